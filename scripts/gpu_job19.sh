#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 1200 python scripts/gpu_probe_poly_params.py > $O/j19_poly_params.log 2>&1; echo "rc=$?"; cat $O/j19_poly_params.log
