#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 600 python scripts/gpu_probe_trace_poly.py > $O/j20_trace_poly.log 2>&1; echo "poly trace rc=$?"; grep -v Warning $O/j20_trace_poly.log | grep -v "^  "
