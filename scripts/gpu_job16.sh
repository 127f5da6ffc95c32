#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 600 python scripts/gpu_probe_trace_poly.py > $O/j16_trace_poly.log 2>&1; echo "poly trace rc=$?"; cat $O/j16_trace_poly.log
timeout 600 python scripts/gpu_probe_trace.py 1 > $O/j16_trace_hex.log 2>&1; echo "hex trace rc=$?"; head -8 $O/j16_trace_hex.log
