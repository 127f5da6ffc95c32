#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
for l in ab/libb200foam_old.so shallowinterfoam_b200/libb200foam.so ab/libb200foam_old.so shallowinterfoam_b200/libb200foam.so; do
timeout 300 python scripts/gpu_ab_sweeps.py $l 2>&1 | tail -2
done
timeout 600 python scripts/gpu_probe_trace_poly.py > $O/j18_trace_poly.log 2>&1; echo "poly trace rc=$?"; grep -v Warning $O/j18_trace_poly.log | grep -v "^  "
