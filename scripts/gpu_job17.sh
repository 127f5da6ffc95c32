#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for i in 1 2; do
for l in ab/libb200foam_old.so shallowinterfoam_b200/libb200foam.so; do
timeout 300 python scripts/gpu_ab_sweeps.py $l 2>&1 | tail -2
done
done
