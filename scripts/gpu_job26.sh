#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
lscpu | grep -E "Model name|Socket|NUMA node\(s\)|^CPU\(s\)"
timeout 600 python scripts/gpu_probe_hostcopy.py > $O/j26_hostcopy.log 2>&1; echo rc=$?; cat $O/j26_hostcopy.log
