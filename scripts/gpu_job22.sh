#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
free -g | head -2
(time timeout 1500 python -m pytest tests/test_atsize_region.py -m gpu -q -s) > $O/j22_atsize.log 2>&1; echo "pytest rc=$?"; tail -25 $O/j22_atsize.log
