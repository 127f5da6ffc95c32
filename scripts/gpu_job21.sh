#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python scripts/gpu_poly_roofline.py > $O/j21_poly_2.56M.json 2> $O/j21_poly_2.56M.err; echo "poly rc=$?"; cat $O/j21_poly_2.56M.json; tail -3 $O/j21_poly_2.56M.err
timeout 1200 python scripts/gpu_poly_roofline.py 300 180 180 > $O/j21_poly_4M.json 2> $O/j21_poly_4M.err; echo "poly4M rc=$?"; cat $O/j21_poly_4M.json; tail -3 $O/j21_poly_4M.err
timeout 900 python -m pytest tests -m gpu -x -q > $O/j21_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $O/j21_gpu_tests.log
