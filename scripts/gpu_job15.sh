#!/bin/bash
# chunked sweep plan (polyhedral rows as chains of <=3-source chunks): GPU tests, polyhedral roofline at 2.56M and 4.05M cells, hex bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/j15_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $O/j15_gpu_tests.log
timeout 900 python scripts/gpu_poly_roofline.py > $O/j15_poly_2.56M.json 2> $O/j15_poly_2.56M.err; echo "poly rc=$?"; cat $O/j15_poly_2.56M.json; tail -3 $O/j15_poly_2.56M.err
timeout 1200 python scripts/gpu_poly_roofline.py 300 180 180 > $O/j15_poly_4M.json 2> $O/j15_poly_4M.err; echo "poly4M rc=$?"; cat $O/j15_poly_4M.json; tail -3 $O/j15_poly_4M.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-strong-leg > $O/j15_bench.json 2> $O/j15_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/j15_bench.json"))
print("%.2fM cell-steps/s, %.2f ms/step, precondition %.1f us" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["kernels"]["dic_precondition"]["us"]))
PY
