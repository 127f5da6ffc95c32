#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_ldu.py tests/test_region.py tests/test_fullsize.py -m gpu -x -q > $O/j30_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -2 $O/j30_gpu_tests.log
for i in 1 2; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-poly-leg --no-preflight > $O/j30_bench_$i.json 2> $O/j30_bench_$i.err; echo "bench rc=$?"
python - $i <<'PY'
import json,sys
d=json.load(open("gpurun_out/j30_bench_%s.json"%sys.argv[1]))
k=d["roofline"]["kernels"]; s=d["strong_scaling_leg"]
print("value %.2fM (%.2f ms/step, %.1f us/it) e2e %.2fM | pcg_vec %.1f us amul %.1f precond %.1f | 7.88M: %.2fM %.1f us/it" % (d["value"]/1e6, d["ms_per_step"], d["us_per_pcg_iteration"], d["e2e"]["value"]/1e6, k["pcg_vec"]["us"], k["amul"]["us"], k["dic_precondition"]["us"], s["value"]/1e6, s["us_per_pcg_iteration"]))
PY
done
