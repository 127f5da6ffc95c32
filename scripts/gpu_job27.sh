#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_ldu.py tests/test_region.py tests/test_fullsize.py tests/test_golden.py -m gpu -x -q > $O/j27_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $O/j27_gpu_tests.log
for f in 0 1 0 1; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-poly-leg --no-preflight --opt pcg_fuse_p=$f > $O/j27_bench_f$f.json 2> $O/j27_bench_f$f.err; echo "bench fuse=$f rc=$?"
python - $f <<'PY'
import json,sys
d=json.load(open("gpurun_out/j27_bench_f%s.json"%sys.argv[1]))
s=d["strong_scaling_leg"]
print("fuse %s: value %.2fM (%.2f ms/step, %.1f it/step, %.1f us/it) e2e %.2fM | 7.88M: %.2fM %.1f ms/step %.1f us/it" % (sys.argv[1], d["value"]/1e6, d["ms_per_step"], d["pcg_iterations_per_step"], d["us_per_pcg_iteration"], d["e2e"]["value"]/1e6, s["value"]/1e6, s["ms_per_step"], s["us_per_pcg_iteration"]))
PY
done
