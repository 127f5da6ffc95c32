#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 300 python scripts/gpu_probe_trace.py 1 > $O/j9_trace_1M.log 2>&1; head -8 $O/j9_trace_1M.log
timeout 300 python scripts/gpu_probe_trace.py 1 400 100 200 > $O/j9_trace_8M.log 2>&1; head -8 $O/j9_trace_8M.log
timeout 600 python -m pytest tests/test_ldu.py tests/test_pbicg.py tests/test_region.py tests/test_fullsize.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/j9_bench.json 2> $O/j9_bench.err
timeout 300 python bench.py --steps 3 --warmup 2 --scaling strong --no-cpu-baseline > $O/j9_bench_strong.json 2> $O/j9_bench_strong.err
python - <<'PY'
import json
for f in ("j9_bench","j9_bench_strong"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); k=d["roofline"]["kernels"]
        print(f, "%.2fM cell-steps/s, %.2f ms/step, it/step %.1f, precondition %.1f us (frac %.3f), amul %.1f us" % (d["value"]/1e6, d["ms_per_step"], d["pcg_iterations_per_step"], k["dic_precondition"]["us"], k["dic_precondition"]["frac_of_measured_peak"], k["amul"]["us"]))
    except Exception as e: print(f, "ERR", e)
PY
