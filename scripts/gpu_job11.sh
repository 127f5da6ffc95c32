#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 300 python scripts/gpu_probe_trace.py 2 > $O/j11_trace_hybrid.log 2>&1; head -10 $O/j11_trace_hybrid.log
timeout 300 python -m pytest tests/test_ldu.py -m gpu -x -q -k "publish or tiling" 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --opt sweep_publish=2 > $O/j11_bench_pub2.json 2> $O/j11_bench_pub2.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/j11_bench_pub1.json 2> $O/j11_bench_pub1.err
timeout 300 python bench.py --steps 3 --warmup 2 --scaling strong --no-cpu-baseline --opt sweep_publish=2 > $O/j11_bench_strong2.json 2> $O/j11_bench_strong2.err
python - <<'PY'
import json
for f in ("j11_bench_pub1","j11_bench_pub2","j11_bench_strong2"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); k=d["roofline"]["kernels"]
        print(f, "%.2fM cell-steps/s, %.2f ms/step, it/step %.1f, precondition %.1f us (frac %.3f), amul %.1f us" % (d["value"]/1e6, d["ms_per_step"], d["pcg_iterations_per_step"], k["dic_precondition"]["us"], k["dic_precondition"]["frac_of_measured_peak"], k["amul"]["us"]))
    except Exception as e: print(f, "ERR", e)
PY
