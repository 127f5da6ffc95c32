#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/j12_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $O/j12_gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/j12_bench.json 2> $O/j12_bench.err; echo "bench rc=$?"
tail -5 $O/j12_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/j12_bench.json")); k=d["roofline"]["kernels"]
print("%.2fM cell-steps/s, %.2f ms/step, it/step %.1f, e2e %.2fM, precondition %.1f us (frac %.3f, dram %.0f GB/s), amul %.1f us" % (d["value"]/1e6, d["ms_per_step"], d["pcg_iterations_per_step"], d["e2e"]["value"]/1e6, k["dic_precondition"]["us"], k["dic_precondition"]["frac_of_measured_peak"], k["dic_precondition"].get("dram_GBps",0), k["amul"]["us"]))
print("coupled", d["e2e_coupled"]); print("strong", d.get("strong_scaling_leg")); print("plugin", {a:b for a,b in d["e2e_plugin"].items() if a!="api"}); print("cpu", d["cpu_baseline"]); print("traffic", d["roofline"]["traffic"])
PY
