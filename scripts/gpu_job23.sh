#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
(time timeout 900 python bench.py --steps 5 --warmup 3 > $O/j23_bench.json 2> $O/j23_bench.err); echo "bench rc=$?"; tail -3 $O/j23_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/j23_bench.json"))
print("%.2fM cell-steps/s, %.2f ms/step, precondition %.1f us" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["kernels"]["dic_precondition"]["us"]))
print("poly", d.get("polyhedral_leg"))
print("strong", d.get("strong_scaling_leg",{}).get("value"))
PY
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/j23_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 $O/j23_gpu_tests.log
