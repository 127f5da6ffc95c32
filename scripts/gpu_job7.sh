#!/bin/bash
# round 2, job 7: sweep lanes / tiles at 7.88 M cells (throughput-bound regime): 2 CTAs of 128 lanes per SM vs 1 CTA of 256
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
for cfg in "256 4096" "128 2048" "192 3072" "160 2560"; do set -- $cfg
timeout 300 python bench.py --steps 2 --warmup 1 --scaling strong --no-cpu-baseline --opt sweep_threads=$1 --opt sweep_tile_rows=$2 > $O/j7_strong_$1.json 2> $O/j7_strong_$1.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/j7_strong_$1.json")); k=d["roofline"]["kernels"]
    print("lanes $1 rows $2: %.2fM cell-steps/s, %.1f ms/step, precondition %.1f us (frac %.3f), amul %.1f us" % (d["value"]/1e6, d["ms_per_step"], k["dic_precondition"]["us"], k["dic_precondition"]["frac_of_measured_peak"], k["amul"]["us"]))
except Exception as e: print("lanes $1 ERR", e)
PY
done
