#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
nproc

for t in 1 4 1 4; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-strong-leg --no-poly-leg --no-preflight --opt host_copy_threads=$t > $O/j25_bench_t$t.json 2> $O/j25_bench_t$t.err; echo "bench t=$t rc=$?"
python - $t <<'PY'
import json,sys
d=json.load(open("gpurun_out/j25_bench_t%s.json"%sys.argv[1]))
p=d["e2e_plugin"]
print("threads %s: value %.2fM e2e %.2fM plugin %.1f ms/step (resident %.1f): solves %s mules %s" % (sys.argv[1], d["value"]/1e6, d["e2e"]["value"]/1e6, p["ms_per_step_in_plugin_calls"], p["resident_ms_same_calls"], [round(x,1) for x in p["pcg_solve_ms"]], [round(x,1) for x in p["mules_ms"]]))
PY
done
