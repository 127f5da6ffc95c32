#!/bin/bash
# final tree: full GPU suite, default bench line (all legs), reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/j31_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -6 $O/j31_gpu_tests.log
(time timeout 900 python bench.py --steps 5 --warmup 3 > $O/j31_bench.json 2> $O/j31_bench.err) 2> $O/j31_bench.time; echo "bench rc=$?"; grep real $O/j31_bench.time
(time timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > $O/j31_ref.json 2> $O/j31_ref.err) 2> $O/j31_ref.time; echo "ref rc=$?"; grep real $O/j31_ref.time
python - <<'PY'
import json
d=json.load(open("gpurun_out/j31_bench.json")); r=json.load(open("gpurun_out/j31_ref.json"))
k=d["roofline"]["kernels"]
print("value %.2fM (%.2f ms/step, %.1f it/step, %.1f us/it) e2e %.2fM coupled %.2fM" % (d["value"]/1e6, d["ms_per_step"], d["pcg_iterations_per_step"], d["us_per_pcg_iteration"], d["e2e"]["value"]/1e6, d["e2e_coupled"]["value"]/1e6))
print("roofline frac %.4f achieved %.1f traffic %s | precond %.1f amul %.1f vec %.1f mules %.1f" % (d["roofline"]["frac"], d["roofline"]["achieved"], d["roofline"]["traffic"], k["dic_precondition"]["us"], k["amul"]["us"], k["pcg_vec"]["us"], k["mules_explicitSolve"]["us"]))
p=d["e2e_plugin"]; print("plugin %.1f ms (resident %.1f) solves %s mules %s" % (p["ms_per_step_in_plugin_calls"], p["resident_ms_same_calls"], [round(x,1) for x in p["pcg_solve_ms"]], [round(x,1) for x in p["mules_ms"]]))
s=d["strong_scaling_leg"]; print("strong %.2fM %.1f ms/step %.1f us/it" % (s["value"]/1e6, s["ms_per_step"], s["us_per_pcg_iteration"]))
q=d["polyhedral_leg"]; print("poly amul %.1f us (%.3f) precond %.1f us (%.3f) its %d" % (q["amul"]["us"], q["amul"]["frac_of_measured_peak"], q["dic_precondition"]["us"], q["dic_precondition"]["frac_of_measured_peak"], q["pcg_iterations"]))
c=d["cpu_baseline"]; print("cpu %.2fM on %d cores; ref arm %.2fM (%s cores)" % (c["value"]/1e6, c["cores"], r["value"]/1e6, r["cpu_baseline"]["cores"]))
print("launches", d["gpu_launches"], "clocks", d["clocks"])
PY
