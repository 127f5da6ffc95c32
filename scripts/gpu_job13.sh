#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
P=29811
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 5 --warmup 3 > $O/j13_bench_n2.json 2> $O/j13_bench_n2.err; echo "bench n2 rc=$?"
tail -4 $O/j13_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/j13_ref_n2.json 2> $O/j13_ref_n2.err; echo "ref n2 rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/j13_bench_n2.json"))
print("%.2fM cell-steps/s, %.2f ms/step, it/step %.1f, e2e %.2fM us/iter %.1f" % (d["value"]/1e6, d["ms_per_step"], d["pcg_iterations_per_step"], d["e2e"]["value"]/1e6, d["us_per_pcg_iteration"]))
print("coupled", {k:v for k,v in d["e2e_coupled"].items() if k!="api"}); print("strong", d.get("strong_scaling_leg")); print("parity", d["parity_check"]["ok"], d["config"].get("decomposition"))
r=json.load(open("gpurun_out/j13_ref_n2.json")); print("ref", r["value"], r["cpu_baseline"]["sample"])
PY
