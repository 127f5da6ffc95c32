#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
P=29821
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+N)) bench.py --gpus $N --steps 5 --warmup 3 > $O/j14_bench_n$N.json 2> $O/j14_bench_n$N.err; echo "bench n$N rc=$?"
tail -3 $O/j14_bench_n$N.err
python - $N <<'PY'
import json,sys
d=json.load(open("gpurun_out/j14_bench_n%s.json"%sys.argv[1]))
print("%.2fM cell-steps/s, %.2f ms/step, it/step %.1f, e2e %.2fM us/iter %.1f" % (d["value"]/1e6, d["ms_per_step"], d["pcg_iterations_per_step"], d["e2e"]["value"]/1e6, d["us_per_pcg_iteration"]))
print("coupled", {k:v for k,v in d["e2e_coupled"].items() if k!="api"}); print("strong", d.get("strong_scaling_leg")); print("parity", d["parity_check"]["ok"], d["config"].get("decomposition"))
PY
done
