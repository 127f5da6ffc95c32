#!/bin/bash
# final-tree validation under torchrun, N GPUs = $1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}
O=gpurun_out; mkdir -p $O
P=29900
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+N)) bench.py --gpus $N --steps 5 --warmup 3 > $O/j28_bench_n$N.json 2> $O/j28_bench_n$N.err; echo "bench n$N rc=$?"
tail -3 $O/j28_bench_n$N.err
python - $N <<'PY'
import json,sys
d=json.load(open("gpurun_out/j28_bench_n%s.json"%sys.argv[1]))
print("%.2fM cell-steps/s, %.2f ms/step, it/step %.1f, e2e %.2fM us/iter %.1f" % (d["value"]/1e6, d["ms_per_step"], d["pcg_iterations_per_step"], d["e2e"]["value"]/1e6, d["us_per_pcg_iteration"]))
print("coupled %.2fM" % (d["e2e_coupled"]["value"]/1e6)); s=d.get("strong_scaling_leg"); print("strong %.2fM %.1f ms/step %s" % (s["value"]/1e6, s["ms_per_step"], s["decomposition"])); print("parity", d["parity_check"]["ok"], d["config"].get("decomposition"))
PY
if [ "$N" = "2" ]; then
timeout 900 python -m pytest tests/test_multirank.py -m gpu -x -q > $O/j28_mr_tests.log 2>&1; echo "multirank pytest rc=$?"; tail -3 $O/j28_mr_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
fi
