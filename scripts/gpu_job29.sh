#!/bin/bash
# side-stream halo push A/B on 2 GPUs (option p2p_pdl_iface) + decomposed parity
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
P=29950
timeout 900 python -m pytest tests/test_multirank.py -m gpu -x -q > $O/j29_mr_tests.log 2>&1; echo "multirank pytest rc=$?"; tail -2 $O/j29_mr_tests.log
i=0
for f in 0 1 0 1; do
i=$((i+1))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+i)) bench.py --gpus 2 --steps 5 --warmup 3 --no-strong-leg --opt pcg_pdl_iface=$f > $O/j29_bench_f$f.json 2> $O/j29_bench_f$f.err; echo "bench pdl_iface=$f rc=$?"
python - $f <<'PY'
import json,sys
d=json.load(open("gpurun_out/j29_bench_f%s.json"%sys.argv[1]))
print("pdl_iface %s: %.2fM cell-steps/s, %.2f ms/step, it/step %.1f, us/iter %.1f, parity %s" % (sys.argv[1], d["value"]/1e6, d["ms_per_step"], d["pcg_iterations_per_step"], d["us_per_pcg_iteration"], d["parity_check"]["ok"]))
PY
done
