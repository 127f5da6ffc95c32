#!/bin/bash
# round 2, GPU job 4 (2 GPUs): folded all-reduces -- decomposed parity tests + weak bench at N = 2, fold on / off
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_multirank.py -m gpu -q > $O/j4_mr_tests.log 2>&1; echo "pytest rc=$?"; tail -4 $O/j4_mr_tests.log
P=29611
for fold in 1 0; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 5 --warmup 3 --opt p2p_fold=$fold > $O/j4_bench_n2_fold$fold.json 2> $O/j4_bench_n2_fold$fold.err; echo "bench n2 fold=$fold rc=$?"
P=$((P+1))
done
python - <<'PY'
import json
for f in ("j4_bench_n2_fold1","j4_bench_n2_fold0"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, "%.2fM"%(d["value"]/1e6), "ms/step %.2f"%d["ms_per_step"], "us/iter %.1f"%d["us_per_pcg_iteration"], d.get("parity_check",{}).get("ok"), d.get("pcg_iterations_per_step"))
    except Exception as e: print(f, "ERR", e)
PY
