#!/bin/bash
# round 2, GPU job 3 (2 GPUs): all parity tests (incl. the 2-GPU decomposed ones), weak + strong bench at N = 2 with the parity pre-flight
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/j3_smi.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/j3_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 $O/j3_gpu_tests.log
P=29511
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 5 --warmup 3 > $O/j3_bench_n2.json 2> $O/j3_bench_n2.err; echo "bench n2 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 2 --steps 3 --warmup 2 --scaling strong > $O/j3_bench_strong_n2.json 2> $O/j3_bench_strong_n2.err; echo "strong n2 rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/j3_bench_n1.json 2> $O/j3_bench_n1.err; echo "bench n1 rc=$?"
head -c 300 $O/j3_bench_n2.json; echo; python - <<'PY'
import json
for f in ("j3_bench_n2","j3_bench_strong_n2"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d.get("value"), d.get("parity_check",{}).get("ok"), d.get("pcg_iterations_per_step"))
    except Exception as e: print(f, "ERR", e)
PY
