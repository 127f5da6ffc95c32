#!/bin/bash
# round 2, GPU job 5 (8 GPUs): weak + strong scaling at N = 8 (and N = 4 weak) with the parity pre-flight
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/j5_smi.txt
P=29711
run() { n=$1; shift; tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n "$@" > $O/j5_$tag.json 2> $O/j5_$tag.err; echo "$tag rc=$?"; P=$((P+1)); }
run 8 weak_n8 --steps 5 --warmup 3
run 8 strong_n8 --steps 5 --warmup 3 --scaling strong
run 4 weak_n4 --steps 5 --warmup 3
run 4 strong_n4 --steps 5 --warmup 3 --scaling strong
python - <<'PY'
import json
for f in ("j5_weak_n8","j5_strong_n8","j5_weak_n4","j5_strong_n4"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, "%.2fM"%(d["value"]/1e6), "ms/step %.2f"%d["ms_per_step"], "us/iter %.1f"%d["us_per_pcg_iteration"], "it/step", d.get("pcg_iterations_per_step"), "parity", d.get("parity_check",{}).get("ok"))
    except Exception as e: print(f, "ERR", e)
PY
