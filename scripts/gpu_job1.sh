#!/bin/bash
# round 2, GPU job 1: parity tests, register-operand sweeps A/B, new bench flow, ncu launch list + full capture of the PCG kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/j1_smi.txt 2>&1
timeout 300 python scripts/gpu_probe_trace.py 1 > $O/j1_trace_reg1.log 2>&1; echo "trace rc=$?"
timeout 300 python scripts/gpu_probe_trace.py 0 > $O/j1_trace_reg0.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/j1_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $O/j1_gpu_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-budget 20 > $O/j1_bench.json 2> $O/j1_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --opt sweep_reg_ops=0 > $O/j1_bench_reg0.json 2> $O/j1_bench_reg0.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --delta-t 0.002 > $O/j1_bench_dt002.json 2> $O/j1_bench_dt002.err
timeout 300 python scripts/gpu_step_profile.py > $O/j1_stepprof.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1500 --csv --log-file $O/j1_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/j1_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_dic_fwd|k_dic_bwd|k_amul|k_pcg_p_update' -s 400 -c 8 -o $O/j1_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/j1_ncu2.log 2>&1; echo "ncu2 rc=$?"
head -c 600 $O/j1_bench.json; echo
