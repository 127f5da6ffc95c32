#!/bin/bash
# round 2, GPU job 2 (1 GPU): all parity tests incl. non-orth / PBiCG / adapters, bench with the N-way CPU baseline, reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
nproc > $O/j2_nproc.txt; lscpu | head -20 >> $O/j2_nproc.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/j2_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $O/j2_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/j2_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 > $O/j2_bench.json 2> $O/j2_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > $O/j2_ref.json 2> $O/j2_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 3 --warmup 2 --scaling strong --no-cpu-baseline > $O/j2_bench_strong1.json 2> $O/j2_bench_strong1.err; echo "strong rc=$?"
head -c 400 $O/j2_bench.json; echo; head -c 600 $O/j2_ref.json; echo
