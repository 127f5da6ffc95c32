#!/bin/bash
# final-build ncu evidence: launch list of the bench command + one --set full capture of the PCG kernels (numbers under ncu are never bench values)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-preflight --no-strong-leg --no-poly-leg"
timeout 300 $CMD > $O/j24_plain.json 2> $O/j24_plain.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1500 --csv --log-file $O/j24_launches.csv $CMD > $O/j24_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_dic_fwd|k_dic_bwd|k_amul|k_pcg_p_update' -s 400 -c 8 -o $O/j24_prof $CMD > $O/j24_ncu2.log 2>&1; echo "ncu2 rc=$?"
ncu -i $O/j24_prof.ncu-rep --page raw --csv > $O/j24_full_raw.csv 2>/dev/null; echo "export rc=$?"
ls -la $O | grep j24
