#!/bin/bash
# final single-GPU record of round 2: bench line, reference arm, ncu launch list of the bench command, ncu --set full of the
# two changed copy-permutation kernels
mkdir -p gpurun_out
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2_final.json'));print(d['value'],d['roofline'],d['e2e']['value'],d['clocks'])
for k in ('prove','prove_non_recursive'):
    p=d[k];print(k,p['seconds'],p['verified'],p['stages_s'])"; tail -2 gpurun_out/bench_r2_final.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/bench_r2_reference_arm.json | cut -c1-300
echo "== ncu launch list of the bench command"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --prove-log-n 0 > gpurun_out/ncu_launches.log 2>&1; grep -c ntt_pass gpurun_out/r2_launches_bench.csv
echo "== ncu full: copy-permutation kernels after the change"
WARM=0 timeout 1200 ncu --set full --clock-control none -k regex:'quotient_copy_perm|copy_perm_ratios' -c 2 -f -o /tmp/prof_cp python tools/prove_once.py 22 blake2s > gpurun_out/ncu_cp.log 2>&1
python tools/ncu_summary.py /tmp/prof_cp.ncu-rep > gpurun_out/r2_ncu_copy_perm_after_summary.txt 2>&1; rm -f /tmp/prof_cp.ncu-rep
grep -E "Kernel Name|time_duration|pipe_alu|pipe_fma|issue_active|stalls \(warps" gpurun_out/r2_ncu_copy_perm_after_summary.txt | cut -c1-200
