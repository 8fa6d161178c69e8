#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_r2_final.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== merkle"; timeout 300 python tools/time_merkle.py | tee gpurun_out/time_merkle_r2_final.json
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2_final.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value'],d['clocks'],d['merkle'])
for k in ('prove','prove_non_recursive'):
    p=d[k];print(k,p['seconds'],p['verified'],p['stages_s'])"; tail -2 gpurun_out/bench_r2_final.err
echo "== ncu launch list of the bench command"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --prove-log-n 0 > gpurun_out/ncu_launches.log 2>&1; grep -c ntt_pass gpurun_out/r2_launches_bench.csv
echo "== ncu full: poseidon2 leaf after the change"
timeout 900 ncu --set full --clock-control none -k regex:'poseidon2_leaf' -c 1 -f -o /tmp/prof_p2 python tools/prof_ntt.py merkle > gpurun_out/ncu_p2b.log 2>&1
python tools/ncu_summary.py /tmp/prof_p2.ncu-rep > gpurun_out/r2_ncu_p2_after_summary.txt 2>&1; rm -f /tmp/prof_p2.ncu-rep
grep -E "Kernel Name|time_duration|inst_executed|pipe_alu|pipe_fma|issue_active|stalls \(warps" gpurun_out/r2_ncu_p2_after_summary.txt | cut -c1-220
