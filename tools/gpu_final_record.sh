#!/bin/bash
# final single-GPU record of round 2: full GPU test suite, smoke, bench (incl. the production-shaped proof), launch list of the
# bench command, ncu of the final gate interpreter on the Poseidon2 flattened gate
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_r2_final.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2_final.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value'],d['clocks'],d['merkle'])
for k in ('prove','prove_non_recursive'):
    p=d[k];print(k,p['seconds'],p['verified'],p['stages_s'])
print(d['prove_production_shape'])"; tail -2 gpurun_out/bench_r2_final.err
echo "== ncu launch list of the bench command"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --prove-log-n 0 > gpurun_out/ncu_launches.log 2>&1; grep -c ntt_pass gpurun_out/r2_launches_bench.csv
echo "== ncu full: gate interpreter (final), Poseidon2 flattened gate"
ONLY=poseidon2_flattened timeout 900 ncu --set full --clock-control none -k regex:gate_eval -s 1 -c 1 -f -o /tmp/prof_gate python tools/time_gates.py > gpurun_out/ncu_gate_final.log 2>&1
python tools/ncu_summary.py /tmp/prof_gate.ncu-rep > gpurun_out/r2_ncu_gate_p2_final_summary.txt 2>&1; rm -f /tmp/prof_gate.ncu-rep
grep -E "Kernel Name|time_duration|inst_executed.sum|pipe_alu|pipe_fma|issue_active|local_op_ld_hit|dram__bytes|stalls \(warps" gpurun_out/r2_ncu_gate_p2_final_summary.txt | cut -c1-220
