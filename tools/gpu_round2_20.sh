#!/bin/bash
mkdir -p gpurun_out
echo "== gate tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_prove.py -m gpu -q -x -k "gate or prove or prover or quotient" 2>&1 | grep -E "Error|error|passed|failed|^E " | head -20
echo "== gate evaluator timing"; timeout 600 python tools/time_gates.py > gpurun_out/time_gates3.json 2> gpurun_out/time_gates3.err; tail -3 gpurun_out/time_gates3.err; cat gpurun_out/time_gates3.json
for k in 0; do
ONLY=poseidon2_flattened BJ_GATE_POINTS_PER_THREAD=$k timeout 900 ncu --set full --clock-control none -k regex:gate_eval -s 1 -c 1 -f -o /tmp/prof_gate_k$k python tools/time_gates.py > gpurun_out/ncu_gate_k$k.log 2>&1
tail -2 gpurun_out/ncu_gate_k$k.log
python tools/ncu_summary.py /tmp/prof_gate_k$k.ncu-rep > gpurun_out/r2_ncu_gate_p2_smem_summary.txt 2>&1
rm -f /tmp/prof_gate_k$k.ncu-rep
grep -v "hit_rate\|mem_local" gpurun_out/r2_ncu_gate_p2_smem_summary.txt
done
