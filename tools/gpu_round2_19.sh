#!/bin/bash
mkdir -p gpurun_out
for k in 4 1; do
ONLY=poseidon2_flattened BJ_GATE_POINTS_PER_THREAD=$k timeout 900 ncu --set full --clock-control none -k regex:gate_eval -s 1 -c 1 -f -o /tmp/prof_gate_k$k python tools/time_gates.py > gpurun_out/ncu_gate_k$k.log 2>&1
tail -2 gpurun_out/ncu_gate_k$k.log
python tools/ncu_summary.py /tmp/prof_gate_k$k.ncu-rep > gpurun_out/r2_ncu_gate_p2_k${k}_summary.txt 2>&1
rm -f /tmp/prof_gate_k$k.ncu-rep
cat gpurun_out/r2_ncu_gate_p2_k${k}_summary.txt
done
