#!/bin/bash
# A/B of the butterfly pipe balance: shipped (sum's wrap correction on the ALU pipe) vs BJ_LIB_VARIANT=addfma (on the FMA pipe)
mkdir -p gpurun_out
echo "== NTT parity (shipped lib)"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ntt or lde" 2>&1 | tail -3
echo "== timing, add on ALU pipe (shipped)"; timeout 600 python tools/time_ntt.py | tee gpurun_out/time_ntt_r2_addalu.json | cut -c1-700
echo "== timing, add on FMA pipe (variant addfma)"; BJ_LIB_VARIANT=addfma timeout 600 python tools/time_ntt.py | tee gpurun_out/time_ntt_r2_addfma.json | cut -c1-700
timeout 900 ncu --set full --clock-control none -k regex:ntt_pass -s 4 -c 2 -f -o /tmp/prof_ntt_addalu python tools/prof_ntt.py ntt > gpurun_out/ncu_ntt_addalu.log 2>&1
python tools/ncu_summary.py /tmp/prof_ntt_addalu.ncu-rep > gpurun_out/r2_ncu_ntt_addalu_summary.txt 2>&1
rm -f /tmp/prof_ntt_addalu.ncu-rep
grep -E "Kernel Name|time_duration|issue_active|pipe_alu|pipe_fma|stalls \(warps" gpurun_out/r2_ncu_ntt_addalu_summary.txt | cut -c1-220
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
