#!/bin/bash
mkdir -p gpurun_out
echo "== NTT parity quick"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ntt or lde" 2>&1 | tail -3
for v in 0 1; do
  echo "== timing, BJ_NTT_L2_PERSIST=$v"; BJ_NTT_L2_PERSIST=$v timeout 600 python tools/time_ntt.py | tee gpurun_out/time_ntt_r2_l2persist$v.json | cut -c1-600
  BJ_NTT_L2_PERSIST=$v timeout 900 ncu --set full --clock-control none -k regex:ntt_pass -s 4 -c 2 -f -o /tmp/prof_ntt_l2p$v python tools/prof_ntt.py ntt > gpurun_out/ncu_ntt_l2p$v.log 2>&1
  python tools/ncu_summary.py /tmp/prof_ntt_l2p$v.ncu-rep > gpurun_out/r2_ncu_ntt_l2persist${v}_summary.txt 2>&1
  rm -f /tmp/prof_ntt_l2p$v.ncu-rep
done
grep -E "Kernel Name|time_duration|dram__bytes|issue_active|stalls \(warps" gpurun_out/r2_ncu_ntt_l2persist*_summary.txt | cut -c1-220
