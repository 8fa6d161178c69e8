#!/bin/bash
# NTT experiments of round 2: bulk-copy (TMA) staged pass vs register-staged pass, L2-resident vs HBM batch.
mkdir -p gpurun_out
echo "== parity of the bulk-copy pass"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bulk_copy" 2>&1 | tail -3
echo "== timing, register-staged (shipped)"; BJ_NTT_BULK=0 timeout 600 python tools/time_ntt.py | tee gpurun_out/time_ntt_r2_base.json
echo "== timing, bulk-copy staged"; BJ_NTT_BULK=1 timeout 600 python tools/time_ntt.py | tee gpurun_out/time_ntt_r2_bulk.json
for v in 0 1; do
  BJ_NTT_BULK=$v timeout 900 ncu --set full --clock-control none -k regex:ntt_pass -s 4 -c 2 -f -o /tmp/prof_ntt_bulk$v python tools/prof_ntt.py ntt > gpurun_out/ncu_ntt_bulk$v.log 2>&1
  python tools/ncu_summary.py /tmp/prof_ntt_bulk$v.ncu-rep > gpurun_out/r2_ncu_ntt_bulk${v}_summary.txt 2>&1
  rm -f /tmp/prof_ntt_bulk$v.ncu-rep
done
grep -E "Kernel Name|time_duration|dram__bytes|issue_active|registers_per|warps_active" gpurun_out/r2_ncu_ntt_bulk*_summary.txt | cut -c1-170
