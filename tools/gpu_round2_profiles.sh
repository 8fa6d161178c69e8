#!/bin/bash
# ncu --set full captures of the prover's non-NTT kernels at 2^22 rows + GPU tests.  The .ncu-rep files
# (tens of MB each with sources) are summarised ON the box (tools/ncu_summary.py) and removed: gpurun_out/ must stay < 64 MiB.
mkdir -p gpurun_out
if [ "$TESTS" != "0" ]; then
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r2.log
fi
[ "$PROFILES" = "0" ] && exit 0
cap() {  # name regex count hasher
  WARM=0 timeout 1200 ncu --set full --clock-control none -k regex:"$2" -c $3 -f -o /tmp/prof_$1 python tools/prove_once.py 22 $4 > gpurun_out/ncu_$1.log 2>&1
  tail -2 gpurun_out/ncu_$1.log
  python tools/ncu_summary.py /tmp/prof_$1.ncu-rep > gpurun_out/r2_ncu_$1_summary.txt 2>&1
  rm -f /tmp/prof_$1.ncu-rep
}
echo "== ncu full: prover kernels (poseidon2 config)"; cap prover 'deep_group|quotient_copy_perm|gate_eval|copy_perm_ratios|lookup_polys|quotient_lookup' 8 poseidon2
echo "== ncu full: poseidon2 node + leaf"; cap p2 'poseidon2_node|poseidon2_leaf' 2 poseidon2
echo "== ncu full: blake2s leaf + node"; cap b2s 'blake2s_leaf|blake2s_node' 2 blake2s
ls -la gpurun_out/
