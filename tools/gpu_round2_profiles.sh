#!/bin/bash
# ncu --set full captures of the prover's non-NTT kernels at 2^22 rows + GPU tests.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r2.log
echo "== ncu full: prover kernels (poseidon2 config)"
WARM=0 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'deep_group|quotient_copy_perm|gate_eval|copy_perm_ratios|lookup_polys|quotient_lookup' -c 8 -f -o gpurun_out/prof_prover_r2 python tools/prove_once.py 22 poseidon2 > gpurun_out/ncu_prover.log 2>&1; tail -2 gpurun_out/ncu_prover.log
echo "== ncu full: poseidon2 node + leaf"
WARM=0 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'poseidon2_node|poseidon2_leaf' -c 2 -f -o gpurun_out/prof_p2_r2 python tools/prove_once.py 22 poseidon2 > gpurun_out/ncu_p2.log 2>&1; tail -2 gpurun_out/ncu_p2.log
echo "== ncu full: blake2s leaf + node"
WARM=0 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'blake2s_leaf|blake2s_node' -c 2 -f -o gpurun_out/prof_b2s_r2 python tools/prove_once.py 22 blake2s > gpurun_out/ncu_b2s.log 2>&1; tail -2 gpurun_out/ncu_b2s.log
ls -la gpurun_out/*.ncu-rep
