#!/bin/bash
# compute-sanitizer passes over a subset of the GPU tests: memcheck (out-of-bounds / misaligned accesses, leaks of the runtime
# API) on small proofs, the sharded native prover, the gate interpreter and the NTT family; racecheck (shared-memory hazards) on
# the NTT pass kernels.
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
SEL_MEM='test_native_cxx_prover_equals_python_driver or test_gates_over_specialized or (test_native_sharded_prover_equals_single_gpu and 4-8-60) or test_ntt_config1 or test_lde_matches_oracle or test_gate or test_merkle_blake2s or test_lookup_polys or test_do_fri or test_production_shaped or test_quotient_degree_above'
echo "== memcheck"; timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/sanitizer_memcheck.log python -m pytest tests/test_gpu_prove.py tests/test_gpu_parity.py -m gpu -q -x -k "$SEL_MEM" 2>&1 | tail -3
echo "exit: $?"; grep -E "ERROR SUMMARY|Invalid|misaligned|out of bounds" gpurun_out/sanitizer_memcheck.log | sort | uniq -c | head -10
echo "== racecheck (NTT)"; BJ_NTT_BULK=0 timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file gpurun_out/sanitizer_racecheck.log python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_ntt_config1 or (test_ntt_forward_matches_oracle and (13 or 17)) or (test_ntt_inverse_matches_oracle and 14)" 2>&1 | tail -3
grep -E "RACECHECK SUMMARY|hazard" gpurun_out/sanitizer_racecheck.log | sort | uniq -c | head -10
tail -c 1500 gpurun_out/sanitizer_memcheck.log > gpurun_out/sanitizer_memcheck_tail.txt; tail -c 1500 gpurun_out/sanitizer_racecheck.log > gpurun_out/sanitizer_racecheck_tail.txt
