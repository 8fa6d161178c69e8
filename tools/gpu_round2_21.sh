#!/bin/bash
mkdir -p gpurun_out
echo "== gate tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_prove.py -m gpu -q -x -k "gate or prove or prover or quotient" 2>&1 | grep -E "Error|error|passed|failed|^E " | head -20
for v in "" gate_np gate_mb4 gate_np_mb4; do
  for only in poseidon2_flattened fma; do
    echo "variant=[$v] $only: $(BJ_LIB_VARIANT=$v ONLY=$only timeout 300 python tools/time_gates.py 2>&1 | tail -1 | cut -c40-)"
  done
done
echo "== full timing (default build)"; timeout 600 python tools/time_gates.py > gpurun_out/time_gates4.json 2> gpurun_out/time_gates4.err; tail -3 gpurun_out/time_gates4.err; cat gpurun_out/time_gates4.json
