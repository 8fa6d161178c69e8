#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_prove.py tests/test_proof_fixtures.py -m gpu -q -x -k "gate or prove or prover or quotient or production" 2>&1 | grep -E "Error|error|passed|failed|^E " | head -20
echo "== gate evaluator timing"; timeout 600 python tools/time_gates.py > gpurun_out/time_gates5.json 2> gpurun_out/time_gates5.err; tail -3 gpurun_out/time_gates5.err; cat gpurun_out/time_gates5.json
for h in poseidon2 blake2s; do
  timeout 900 python tools/prove_production_shape.py 20 $h > gpurun_out/production_shape_$h.json 2> gpurun_out/production_shape_$h.err; tail -2 gpurun_out/production_shape_$h.err; cat gpurun_out/production_shape_$h.json
done
