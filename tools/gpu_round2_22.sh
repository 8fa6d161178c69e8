#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_prove.py -m gpu -q -x -k "production_shaped or quotient_degree_above or native_cxx" 2>&1 | grep -E "Error|error|passed|failed|^E " | head -20
for h in poseidon2 blake2s; do
  timeout 900 python tools/prove_production_shape.py 20 $h > gpurun_out/production_shape_$h.json 2> gpurun_out/production_shape_$h.err; tail -2 gpurun_out/production_shape_$h.err; cat gpurun_out/production_shape_$h.json
done
