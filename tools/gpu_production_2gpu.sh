#!/bin/bash
mkdir -p gpurun_out
for h in poseidon2 blake2s; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/prove_production_shape.py 20 $h > gpurun_out/production_shape_2gpu_$h.json 2> gpurun_out/production_shape_2gpu_$h.err
  tail -3 gpurun_out/production_shape_2gpu_$h.err | cut -c1-300; cat gpurun_out/production_shape_2gpu_$h.json
done
