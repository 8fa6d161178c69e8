#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
echo "== sharded commit check on $N GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -5 | tee gpurun_out/multi_gpu_check_$N.log
echo "== bench --gpus $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -1 | tee gpurun_out/bench_$N.json | cut -c1-600
