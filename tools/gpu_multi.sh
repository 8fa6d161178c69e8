#!/bin/bash
# multi-GPU validation round (run under gpurun --gpus N): sharded-prover tests on one GPU (local transport), then the NCCL
# path on N GPUs (tools/multi_gpu_check.py), then bench.py --gpus N.
mkdir -p gpurun_out
N=${NGPU:-2}
if [ "$TESTS" != "0" ]; then
echo "== pytest tests/test_gpu_prove.py (1 GPU, local transport)"; timeout 900 python -m pytest tests/test_gpu_prove.py -m gpu -q -x --timeout 600 2>&1 | tail -5 | tee gpurun_out/pytest_prove_$N.log
fi
echo "== sharded commit / prover check on $N GPUs"
PROVE_LOG_N=${PROVE_LOG_N:-20} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -8 | tee gpurun_out/multi_gpu_check_$N.log
echo "== bench --gpus $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 2> gpurun_out/bench_$N.err | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -1 | tee gpurun_out/bench_$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'n_gpus')}, d['e2e']['value'])
for k in ('prove', 'prove_non_recursive'):
    p = d[k]; print(k, p['seconds'], p['verified'], p['stages_s'])
"
tail -3 gpurun_out/bench_$N.err
