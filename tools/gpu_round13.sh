#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(d['value']);print(d.get('merkle'));print(json.dumps(d.get('prove')['stages_s']), d['prove']['seconds'])"; tail -8 gpurun_out/bench.err
echo "== ncu merkle"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:poseidon2_leaf -c 1 -f -o gpurun_out/prof_merkle_r13 python tools/prof_ntt.py merkle > gpurun_out/ncu3.log 2>&1; tail -1 gpurun_out/ncu3.log
