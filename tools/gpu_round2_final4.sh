#!/bin/bash
# final single-GPU record of round 2: full GPU test suite, smoke, bench (incl. the production-shaped proof), launch list of the
# bench command, ncu of the final gate interpreter on the Poseidon2 flattened gate
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_r2_final.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2_final.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value'],d['clocks'],d['merkle'])
for k in ('prove','prove_non_recursive'):
    p=d[k];print(k,p['seconds'],p['verified'],p['stages_s'])
print(d['prove_production_shape'])"; tail -2 gpurun_out/bench_r2_final.err
