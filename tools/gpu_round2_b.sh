#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r2b.log
bash tools/ntt_experiments.sh 2>&1 | tail -40
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2b.json'));print(d['value'],d['roofline'],d['ntt_family'],d['e2e']['value'],d.get('merkle'))
for k in ('prove','prove_non_recursive'):
    p=d[k];print(k,p['seconds'],p['verified'],p['python_json_parse_s'],p['stages_s'])
print(d['cpu_baseline']['rows'])"; tail -3 gpurun_out/bench_r2b.err
