#!/bin/bash
mkdir -p gpurun_out
echo "== full pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value'],d['cpu_baseline']['value']);print(d.get('merkle'));print(json.dumps(d.get('prove'),indent=1))"; tail -8 gpurun_out/bench.err
