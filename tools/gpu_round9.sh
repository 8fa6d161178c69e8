#!/bin/bash
mkdir -p gpurun_out
echo "== pytest prove"; timeout 900 python -m pytest tests/test_gpu_prove.py -m gpu -q --timeout 900 2>&1 | tail -5
echo "== bench with prove"; timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'],d['e2e']['value']);print(json.dumps(d.get('prove'),indent=1))"; tail -8 gpurun_out/bench.err
