#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'],d['roofline']['frac'],d['e2e'],d['clocks'],d['cpu_baseline'])"; tail -5 gpurun_out/bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 | tail -1 | cut -c1-400
