#!/bin/bash
mkdir -p gpurun_out
echo "== pytest prove"; timeout 1500 python -m pytest tests/test_gpu_prove.py -m gpu -q --timeout 900 -x 2>&1 | tail -40 | tee gpurun_out/pytest_prove.log
