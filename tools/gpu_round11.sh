#!/bin/bash
mkdir -p gpurun_out
echo "== pytest blake2s + prove"; timeout 1500 python -m pytest tests/test_gpu_prove.py tests/test_gpu_parity.py -m gpu -q --timeout 900 -k "blake2s or prove or lookup" 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
