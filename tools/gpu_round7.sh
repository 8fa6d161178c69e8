#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
