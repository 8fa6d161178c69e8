#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r2_final.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
