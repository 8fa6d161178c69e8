#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r2e.log
echo "== prove_once"; for h in blake2s poseidon2; do timeout 600 python tools/prove_once.py 22 $h 2>&1 | tail -1; done | tee gpurun_out/prove_once_r2e.log
echo "== timing, shipped"; timeout 600 python tools/time_ntt.py | tee gpurun_out/time_ntt_r2_shipped.json | cut -c1-500
echo "== timing, variant subfma"; BJ_LIB_VARIANT=subfma timeout 600 python tools/time_ntt.py | tee gpurun_out/time_ntt_r2_subfma.json | cut -c1-500
