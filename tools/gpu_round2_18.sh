#!/bin/bash
mkdir -p gpurun_out
echo "== gate tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_prove.py -m gpu -q -x -k "gate or prove or prover or quotient" 2>&1 | grep -E "Error|error|passed|failed|^E " | head -20
echo "== gate evaluator timing"; timeout 600 python tools/time_gates.py > gpurun_out/time_gates2.json 2> gpurun_out/time_gates2.err; tail -3 gpurun_out/time_gates2.err; cat gpurun_out/time_gates2.json
