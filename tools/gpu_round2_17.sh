#!/bin/bash
mkdir -p gpurun_out
echo "== Q > L tests"; timeout 900 python -m pytest tests/test_gpu_prove.py tests/test_proof_fixtures.py -m gpu -q 2>&1 | grep -E "Error|error|passed|failed|^E " | head -40
