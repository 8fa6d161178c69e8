#!/bin/bash
for m in 0 1 2 3; do
echo "== BJ_GATE_PEEPHOLE=$m"; BJ_GATE_PEEPHOLE=$m timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gate" 2>&1 | grep -E "passed|failed|^FAILED" | head -8
done
