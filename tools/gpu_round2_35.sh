#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_prove.py -m gpu -q -x -k "cpu_oracle_prover" 2>&1 | grep -E "Error|error|passed|failed|^E " | head -20
