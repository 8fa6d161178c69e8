#!/bin/bash
# gate interpreter: parity tests, proof equality with the CPU oracle prover, timing on the fixture's gate set
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_prove.py -m gpu -q -x -k "gate or cpu_oracle_prover or production_shaped or native_cxx" 2>&1 | grep -E "Error|error|passed|failed|^E " | head
timeout 300 python tools/time_gates.py > gpurun_out/time_gates6.json 2> gpurun_out/time_gates6.err; tail -2 gpurun_out/time_gates6.err; python -c "
import json;d=json.load(open('gpurun_out/time_gates6.json'));print({k:(v['ms'] if isinstance(v,dict) else v) for k,v in d.items()})"
