#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'],d['roofline']['frac'],d['sweep'],d['e2e'],d['clocks'])"; tail -5 gpurun_out/bench.err
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_ntt.csv python tools/prof_ntt.py ntt > gpurun_out/ncu1.log 2>&1; grep ntt_pass gpurun_out/launches_ntt.csv | tail -2 | awk -F'","' '{print $5, $8, $9, $NF}'
echo "== ncu full ntt"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 4 -c 2 -f -o gpurun_out/prof_ntt_r4 python tools/prof_ntt.py ntt > gpurun_out/ncu2.log 2>&1; tail -1 gpurun_out/ncu2.log
echo "== ncu full merkle"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:poseidon2_leaf -c 1 -f -o gpurun_out/prof_merkle_r4 python tools/prof_ntt.py merkle > gpurun_out/ncu3.log 2>&1; tail -1 gpurun_out/ncu3.log
