#!/bin/bash
# One GPU validation round (run under gpurun): tests, bench line, ncu launch list of the bench command, full ncu captures
# of the two dominant kernels.  Outputs land in gpurun_out/ (copied into profiles/ by hand afterwards).
mkdir -p gpurun_out
TAG=${TAG:-r}
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'],d['roofline']['frac'],d['sweep'],d['ntt_family'],d['e2e'],d['clocks'],d.get('merkle'),d.get('prove'),d.get('prove_non_recursive'),d.get('cpu_baseline'))"; tail -3 gpurun_out/bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/bench_reference.json | cut -c1-400
echo "== ncu launch list of the bench command"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --prove-log-n 0 > gpurun_out/ncu1.log 2>&1; grep -c ntt_pass gpurun_out/launches_bench.csv
[ "$NCU" = "0" ] && exit 0
echo "== ncu full ntt"; timeout 900 ncu --set full --clock-control none -k regex:ntt_pass -s 4 -c 2 -f -o gpurun_out/prof_ntt_$TAG python tools/prof_ntt.py ntt > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
echo "== ncu full merkle"; timeout 900 ncu --set full --clock-control none -k regex:poseidon2_leaf -c 1 -f -o gpurun_out/prof_merkle_$TAG python tools/prof_ntt.py merkle > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log
