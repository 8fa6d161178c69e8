#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest (PTX arithmetic)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
if grep -q failed gpurun_out/pytest_gpu.log; then
  echo "== pytest (portable arithmetic)"; BJ_LIB_VARIANT=portable timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_portable.log
fi
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_ntt.csv python tools/prof_ntt.py ntt > gpurun_out/ncu1.log 2>&1; grep ntt_pass gpurun_out/launches_ntt.csv | tail -4 | cut -c1-60,200-400
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 4 -c 2 -f -o gpurun_out/prof_ntt_r2 python tools/prof_ntt.py ntt > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
