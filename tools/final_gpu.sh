#!/bin/bash
# One GPU-box call: A/B lab, GPU tests, bench (own arm), launch list of the bench command. Outputs under gpurun_out/.
set -u
T=${1:-final}
timeout 200 python tools/mixed_lab.py --combos fp64:0,mixed:1,mixedc:1,mixed:0,mixedc:0 --out gpurun_out/mixed_lab_$T.json 2>&1 | cut -c1-250 | tail -12
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 400 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_$T.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-carrington > gpurun_out/ncu_launches_$T.log 2>&1; echo "ncu rc=$?"
