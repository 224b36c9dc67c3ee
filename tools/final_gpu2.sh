#!/bin/bash
# Second GPU-box call of the mixed-kernel work: GPU tests, ncu --set full of the production kernel, bench in both
# arithmetic modes, the other BASELINE configs at full size. Outputs under gpurun_out/.
set -u
T=${1:-final2}
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 400 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_$T.json
timeout 200 python bench.py --arithmetic fp64 --no-cpu-baseline --no-carrington > gpurun_out/bench_fp64_$T.json 2> gpurun_out/bench_fp64_$T.err; echo "bench fp64 rc=$?"; cut -c1-300 gpurun_out/bench_fp64_$T.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lag_corr_roll_kernel -c 1 -f -o gpurun_out/prof_mixed_$T \
  python tools/mixed_lab.py --combos mixed:1 --lagsets config1 --steps 1 --out gpurun_out/mixed_ncu_$T.json > gpurun_out/ncu_mixed_$T.log 2>&1; echo "ncu rc=$?"
timeout 400 python tools/config_runs.py --skip carrington --out gpurun_out/config_runs_$T.json 2>&1 | tail -15
