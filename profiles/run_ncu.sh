#!/bin/bash
# Run on the GPU box (under gpurun) from the repo root. Writes into gpurun_out/.
#   launches.csv : every kernel launch of a short bench run with its device time (cold-cache, serialised)
#   rollout.ncu-rep : full-set capture of the rollout kernel (Philox mode = the bench step; FORCED mode = the rules-only run)
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --sections rollout"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_bench_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_pair_kernel -s 3 -c 2 -f -o gpurun_out/rollout $CMD > gpurun_out/ncu_bench_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rollout_pair_kernel<2" -s 3 -c 1 -f -o gpurun_out/rollout_forced $CMD > gpurun_out/ncu_bench_forced.log 2>&1
ls -la gpurun_out
