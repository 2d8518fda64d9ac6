#!/bin/bash
# Run on the GPU box (under gpurun) from the repo root: launch list of one REINFORCE gradient + full captures of the backward chain
# (trunk_kernel<1>) and of the tensor-core weight-gradient kernel.
set -x
mkdir -p gpurun_out
CMD="python tools/bench_reinforce.py --positions 8192 --games 64"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_k6.csv $CMD > gpurun_out/ncu_k6_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"trunk_kernel|wgrad_tc_kernel" -s 160 -c 6 -f -o gpurun_out/k6 $CMD > gpurun_out/ncu_k6_full.log 2>&1
ls -la gpurun_out
