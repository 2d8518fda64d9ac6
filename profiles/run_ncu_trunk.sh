#!/bin/bash
# Full-set ncu capture of the fused trunk kernel (SLPolicy forward, 16,384 positions, precision 3). Run under gpurun.
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:trunk_kernel -s 3 -c 1 -f -o gpurun_out/trunk python tools/bench_nets.py --steps 3 > gpurun_out/ncu_trunk.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mcts_select_pipe_kernel -s 40 -c 2 -f -o gpurun_out/mcts_select python tools/bench_mcts.py --trees 256 --playouts 8192 --warm 0 > gpurun_out/ncu_mcts_select.log 2>&1
ls -la gpurun_out
