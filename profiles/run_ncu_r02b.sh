#!/bin/bash
# Round-2 (second half) ncu captures of the rollout kernel after the bank-strip / warp-uniform rewrite.  Run under gpurun from the repo
# root; reports land in gpurun_out/, summaries are made here with profiles/ncu_summary.py and profiles/ncu_traffic.py.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --sections rollout"
EXTRA="--metrics sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed.sum,sm__cycles_active.avg"
ncu --set full $EXTRA --clock-control none --import-source on --kernel-name-base demangled -k "regex:rollout_pair_kernel<\(int\)0, \(bool\)0" -s 3 -c 1 -f -o gpurun_out/r02b_rollout_philox $CMD > gpurun_out/r02b_ncu_rollout.log 2>&1
ncu --set full $EXTRA --clock-control none --import-source on --kernel-name-base demangled -k "regex:rollout_pair_kernel<\(int\)2" -s 1 -c 1 -f -o gpurun_out/r02b_rollout_forced $CMD > gpurun_out/r02b_ncu_forced.log 2>&1
# every launch of a short bench run with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --selfplay-games 4096 --mcts-trees 32 --mcts-playouts 2048 --reinforce-games 256 --reinforce-steps 1 --reinforce-1m-games 0 --valuegen-games 2048 > gpurun_out/r02b_ncu_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep
