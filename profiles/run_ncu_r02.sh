#!/bin/bash
# Round-2 ncu captures (run under gpurun from the repo root; reports land in gpurun_out/, summaries are made here with profiles/ncu_summary.py).
set -x
mkdir -p gpurun_out
# the trunk in the inference precision (2) on CTA pairs, and in precision 3
PRECS=2 ncu --set full --clock-control none --import-source on -k regex:trunk_kernel -s 4 -c 1 -f -o gpurun_out/r02_trunk_p2 python tools/bench_nets_prec.py > gpurun_out/r02_ncu_trunk_p2.log 2>&1
PRECS=3 ncu --set full --clock-control none --import-source on -k regex:trunk_kernel -s 4 -c 1 -f -o gpurun_out/r02_trunk_p3 python tools/bench_nets_prec.py > gpurun_out/r02_ncu_trunk_p3.log 2>&1
# the rollout kernel: the bench step (Philox) and the rules-only FORCED replay
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --sections rollout"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:rollout_pair_kernel<\(int\)0, \(bool\)0" -s 3 -c 1 -f -o gpurun_out/r02_rollout_philox $CMD > gpurun_out/r02_ncu_rollout.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:rollout_pair_kernel<\(int\)2" -s 1 -c 1 -f -o gpurun_out/r02_rollout_forced $CMD > gpurun_out/r02_ncu_forced.log 2>&1
# every launch of a short bench run with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --selfplay-games 4096 --mcts-trees 32 --mcts-playouts 2048 --reinforce-games 256 --reinforce-steps 1 --reinforce-1m-games 0 --valuegen-games 2048 > gpurun_out/r02_ncu_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep
