#!/bin/bash
# Round-2 first GPU call (run under gpurun from the repo root): CTA-pair tensor-path check + rates, trunk baseline by precision,
# pipeline trace of the trunk, compute-sanitizer on the integer kernels, ncu of the FORCED-mode (rules-only) rollout kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
./tools/micro/umma2_check > gpurun_out/r02_umma2_check.log 2>&1
python tools/bench_nets_prec.py > gpurun_out/r02_trunk_prec_base.log 2>&1
# pipeline trace (debug build), then back to the product build
IAGO_NVCC_EXTRA=-DIAGO_TRUNK_TRACE python -m iago_b200.build --force > /dev/null 2>&1 && python tools/trace_trunk.py > gpurun_out/r02_trunk_trace_base.log 2>&1
python -m iago_b200.build --force > /dev/null 2>&1
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_run.py rollout mcts selfplay > gpurun_out/r02_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r02_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --kernel-regex kns=mcts python tools/sanitize_run.py mcts > gpurun_out/r02_racecheck_mcts.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r02_racecheck_mcts.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --kernel-regex kns=rollout python tools/sanitize_run.py rollout > gpurun_out/r02_racecheck_rollout.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r02_racecheck_rollout.log
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --sections rollout"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:rollout_pair_kernel<2" -s 1 -c 1 -f -o gpurun_out/r02_rollout_forced $CMD > gpurun_out/r02_ncu_forced.log 2>&1
ls -la gpurun_out | tail -20
