#!/bin/bash
# compute-sanitizer on the integer kernels and the search (run under gpurun from the repo root; logs land in gpurun_out/).
# tools/sanitize_run.py runs tiny problem sizes and still checks the rollout against the CPU oracle.
set -x
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_run.py rollout mcts selfplay reinforce > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 30000 --kernel-regex kns=mcts python tools/sanitize_run.py mcts > gpurun_out/sanitizer_racecheck_mcts.log 2>&1; echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck_mcts.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --kernel-regex kns=rollout python tools/sanitize_run.py rollout > gpurun_out/sanitizer_racecheck_rollout.log 2>&1; echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck_rollout.log
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_run.py rollout mcts > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck exit $?" >> gpurun_out/sanitizer_synccheck.log
python profiles/sanitizer_summary.py gpurun_out > gpurun_out/sanitizer_summary.txt
cat gpurun_out/sanitizer_summary.txt
# K6 alone (after the weight-gradient layout change): memcheck of the gradient kernels -> profiles/r02_sanitizer_memcheck_reinforce.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_run.py reinforce > gpurun_out/sanitizer_memcheck_reinforce.log 2>&1
