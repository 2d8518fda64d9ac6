#!/bin/bash
# A/B of kernel builds: tools/ab_rollout.sh lib1.so lib2.so ...  (run on the GPU box; prints plies/s per library)
for lib in "$@"; do
  IAGO_B200_LIB=$PWD/$lib python bench.py --sections rollout --no-cpu --steps 30 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('$lib', '%.4g plies/s' % l['value'], '%.4f ms' % l['ms_per_step'], 'frac %.3f' % l['roofline']['frac'], 'e2e %.4g' % l['e2e']['value'], 'rules-only %.4g plies/s' % l['roofline_movegen']['plies_per_s'])"
done
