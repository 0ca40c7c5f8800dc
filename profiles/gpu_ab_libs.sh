#!/bin/bash
# A/B of library builds on one box: bash profiles/gpu_ab_libs.sh name=path [name=path ...]   (path relative to the repo; two repetitions)
mkdir -p gpurun_out; ARGS="$*"
for rep in 1 2; do
for nv in $ARGS; do
  v=${nv%%=*}; LIB=$PWD/${nv#*=}
  for st in "20 5" "1000 50"; do
    set -- $st
    BGW_LIB=$LIB python bench.py --steps $1 --warmup $2 --no-cpu --e2e-steps 4 --observe-launches 5 2>/dev/null | python -c "
import sys, json
r = json.loads(sys.stdin.readline())
print('$v steps=$1', 'ms/step %.5f' % r['ms_per_step'], 'frac %.4f' % r['roofline']['frac'], 'given %.5f' % r['roofline']['kernel_ms_given_actions'], 'iso %.5f' % r['roofline']['kernel_ms_isolated'])"
  done
done
done | tee gpurun_out/ab_libs.txt
