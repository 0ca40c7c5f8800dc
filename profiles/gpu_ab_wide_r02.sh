#!/bin/bash
# A/B of the 4-warp C5 instantiation on under-filled batches (512 and 1024 envs on one GPU), plus parity of the selection.
mkdir -p gpurun_out
for E in 256 512 1024 2048; do
  for w in 0 1; do
    for st in "20 5" "1000 50"; do
      set -- $st
      BGW_C5_WIDE=$w python bench.py --envs-per-gpu $E --steps $1 --warmup $2 --no-cpu --e2e-steps 0 --kernel-steps 0 --given-steps 0 2>/dev/null | python -c "
import sys, json
r = json.loads(sys.stdin.readline()); print('E=$E wide=$w steps=$1', 'ms/step %.5f' % r['ms_per_step'], 'value %.4g' % r['value'])"
    done
  done
done | tee gpurun_out/ab_wide.txt
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
