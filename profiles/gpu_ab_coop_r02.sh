#!/bin/bash
# A/B: cooperative row gather (default build) against one lane per learner (libbgw_onelane.so, -DBGW_OBS_ONE_LANE)
# build the variant first:  BGW_DEFINES=BGW_OBS_ONE_LANE BGW_OUT=$PWD/abmarl_b200/csrc/libbgw_onelane.so python -m abmarl_b200.csrc.build
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for rep in 1 2; do
for v in coop onelane; do
  LIB=""; [ $v = onelane ] && LIB=/root/repo/abmarl_b200/csrc/libbgw_onelane.so
  for st in "20 5" "1000 50"; do
    set -- $st
    BGW_LIB=$LIB python bench.py --steps $1 --warmup $2 --no-cpu --e2e-steps 4 2>/dev/null | python -c "
import sys, json
r = json.loads(sys.stdin.readline()); o = r['roofline']['observe_kernel']
print('$v steps=$1', 'ms/step %.5f' % r['ms_per_step'], 'frac %.4f' % r['roofline']['frac'], 'given %.5f' % r['roofline']['kernel_ms_given_actions'], 'observe ms %.5f GB/s %.0f frac %.3f' % (o['ms_per_launch'], o['achieved'], o['frac']))"
  done
done
done | tee gpurun_out/ab_coop.txt
python profiles/bench_configs.py tb_c2 2>/dev/null | tail -1 | cut -c1-330
BGW_LIB=/root/repo/abmarl_b200/csrc/libbgw_onelane.so python profiles/bench_configs.py tb_c2 2>/dev/null | tail -1 | cut -c1-330
