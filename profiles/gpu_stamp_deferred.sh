#!/bin/bash
# deferred env stamps: chaining / rollout / jitter tests, soak + chain probes, then A/B against immediate stamps.  The script ran with
# deferred stamps as the default build and -DBGW_STAMP_IMMEDIATE as libbgw_si.so; today the default is immediate and the variant is
#   BGW_FAST_DEFINES=BGW_STAMP_DEFERRED BGW_OUT=$PWD/abmarl_b200/csrc/libbgw_sd.so python -m abmarl_b200.csrc.build
python -m pytest tests/test_gpu_parity.py tests/test_gpu_jitter.py -x -q -m gpu -k "rollout or chained or full_size or cuda_graph or step_sampled or jitter or specialized_team or tb_c2 or tb_c5" 2>&1 | tail -3
timeout -k 5 400 python tests/soak_c5.py 700 rollout 2>&1 | tail -2
timeout -k 5 200 python tests/chain_probe.py 50 400 1000 3000 2>&1 | grep "rollout(" | grep -c identical
PROBE_HORIZON=12 timeout -k 5 200 python tests/chain_probe.py 100 1000 2>&1 | grep -c identical
bash profiles/gpu_ab_libs.sh deferred=abmarl_b200/csrc/libbgw.so immediate=abmarl_b200/csrc/libbgw_si.so
for l in "" abmarl_b200/csrc/libbgw_si.so; do BGW_LIB=${l:+$PWD/$l} python profiles/bench_configs.py tb_c2 2>/dev/null | python -c "
import sys, json; r = json.loads(sys.stdin.readline()); print('tb_c2 ${l:-deferred}', 'ms/step %.5f' % r['ms_per_step'], '%.4g' % r['agent_steps_per_s'], 'frac %.4f' % r['roofline_frac'])"; done
