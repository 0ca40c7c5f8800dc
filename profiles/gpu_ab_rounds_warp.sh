#!/bin/bash
# A/B: reservation rounds on the env's last warp (now the default) against the first.  The script ran when the default still was
# the first warp; today build the variant with  BGW_FAST_DEFINES=BGW_ROUNDS_FIRST BGW_OUT=$PWD/abmarl_b200/csrc/libbgw_rl.so python -m abmarl_b200.csrc.build
BGW_LIB=$PWD/abmarl_b200/csrc/libbgw_rl.so python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tb_c5 or tb_c2 or tb_dense or rollout or chained or full_size" 2>&1 | tail -2
bash profiles/gpu_ab_libs.sh base=abmarl_b200/csrc/libbgw.so roundslast=abmarl_b200/csrc/libbgw_rl.so
