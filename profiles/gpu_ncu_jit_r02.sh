#!/bin/bash
# ncu --set full of the run-time compiled general kernel on pacman (BASELINE configs[2]); usage: bash profiles/gpu_ncu_jit_r02.sh [config] [tag]
C=${1:-pacman_c3}; R=${2:-r02_jit_$C}
BGW_SPECIALIZE=1 timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:bgw_step_jit --launch-skip 20 --launch-count 1 -o gpurun_out/prof_${R} -f python profiles/bench_configs.py $C > gpurun_out/${R}_ncu.log 2>&1
tail -2 gpurun_out/${R}_ncu.log
