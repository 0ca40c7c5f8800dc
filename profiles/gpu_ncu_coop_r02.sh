#!/bin/bash
# ncu --set full of the fused rollout launch and of the observe kernel after the cooperative gather
bash profiles/gpu_ncu_fused.sh r02g_fused_early
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:bgw_observe_fast --launch-skip 5 --launch-count 1 -o gpurun_out/prof_r02g_observe -f python bench.py --steps 20 --warmup 5 --no-cpu --e2e-steps 4 --spinup-steps 10 > gpurun_out/r02g_observe_ncu.log 2>&1
tail -2 gpurun_out/r02g_observe_ncu.log
