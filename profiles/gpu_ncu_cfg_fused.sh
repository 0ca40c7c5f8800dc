# ncu --set full of the fused rollout launch of a side config: bash profiles/gpu_ncu_cfg_fused.sh <config> <kernel regex> <tag>
C=$1; K=$2; R=$3
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 1 --launch-count 1 -o gpurun_out/prof_${R} -f python profiles/bench_configs.py $C > gpurun_out/${R}_ncu.log 2>&1
tail -2 gpurun_out/${R}_ncu.log
