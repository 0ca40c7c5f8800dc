# ncu --set full of one step launch of a side config: bash profiles/gpu_ncu_cfg.sh <config> <kernel regex> <skip> <tag>
C=$1; K=$2; SKIP=$3; R=$4
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip $SKIP --launch-count 1 -o gpurun_out/prof_${R} -f env BGW_ROLLOUT_FUSED=0 python profiles/bench_configs.py $C > gpurun_out/${R}_ncu.log 2>&1
tail -2 gpurun_out/${R}_ncu.log
