# ncu --set full of the FUSED rollout launch of the headline kernel (20 manager steps of 4096 envs in one launch: steady state,
# no ramp / tail per step).  usage: bash profiles/gpu_ncu_fused.sh <tag> [lib]
R=$1; LIB=${2:-abmarl_b200/csrc/libbgw.so}
N="--no-cpu --e2e-steps 4 --kernel-steps 1 --given-steps 1"
export BGW_LIB=$PWD/$LIB
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:bgw_step_fast --launch-skip 1 --launch-count 1 -o gpurun_out/prof_${R} -f python bench.py --steps 20 --warmup 5 $N > gpurun_out/${R}_ncu.log 2>&1
tail -2 gpurun_out/${R}_ncu.log
