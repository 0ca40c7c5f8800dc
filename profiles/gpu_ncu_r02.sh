# ncu --set full of one early (launch 20 after reset) and one late (launch 180) step launch of the headline kernel.
# usage: bash profiles/gpu_ncu_r02.sh <tag> [lib]    -> gpurun_out/prof_<tag>_{early,late}.ncu-rep
R=$1; LIB=${2:-abmarl_b200/csrc/libbgw.so}
N="--no-cpu --e2e-steps 4 --kernel-steps 1"
export BGW_LIB=$PWD/$LIB
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:bgw_step_fast --launch-skip 15 --launch-count 1 -o gpurun_out/prof_${R}_early -f python bench.py --steps 20 --warmup 5 $N > gpurun_out/${R}_ncu_early.log 2>&1
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:bgw_step_fast --launch-skip 180 --launch-count 1 -o gpurun_out/prof_${R}_late -f python bench.py --steps 200 --warmup 20 $N > gpurun_out/${R}_ncu_late.log 2>&1
tail -2 gpurun_out/${R}_ncu_early.log
