# A/B of kernel build variants / launch knobs on the headline workload (run under gpurun); prints ms per step
mkdir -p gpurun_out
B="timeout -k 5 120 python bench.py --no-cpu --e2e-steps 4 --kernel-steps 1 --steps 1000 --warmup 50"
run() {  # name, env...
  name=$1; shift
  env "$@" $B > gpurun_out/r01l_ab_$name.log 2>&1
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([x for x in open(f'gpurun_out/r01l_ab_{n}.log') if x.startswith('{')][-1])
    print(f"{n:28s} {d['ms_per_step']:.5f} ms/step  {d['value']:.4e}  frac {d['roofline']['frac']:.4f}  {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(n, 'ERR', e, open(f'gpurun_out/r01l_ab_{n}.log').read()[-800:])
PY
}
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -x -q -k "rollout or chained or step_sampled or full_size_c5" > gpurun_out/r01l_pytest_chain.log 2>&1; echo "rc=$?" >> gpurun_out/r01l_pytest_chain.log
tail -3 gpurun_out/r01l_pytest_chain.log
run prev BGW_LIB=$PWD/abmarl_b200/csrc/libbgw_prev.so
run cur X=1
run prev2 BGW_LIB=$PWD/abmarl_b200/csrc/libbgw_prev.so
run cur2 X=1
run cur_g1110 BGW_CHAIN_GRID=1110
run cur_g740 BGW_CHAIN_GRID=740
run cur_g370 BGW_CHAIN_GRID=370
run t64 BGW_LIB=$PWD/abmarl_b200/csrc/libbgw_t64.so BGW_THREADS=64
run t64_g740 BGW_LIB=$PWD/abmarl_b200/csrc/libbgw_t64.so BGW_THREADS=64 BGW_CHAIN_GRID=740
run t128 BGW_LIB=$PWD/abmarl_b200/csrc/libbgw_t128.so BGW_THREADS=128
