# The interleaving-perturbation build (-DBGW_JITTER, bgw_fast.cuh) under the parity, chained / fused rollout and soak tests.
export BGW_LIB=$PWD/abmarl_b200/csrc/libbgw_jitter.so
timeout -k 5 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kat.py tests/test_gpu_managers.py -x -q 2>&1 | tail -3
timeout -k 5 600 python tests/soak_c5.py 450 rollout 2>&1 | tail -1
timeout -k 5 300 python tests/soak_c5.py 120 2>&1 | tail -1
PROBE_HORIZON=12 timeout -k 5 300 python tests/chain_probe.py 100 500 2>&1 | grep identical
