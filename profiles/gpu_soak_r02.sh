timeout -k 5 400 python tests/soak_c5.py 700 rollout 2>&1 | tail -2
timeout -k 5 200 python tests/chain_probe.py 50 400 1000 3000 2>&1 | grep "rollout(" | grep identical
PROBE_HORIZON=12 timeout -k 5 200 python tests/chain_probe.py 100 1000 2>&1 | grep identical
bash profiles/gpu_ab_r02b.sh --no-tests cur t96
