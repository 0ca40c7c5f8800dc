# A/B of kernel build variants / launch knobs on the headline workload (run under gpurun); prints ms per step.
# Variants are extra builds (BGW_DEFINES=... BGW_OUT=abmarl_b200/csrc/libbgw_<name>.so python -m abmarl_b200.csrc.build --force)
# selected with BGW_LIB; usage: bash profiles/gpu_ab_variants.sh name[:ENV=VAL,...] ...
mkdir -p gpurun_out
B="timeout -k 5 120 python bench.py --no-cpu --e2e-steps 4 --kernel-steps 1 --steps 1000 --warmup 50"
run() {  # name, env...
  name=$1; shift
  env "$@" $B > gpurun_out/ab_$name.log 2>&1
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([x for x in open(f'gpurun_out/ab_{n}.log') if x.startswith('{')][-1])
    print(f"{n:28s} {d['ms_per_step']:.5f} ms/step  {d['value']:.4e}  frac {d['roofline']['frac']:.4f}  {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(n, 'ERR', e, open(f'gpurun_out/ab_{n}.log').read()[-800:])
PY
}
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -x -q -k "rollout or chained or step_sampled or full_size_c5 or engine_matches_oracle" > gpurun_out/ab_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/ab_pytest.log
tail -3 gpurun_out/ab_pytest.log
for rep in 1 2; do
for v in "$@"; do
  name=${v%%:*}; envs=${v#*:}; [ "$envs" = "$v" ] && envs="X=1"
  lib=abmarl_b200/csrc/libbgw_$name.so; [ "$name" = cur ] && lib=abmarl_b200/csrc/libbgw.so
  run ${name}_$rep BGW_LIB=$PWD/$lib $(echo $envs | tr ',' ' ')
done
done
