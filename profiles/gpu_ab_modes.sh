# A/B of build variants under the three launch modes: chained rollout (default), rollout with BGW_CHAIN=0, per-step calls.
# usage: bash profiles/gpu_ab_modes.sh name ...
mkdir -p gpurun_out
N="--no-cpu --e2e-steps 4 --kernel-steps 1"
run() {  # tag lib steps warmup extra-env extra-args
  env BGW_LIB=$PWD/$2 $5 timeout -k 5 180 python bench.py $N --steps $3 --warmup $4 $6 > gpurun_out/abm_$1.log 2>&1
  python - "$1" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([x for x in open(f'gpurun_out/abm_{n}.log') if x.startswith('{')][-1])
    print(f"{n:34s} {d['ms_per_step']:.5f} ms/step  {d['value']:.4e}  frac {d['roofline']['frac']:.4f} iso {d['roofline']['kernel_ms_isolated']:.5f} {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(n, 'ERR', e, open(f'gpurun_out/abm_{n}.log').read()[-600:])
PY
}
for name in "$@"; do
  lib=abmarl_b200/csrc/libbgw_$name.so; [ "$name" = cur ] && lib=abmarl_b200/csrc/libbgw.so
  run ${name}_early_chain $lib 20 5 X=1 ""
  run ${name}_early_nochain $lib 20 5 BGW_CHAIN=0 ""
  run ${name}_early_perstep $lib 20 5 X=1 --per-step-calls
  run ${name}_full_chain $lib 1000 50 X=1 ""
  run ${name}_full_nochain $lib 1000 50 BGW_CHAIN=0 ""
done
