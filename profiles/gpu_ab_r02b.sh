# like gpu_ab_r02.sh but variants are (name:ENV=VAL,...) of the CURRENT library unless libbgw_<name>.so exists
mkdir -p gpurun_out
N="--no-cpu --e2e-steps 4 --kernel-steps 1"
run() {  # tag lib steps warmup envs
  env BGW_LIB=$PWD/$2 $5 timeout -k 5 180 python bench.py $N --steps $3 --warmup $4 > gpurun_out/ab_$1.log 2>&1
  python - "$1" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([x for x in open(f'gpurun_out/ab_{n}.log') if x.startswith('{')][-1])
    print(f"{n:28s} {d['ms_per_step']:.5f} ms/step  {d['value']:.4e}  frac {d['roofline']['frac']:.4f}  {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(n, 'ERR', e, open(f'gpurun_out/ab_{n}.log').read()[-800:])
PY
}
if [ "$1" = "--no-tests" ]; then shift; else
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kat.py tests/test_gpu_managers.py -x -q > gpurun_out/ab_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/ab_pytest.log
tail -4 gpurun_out/ab_pytest.log
fi
for rep in 1 2; do
for v in "$@"; do
  name=${v%%:*}; envs=${v#*:}; [ "$envs" = "$v" ] && envs="X=1"
  lib=abmarl_b200/csrc/libbgw_$name.so; [ -f $lib ] || lib=abmarl_b200/csrc/libbgw.so
  run ${name}_early_$rep $lib 20 5 "$(echo $envs | tr ',' ' ')"
  run ${name}_full_$rep $lib 1000 50 "$(echo $envs | tr ',' ' ')"
done
done
