# A/B of the launch chaining (run under gpurun): new parity tests first, then the bench in four launch modes
set -x
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -x -q -k "rollout or chained or step_sampled" > gpurun_out/r01k_pytest_chain.log 2>&1; echo "rc=$?" >> gpurun_out/r01k_pytest_chain.log
tail -5 gpurun_out/r01k_pytest_chain.log
B="timeout -k 5 300 python bench.py --no-cpu --e2e-steps 4 --steps 1000 --warmup 50"
BGW_PDL=0 $B --per-step-calls > gpurun_out/r01k_ab_nopdl.log 2>&1
$B --per-step-calls > gpurun_out/r01k_ab_pdl_wait.log 2>&1
BGW_CHAIN=0 $B > gpurun_out/r01k_ab_rollout_nochain.log 2>&1
$B > gpurun_out/r01k_ab_chain.log 2>&1
for f in nopdl pdl_wait rollout_nochain chain; do echo $f; python - <<PY
import json
try:
    l=[x for x in open('gpurun_out/r01k_ab_$f.log') if x.startswith('{')][-1]; d=json.loads(l)
    print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_isolated'], d['clocks'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r01k_ab_$f.log').read()[-1500:])
PY
done
