# general-kernel check after a change of bgw_step_kernel: all GPU tests, then the side configs (R = file prefix)
R=${R:-r01o}
timeout -k 5 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest.log
tail -3 gpurun_out/${R}_pytest.log
timeout -k 5 900 python profiles/bench_configs.py > gpurun_out/${R}_configs.jsonl 2> gpurun_out/${R}_configs.err
python - <<PY
import json
for l in open('gpurun_out/${R}_configs.jsonl'):
    d = json.loads(l); print(d['config'], round(d['ms_per_step'], 4), f"{d['agent_steps_per_s']:.3e}", round(d['roofline_frac'], 4))
PY
