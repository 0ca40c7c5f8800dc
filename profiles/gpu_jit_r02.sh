#!/bin/bash
# bgw_specialize (run-time compilation of the general kernel per spec): parity tests, then the side configs stock vs specialised
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "specializ" 2>&1 | tail -15
CFG="pacman_c3 maze_c1 tb_blocking tb_encoding tb_restricted tb_selective_stacked tb_ammo_selective reach_target traffic mm_c4 mm_allstep pacman_simple"
python profiles/bench_configs.py $CFG 2>gpurun_out/jit_stock.err > gpurun_out/jit_stock.jsonl
BGW_SPECIALIZE=1 BGW_JIT_CACHE=/tmp/bgw_jit python profiles/bench_configs.py $CFG 2>gpurun_out/jit_spec.err > gpurun_out/jit_spec.jsonl
python - <<'PY'
import json
a = [json.loads(l) for l in open('gpurun_out/jit_stock.jsonl')]
b = {r['config']: r for r in (json.loads(l) for l in open('gpurun_out/jit_spec.jsonl'))}
for r in a:
    s = b.get(r['config'])
    if s: print('%-22s stock %.4f ms/step %.3e   specialised %.4f ms/step %.3e  x%.2f  (compile %.1f s)' % (r['config'], r['ms_per_step'], r['agent_steps_per_s'], s['ms_per_step'], s['agent_steps_per_s'], r['ms_per_step'] / s['ms_per_step'], s['specialize_seconds']))
PY
tail -3 gpurun_out/jit_spec.err
