#!/bin/bash
# bgw_specialize on sims of the specialised team-battle kernel (their own compile-time shape): parity, then stock vs specialised
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "specializ or caller_supplied" 2>&1 | tail -8
# the same with PyTorch's bundled NVRTC (12.8: no 256-bit stores, -DBGW_NO_ST256)
BGW_NVRTC=$(python -c "import nvidia.cuda_nvrtc, os; print(os.path.join(os.path.dirname(nvidia.cuda_nvrtc.__file__), 'lib', 'libnvrtc.so.12'))") python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "specialized_team_battle" 2>&1 | tail -3
CFG="tb_c5_small tb_dense"
python profiles/bench_configs.py $CFG 2>gpurun_out/jitf_stock.err > gpurun_out/jitf_stock.jsonl
BGW_SPECIALIZE=1 python profiles/bench_configs.py $CFG 2>gpurun_out/jitf_spec.err > gpurun_out/jitf_spec.jsonl
BGW_DYNAMIC_SHAPES=1 python profiles/bench_configs.py tb_c2 2>>gpurun_out/jitf_stock.err | sed 's/"tb_c2"/"tb_c2 (run-time shapes)"/' >> gpurun_out/jitf_stock.jsonl
BGW_DYNAMIC_SHAPES=1 BGW_SPECIALIZE=1 python profiles/bench_configs.py tb_c2 2>>gpurun_out/jitf_spec.err | sed 's/"tb_c2"/"tb_c2 (run-time shapes)"/' >> gpurun_out/jitf_spec.jsonl
python profiles/bench_configs.py tb_c2 2>/dev/null | sed 's/"tb_c2"/"tb_c2 (shipped shape)"/' >> gpurun_out/jitf_stock.jsonl
python - <<'PY'
import json
a = [json.loads(l) for l in open('gpurun_out/jitf_stock.jsonl')]
b = {r['config']: r for r in (json.loads(l) for l in open('gpurun_out/jitf_spec.jsonl'))}
for r in a:
    s = b.get(r['config'])
    if s: print('%-26s stock %.4f ms/step %.3e   specialised %.4f ms/step %.3e  x%.2f  (compile %.1f s)' % (r['config'], r['ms_per_step'], r['agent_steps_per_s'], s['ms_per_step'], s['agent_steps_per_s'], r['ms_per_step'] / s['ms_per_step'], s['specialize_seconds']))
    else: print('%-26s stock %.4f ms/step %.3e' % (r['config'], r['ms_per_step'], r['agent_steps_per_s']))
PY
tail -3 gpurun_out/jitf_spec.err
