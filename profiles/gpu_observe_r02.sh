#!/bin/bash
# bgw_observe: parity tests, the bench line with the observe_kernel record, and one ncu --set full capture of the kernel.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "observe" 2>&1 | tail -5
python -m pytest tests/test_gpu_kat.py -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --no-cpu --e2e-steps 4 > gpurun_out/observe_bench.json 2> gpurun_out/observe_bench.err || tail -5 gpurun_out/observe_bench.err
python -c "
import json; r = json.load(open('gpurun_out/observe_bench.json')); print(json.dumps(r['roofline']['observe_kernel'], indent=1)); print('frac', r['roofline']['frac'], 'ms/step', r['ms_per_step'])"
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:bgw_observe_fast --launch-skip 5 --launch-count 1 -o gpurun_out/prof_r02_observe -f python bench.py --steps 20 --warmup 5 --no-cpu --e2e-steps 4 --spinup-steps 10 > gpurun_out/r02_observe_ncu.log 2>&1
tail -2 gpurun_out/r02_observe_ncu.log
