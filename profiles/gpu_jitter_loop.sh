#!/bin/bash
# the jitter-build parity run several times over (an intermittent failure would be an ordering bug): logs under gpurun_out/
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
  python -m pytest tests/test_gpu_jitter.py -x -q -m gpu > gpurun_out/jitter_loop_$i.log 2>&1; echo "run $i rc=$?"
done
grep -l "failed\|Error" gpurun_out/jitter_loop_*.log
