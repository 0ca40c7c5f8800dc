"""One-off probe of the chained rollout (bgw_rollout_sampled) at the headline size: growing chain lengths, each timed and
compared with serialised launches of a second engine.  Run under gpurun with a timeout."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from abmarl_b200.engine import BatchedGridWorld

lengths = [int(x) for x in sys.argv[1:]] or [50, 100, 200, 400, 1000]
spec = bench.build_spec(4096, 0)
if os.environ.get('PROBE_HORIZON'):
    spec.horizon = int(os.environ['PROBE_HORIZON'])     # short episodes: many envs reset inside every chain
eng, ser = BatchedGridWorld(spec, device='cuda:0'), BatchedGridWorld(spec, device='cuda:0')
eng.reset()
ser.reset()
torch.cuda.synchronize()
for n in lengths:
    print(f'rollout({n}) ...', flush=True)
    t0 = time.perf_counter()
    eng.rollout_sampled(n)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    for _ in range(n):
        ser.step_sampled()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    same = all(torch.equal(getattr(eng, k), getattr(ser, k)) for k in ('obs', 'reward', 'done', 'all_done', 'actions'))
    same = same and all(torch.equal(eng.state[k], ser.state[k]) for k in ('cell', 'flags', 'health', 'step', 'episode', 'env_flags', 'stats'))
    print(f'rollout({n}): enqueue {1e3 * (t1 - t0):.1f} ms, done after {1e3 * (t2 - t0):.1f} ms ({1e3 * (t2 - t0) / n:.4f} ms/step); '
          f'serialised {1e3 * (t3 - t2) / n:.4f} ms/step; identical={same}', flush=True)
