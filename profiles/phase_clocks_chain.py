"""Per-phase clock64 timeline of the specialised kernel inside a CHAINED rollout (build with BGW_PROFILE=1; run with
BGW_PROF_FILE=... BGW_PROF_LAZY=1): python profiles/phase_clocks_chain.py <warmup steps> <rollout steps>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import bench
from abmarl_b200.engine import BatchedGridWorld
eng = BatchedGridWorld(bench.build_spec(4096, 0), device='cuda:0')
eng.reset()
W, N = int(sys.argv[1]), int(sys.argv[2])
eng.rollout_sampled(W)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); eng.rollout_sampled(N); t1.record()
torch.cuda.synchronize()
print('ms/step', t0.elapsed_time(t1) / N)
a = np.fromfile(os.environ['BGW_PROF_FILE'], dtype=np.int64).reshape(-1, 8, 16)
names = ['wait', 'zero+ctr', 'compact', 'lists', 'att-pre', 'att-rounds', 'settle/classify', 'move-phase', 'emit', 'obs', 'store/clean']
d = np.diff(a[:, :, :12], axis=2).astype(np.float64)
valid = (a[:, :, 11] > 0) & (a[:, :, 0] > 0) & (d > 0).all(axis=2) & (d < 1e6).all(axis=2)
print('valid iterations', valid.sum(), 'of', valid.size)
dv = d[valid]
for n, m, md, mx in zip(names, dv.mean(0), np.median(dv, 0), dv.max(0)):
    print(f'  {n:18s} mean {m:9.0f}  median {md:9.0f}  max {mx:9.0f}')
print('total per env: mean', dv.sum(1).mean(), 'median', np.median(dv.sum(1)))
