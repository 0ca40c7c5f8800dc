import sys, os
sys.path.insert(0, '/root/repo')
import torch, numpy as np
import bench
from abmarl_b200.engine import BatchedGridWorld
eng = BatchedGridWorld(bench.build_spec(4096, 0), device='cuda:0')
eng.reset()
N = int(sys.argv[1])
for t in range(N):
    eng.step(eng.sample_actions())
torch.cuda.synchronize()
a = np.fromfile(os.environ['BGW_PROF_FILE'], dtype=np.int64).reshape(-1, 8, 16)
names = ['wait', 'zero+ctr', 'compact', 'lists', 'att-pre', 'att-rounds', 'settle/classify', 'move-rounds', 'emit', 'obs', 'store/clean']
d = np.diff(a[:, :, :12], axis=2).astype(np.float64)
valid = (a[:, :, 11] > 0) & (a[:, :, 0] > 0)
print('valid iterations', valid.sum(), 'of', valid.size)
dv = d[valid]
print('mean cycles per phase:')
for n, m, mx in zip(names, dv.mean(0), dv.max(0)):
    print(f'  {n:18s} {m:9.0f}  max {mx:9.0f}')
print('total per env', dv.sum(1).mean())
tot = (a[:, :, 11].max(1) - a[:, 0, 0])
print('per-CTA span: mean', tot.mean(), 'max', tot.max())
