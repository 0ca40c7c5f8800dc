"""PCIe copy bandwidth on the bench box (pinned host memory): D2H / H2D alone and both at once, cudaMemcpyAsync."""
import json
import torch
n = 64 << 20
d = torch.empty(n, dtype=torch.uint8, device='cuda'); d2 = torch.empty(n, dtype=torch.uint8, device='cuda')
h = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def both():
    with torch.cuda.stream(s1):
        h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)


out = {"d2h_gbs": n / timed(lambda: h.copy_(d, non_blocking=True)) / 1e6,
       "h2d_gbs": n / timed(lambda: d2.copy_(h2, non_blocking=True)) / 1e6,
       "duplex_each_gbs": n / timed(both) / 1e6}
print(json.dumps(out))
