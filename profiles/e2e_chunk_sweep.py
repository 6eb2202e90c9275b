import os, sys, time, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, os.getcwd())
import torch
import agilex_ntt_b200 as A
n, B = 4096, 65536
ctx = A.Context(n, [1053818881])
host = torch.empty(B * n, dtype=torch.int32).pin_memory()
d = torch.empty(B * n, dtype=torch.int32, device='cuda'); ctx.fill_synthetic(d, seed=1); host.copy_(d)
ctx.fwd_host(host); ctx.inv_host(host)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    ctx.fwd_host(host); ctx.inv_host(host)
dt = time.perf_counter() - t0
print(os.environ.get('AGX_HOST_CHUNK_MB', 'default'), 'MB chunks: pairs/s %.3fM, per-direction GB/s %.1f' % (B * 5 / dt / 1e6, 2 * B * n * 4 * 5 / dt / 1e9), bool((host.cuda() == d).all()))
