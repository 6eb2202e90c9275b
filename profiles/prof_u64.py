"""ncu driver for the u64 frame kernel: a few agx_ref_fwd_dev launches at N (default 16384), 256 MiB of frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import agilex_ntt_b200 as A
from oracle import oracle as O
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
q = O.U64_PRIMES[60]
tw, pre = O.tables_u64(N, q)
frames = (256 << 20) // (N * 8)
x = O.synthetic_u64(N * 8, 7, 4 * q)
d_in = torch.from_numpy(np.tile(x, frames // 8).view(np.int64)).cuda()
d_out = torch.empty_like(d_in)
d_tw, d_pre = torch.from_numpy(tw.view(np.int64)).cuda(), torch.from_numpy(pre.view(np.int64)).cuda()
p = A.RefPipeline()
for _ in range(3):
    p.fwd_dev(N, d_in, d_in, d_out, q, d_tw, d_pre, frames)
torch.cuda.synchronize()
assert (d_out[: N].cpu().numpy().view(np.uint64) == O.ref_fwd_u64(x[:N], x[:N], q, tw, pre, 1)).all()
print("ok")
