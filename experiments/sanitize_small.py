"""Small-case parity loop over every kernel (written for compute-sanitizer, which is closed on this pool's GPU
   boxes; run it plain):  python experiments/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import agilex_ntt_b200 as A
from oracle import oracle as O

Q = O.SEAL_PRIMES_30
for n in (4096, 2048, 1024, 256):
    ctx = A.Context(n, Q)
    P = O.Plan(n, Q)
    x = P.synthetic(3, seed=1)
    d = torch.from_numpy(x.view(np.int32)).cuda()
    ctx.fwd(d)
    assert (d.cpu().numpy().view(np.uint32).reshape(x.shape) == P.fwd(x.copy())).all()
    ctx.inv(d)
    assert (d.cpu().numpy().view(np.uint32).reshape(x.shape) == x).all()
    a, b = P.synthetic(3, seed=2), P.synthetic(3, seed=3)
    da, db = torch.from_numpy(a.view(np.int32)).cuda(), torch.from_numpy(b.view(np.int32)).cuda()
    dc = torch.empty_like(da)
    ctx.polymul(dc, da, db)
    assert (dc.cpu().numpy().view(np.uint32).reshape(a.shape) == P.polymul(a, b)).all()
    ctx.elementwise("mac", dc, da, db)
    ctx.close()
for N in (256, 1024, 16384):          # one-CTA-per-frame kernel; strided<3> + last pass; strided<4>, <3>, <3> + last pass
    p = A.RefPipeline()
    q = Q[0]
    tw, pre = O.tables_u64(N, q)
    xin = np.arange(2 * N, dtype=np.uint64)
    xin2 = xin + np.uint64(1)
    out = np.zeros(2 * N, dtype=np.uint64)
    p.ntt_input_kernel(xin, xin2, np.array([q], dtype=np.uint64), tw, pre, 2)
    p.fwd_ntt_kernel(0)
    p.ntt_output_kernel(out, 2)
    p.wait()
    assert (out == O.ref_fwd_u64(xin, xin2, q, tw, pre, 2)).all()
    p.close()
print("sanitize_small ok")
