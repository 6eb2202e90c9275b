"""Secondary measurements beside bench.py's headline line (BASELINE.json configs 3-5): per-size forward/inverse
throughput, the 3-limb RNS batch, and the fused polymul.  Device-resident data, CUDA events, inputs > L2.
Usage (GPU box): python profiles/bench_extra.py > gpurun_out/bench_extra.jsonl"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import agilex_ntt_b200 as A

Q = (1053818881, 1054015489, 1054212097)
PEAK = 6555.2
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def line(**kw):
    print(json.dumps(kw), flush=True)


for n in (1024, 2048, 4096):
    for L in (1, 3):
        B = (1 << 28) // n // L
        ctx = A.Context(n, Q[:L])
        d = torch.empty(B * L * n, dtype=torch.int32, device="cuda")
        ctx.fill_synthetic(d, seed=1234)
        tf = timed(lambda: ctx.fwd(d))
        ti = timed(lambda: ctx.inv(d))
        byts = 2.0 * 4 * n * B * L
        line(kind="ntt", n=n, nlimbs=L, batch=B, fwd_ms=tf, inv_ms=ti, fwd_transforms_per_s=B * L / tf * 1e3,
             inv_transforms_per_s=B * L / ti * 1e3, pairs_per_s=B * L / (tf + ti) * 1e3,
             fwd_GBps=byts / tf / 1e6, inv_GBps=byts / ti / 1e6, fwd_frac_of_measured_hbm=byts / tf / 1e6 / PEAK,
             inv_frac_of_measured_hbm=byts / ti / 1e6 / PEAK)
        del d
        ctx.close()
for n, B in ((2048, 131072), (4096, 65536), (1024, 262144)):
    ctx = A.Context(n, Q[:1])
    a = torch.empty(B * n, dtype=torch.int32, device="cuda")
    b = torch.empty_like(a)
    c = torch.empty_like(a)
    ctx.fill_synthetic(a, seed=1)
    ctx.fill_synthetic(b, seed=2)
    t = timed(lambda: ctx.polymul(c, a, b), iters=10)
    byts = 3.0 * 4 * n * B
    line(kind="polymul", n=n, batch=B, ms=t, products_per_s=B / t * 1e3, GBps=byts / t / 1e6,
         frac_of_measured_hbm=byts / t / 1e6 / PEAK)
    del a, b, c
    ctx.close()
