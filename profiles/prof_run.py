"""Tiny driver for ncu captures: a few forward and inverse launches of configs[1] (n=4096, 65,536 polynomials) -- or
another size / the polymul with --n / --polymul -- through the C ABI.  Usage (GPU box, after the plain run exited 0):
    ncu --set full --clock-control none --import-source on -k regex:'ntt_|r16_|polymul' -s 4 -c 2 -o gpurun_out/prof \
        python profiles/prof_run.py"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import agilex_ntt_b200 as A

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--limbs", type=int, default=1)
ap.add_argument("--polymul", action="store_true")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
Q = (1053818881, 1054015489, 1054212097)[:a.limbs]
B = a.batch or (1 << 28) // a.n // a.limbs
ctx = A.Context(a.n, Q)
d = torch.empty(B * a.limbs * a.n, dtype=torch.int32, device="cuda")
ctx.fill_synthetic(d, seed=1234)
s0 = ctx.checksum(d)
if a.polymul:
    b = torch.empty_like(d); c = torch.empty_like(d)
    ctx.fill_synthetic(b, seed=2)
    for _ in range(a.reps):
        ctx.polymul(c, d, b)
else:
    for _ in range(a.reps):
        ctx.fwd(d)
        ctx.inv(d)
torch.cuda.synchronize()
assert ctx.checksum(d) == s0
print("ok", ctx.variant(), "launches", ctx.launch_count())
