#!/bin/bash
# forward / inverse ms at the cfg-5 sizes (1 GiB per launch) for several library builds inside ONE gpurun call:
#   bash scripts/ab_sizes.sh "" _pf1184 _pf2368      (variant "_x" = agilex-ntt_b200/lib/libagxntt_x.so)
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
  export AGX_LIB=$PWD/agilex-ntt_b200/lib/libagxntt$v.so
  timeout 300 python - <<PY
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch
import agilex_ntt_b200 as A
def time_ms(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
out = {}
for n in (1024, 2048, 4096):
    c = A.Context(n, [1053818881])
    B = (1 << 28) // n
    d = torch.empty(B * n, dtype=torch.int32, device="cuda")
    c.fill_synthetic(d, seed=3)
    s0 = c.checksum(d)
    tf = time_ms(lambda: c.fwd(d)); 
    # leave the data a valid spectrum chain: time inverse on whatever is there (values stay < q)
    ti = time_ms(lambda: c.inv(d))
    out[n] = (round(tf, 4), round(ti, 4))
    c.close(); del d
print("variant[$v] rep $rep", out, flush=True)
PY
done
done
