"""End-to-end rate of the host-pointer entry points from PAGEABLE caller memory (numpy arrays), against the number of
staging-copy threads per direction (AGX_HOST_COPY_THREADS; csrc/agx_copycrew.h) and against page-locked buffers.
One JSON line per case.  python profiles/e2e_pageable.py  (on a GPU box)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import agilex_ntt_b200 as A
from oracle import oracle as O

n, B, q = 4096, 65536, O.SEAL_PRIMES_30[0]
P = O.Plan(n, [q])
x = P.synthetic(B, seed=5)                       # 1 GiB, pageable
want_head = P.fwd(x[:64].copy())


def run_u32(threads):
    if threads: os.environ["AGX_HOST_COPY_THREADS"] = str(threads)
    else: os.environ.pop("AGX_HOST_COPY_THREADS", None)
    c = A.Context(n, [q])
    out = np.zeros_like(x)                       # pages touched
    c.fwd_host(x, out)                           # warm-up: staging buffers, crews
    ok = bool((out[:64] == want_head).all())
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); c.fwd_host(x, out); ts.append(time.perf_counter() - t0)
    c.inv_host(out)
    ok = ok and bool((out == x).all())
    c.close()
    t = min(ts)
    return {"case": "agx_ntt_fwd_host, n=4096, 65536 polynomials (1 GiB in, 1 GiB out), pageable numpy arrays",
            "copy_threads_per_direction": threads or "default", "ms": t * 1e3, "GB_per_s_each_way": x.nbytes / t / 1e9,
            "transforms_per_s": B / t, "ok": ok}


def run_u32_pinned():
    c = A.Context(n, [q])
    src = torch.from_numpy(x.view(np.int32)).pin_memory()
    dst = torch.empty_like(src).pin_memory()
    c.fwd_host(src, dst)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); c.fwd_host(src, dst); ts.append(time.perf_counter() - t0)
    c.close()
    t = min(ts)
    return {"case": "the same from page-locked buffers", "ms": t * 1e3, "GB_per_s_each_way": x.nbytes / t / 1e9,
            "transforms_per_s": B / t}


print(json.dumps({"host_threads": len(os.sched_getaffinity(0)), "gpu": torch.cuda.get_device_name(0)}))
for th in (1, 2, 4, 8, 0):
    print(json.dumps(run_u32(th)), flush=True)
print(json.dumps(run_u32_pinned()), flush=True)
