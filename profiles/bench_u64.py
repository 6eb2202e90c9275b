"""Throughput of the reference-shaped u64 forward pipeline (agx_ref_input / fwd / output / wait) at the reference's
native sizes, from pageable and from page-locked host buffers (what the host mirror's sycl::buffer allocates).
Usage (GPU box): python profiles/bench_u64.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import agilex_ntt_b200 as A
from oracle import oracle as O

q = 1053818881


def pinned(a):
    t = torch.empty(a.size, dtype=torch.int64).pin_memory()
    v = t.numpy().view(np.uint64)
    v[:] = a
    return t, v


for N, frames in ((1024, 32768), (8192, 4096), (16384, 2048), (32768, 1024)):
    tw, pre = O.tables_u64(N, q)
    x = (np.arange(N * frames, dtype=np.uint64) * np.uint64(2654435761)) % np.uint64(q)
    for kind in ("pageable", "pinned"):
        if kind == "pinned":
            tx, xin = pinned(x)
            to, out = pinned(np.zeros_like(x))
        else:
            xin, out = x, np.zeros_like(x)
        p = A.RefPipeline()
        best = 1e9
        for it in range(4):
            t0 = time.perf_counter()
            p.ntt_input_kernel(xin, xin, np.array([q], dtype=np.uint64), tw, pre, frames)
            p.fwd_ntt_kernel(0)
            p.ntt_output_kernel(out, frames)
            p.wait()
            best = min(best, time.perf_counter() - t0)
        ok = bool((out[:N] == O.ref_fwd_u64(x[:N], x[:N], q, tw, pre, 1)).all()) and \
            bool((out[-N:] == O.ref_fwd_u64(x[-N:], x[-N:], q, tw, pre, 1)).all())
        print(json.dumps({"N": N, "frames": frames, "host_buffers": kind, "e2e_ms": best * 1e3,
                          "frames_per_s_e2e": frames / best, "GBps_each_direction": x.nbytes / best / 1e9,
                          "first_and_last_frame_ok": ok}), flush=True)
        p.close()
