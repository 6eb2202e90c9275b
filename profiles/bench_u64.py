"""Throughput of the reference-shaped u64 forward pipeline (agx_ref_input / fwd / output / wait) on host buffers,
and its kernel alone, at the reference's native sizes.  Usage (GPU box): python profiles/bench_u64.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import agilex_ntt_b200 as A
from oracle import oracle as O

q = 1053818881
for N, frames in ((1024, 8192), (8192, 2048), (16384, 1024), (32768, 512)):
    tw, pre = O.tables_u64(N, q)
    x = (np.arange(N * frames, dtype=np.uint64) * np.uint64(2654435761)) % np.uint64(q)
    out = np.zeros_like(x)
    p = A.RefPipeline()
    best = 1e9
    for it in range(4):
        t0 = time.perf_counter()
        p.ntt_input_kernel(x, x, np.array([q], dtype=np.uint64), tw, pre, frames)
        p.fwd_ntt_kernel(0)
        p.ntt_output_kernel(out, frames)
        p.wait()
        best = min(best, time.perf_counter() - t0)
    ok = bool((out[:N] == O.ref_fwd_u64(x[:N], x[:N], q, tw, pre, 1)).all())
    print(json.dumps({"N": N, "frames": frames, "e2e_ms": best * 1e3, "frames_per_s_e2e": frames / best,
                      "GBps_host_traffic": 3 * x.nbytes / best / 1e9, "first_frame_ok": ok}), flush=True)
    p.close()
