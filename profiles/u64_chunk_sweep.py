"""End-to-end rate of the reference-shaped u64 round (agx_ref_input / fwd / output / wait, page-locked buffers) against the
pipeline's chunk size (AGX_REF_CHUNK_KB; read per round): 16 MiB (the old constant), one wave of frame CTAs
(SMs x 128 KiB = 18.5 MiB on a B200) and multiples.  python profiles/u64_chunk_sweep.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import agilex_ntt_b200 as A
from oracle import oracle as O

q = O.U64_PRIMES[60]
sms = torch.cuda.get_device_properties(0).multi_processor_count
for N in (16384, 8192, 32768):
    tw, pre = O.tables_u64(N, q)
    frames = (1 << 30) // (N * 8)
    t1 = torch.empty(N * frames, dtype=torch.int64).pin_memory(); xin = t1.numpy().view(np.uint64)
    xin[:] = (np.arange(N * frames, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) % np.uint64(q)
    t2 = torch.empty(N * frames, dtype=torch.int64).pin_memory(); out = t2.numpy().view(np.uint64)
    mod = np.array([q], dtype=np.uint64)
    for rep in range(2):
        for kb in (8192, 16384, sms * 128, 32768, 2 * sms * 128, 65536):
            os.environ["AGX_REF_CHUNK_KB"] = str(kb)
            p = A.RefPipeline()
            best = 1e9
            for it in range(5):
                t0 = time.perf_counter()
                p.ntt_input_kernel(xin, xin, mod, tw, pre, frames); p.fwd_ntt_kernel(0); p.ntt_output_kernel(out, frames); p.wait()
                best = min(best, time.perf_counter() - t0)
            ok = all(bool((out[f * N:(f + 1) * N] == O.ref_fwd_u64(xin[f * N:(f + 1) * N], xin[f * N:(f + 1) * N], q, tw, pre, 1)).all())
                     for f in (0, frames - 1))
            print(json.dumps({"N": N, "round_MiB": 1024, "chunk_KiB": kb, "chunk_frames": (kb << 10) // (N * 8), "ms": best * 1e3,
                              "frames_per_s": frames / best, "GBps_each_way": (1 << 30) / best / 1e9, "ok": ok}), flush=True)
            p.close()
