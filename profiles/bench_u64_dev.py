"""Kernel rate of the reference-shaped u64 forward path on device-resident frames (agx_ref_fwd_dev): frames/s, fraction
of the 2*N*8 B HBM roofline and of the u64 integer roofline MEASURED on this GPU (agx_measure_butterfly_peak kind 1: the
ntt.cpp:331-369 butterfly stream alone).  60-bit NTT prime, lazy [0,4q) inputs, chunk sizes from L2-resident to 1 GiB.
Usage (GPU box): python profiles/bench_u64_dev.py > gpurun_out/bench_u64_dev.jsonl"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import agilex_ntt_b200 as A
from oracle import oracle as O      # inputs / tables / spot check only (this is a measurement script, not the product)

PEAK = 6555.2
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
q = O.U64_PRIMES[60]
p = A.RefPipeline()
ctx = A.Context(1024, [1053818881])
peak64, mhz = ctx.measure_butterfly_peak(1, 1024)
peak64_4, _ = ctx.measure_butterfly_peak(1, 512)
peak32, _ = ctx.measure_butterfly_peak(0, 1024)
sms = torch.cuda.get_device_properties(0).multi_processor_count
print(json.dumps({"u64_butterflies_per_clk_per_sm": peak64, "at_4_warps_per_scheduler": peak64_4, "u32_butterflies_per_clk_per_sm": peak32,
                  "implied_sm_mhz": mhz, "sms": sms}), flush=True)


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for N in (1024, 8192, 16384, 32768):
    logn = N.bit_length() - 1
    tw, pre = O.tables_u64(N, q)
    d_tw = torch.from_numpy(tw.view(np.int64)).cuda()
    d_pre = torch.from_numpy(pre.view(np.int64)).cuda()
    for chunk_mib in (16, 1024):
        frames = (chunk_mib << 20) // (N * 8)
        x = O.synthetic_u64(N * min(frames, 64), 7, 4 * q)
        d_in = torch.from_numpy(np.tile(x, frames // min(frames, 64)).view(np.int64)).cuda()
        d_out = torch.empty_like(d_in)
        p.fwd_dev(N, d_in, d_in, d_out, q, d_tw, d_pre, frames)
        torch.cuda.synchronize()
        got = d_out[: N * 2].cpu().numpy().view(np.uint64)
        ok = bool((got == O.ref_fwd_u64(x[: 2 * N], x[: 2 * N], q, tw, pre, 2)).all())
        l0 = p.launch_count()
        p.fwd_dev(N, d_in, d_in, d_out, q, d_tw, d_pre, frames)
        launches = p.launch_count() - l0
        ms = timed(lambda: p.fwd_dev(N, d_in, d_in, d_out, q, d_tw, d_pre, frames), 20 if chunk_mib == 16 else 5)
        fps = frames / (ms * 1e-3)
        bf = (N // 2) * logn
        print(json.dumps({"N": N, "chunk_MiB": chunk_mib, "frames": frames, "ms": ms, "frames_per_s": fps, "launches": launches,
                          "GBps_algorithmic": fps * 2 * N * 8 / 1e9, "frac_of_measured_hbm": fps * 2 * N * 8 / 1e9 / PEAK,
                          "butterflies_per_clk_per_sm": fps * bf / (sms * mhz * 1e6),
                          "frac_of_measured_u64_integer_peak": fps * bf / (sms * mhz * 1e6) / peak64, "parity": "ok" if ok else "FAILED"}),
              flush=True)
        del d_in, d_out
p.close()
ctx.close()
