"""Where does a kernel's time go inside an H2D / kernel / D2H pipeline?  Replays the library's host pipeline (page-locked
buffers, one stream per slot, one chunk per slot; SLOTS=3 by default) in torch with CUDA events around every operation,
for the u32 kernels (n = 4096, 32 MiB chunks) and the u64 frame kernel (N = 16384, 16 MiB chunks), 1 GiB each way:
  (no kernel)  the copies alone: the box's duplex PCIe rate
  dep          the library's structure: upload, kernel, download in the slot's stream
  spin         the kernel replaced by a dependent one-CTA spin of ~60 us that touches no memory
  side         the real kernel on its own stream and buffers, no dependency on the copies
  engines      one stream per engine (uploads / kernels / downloads), chunks handed over through events
  aheadL       the slot streams again, each download issued L chunks after its upload and kernel
Result (profiles/r02_pipeline_timeline.jsonl): a kernel the copies DEPEND on adds about its own duration to the copy
engines' period whatever the structure (dep, spin, engines, ahead at 3 slots); the same kernel running independently
beside the copies costs nothing (side) -- it is the cross-engine hand-over, not memory or SM contention.
python profiles/pipeline_timeline.py   (on a GPU box; one JSON line per case; DUMP=1 prints per-chunk timelines)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import agilex_ntt_b200 as A
from oracle import oracle as O

q = 1053818881
TOTAL = 1 << 30
SLOTS = int(os.environ.get("SLOTS", "3"))


def run(kind, chunk_mib, with_kernel, mode="dep"):
    chunk = chunk_mib << 20
    nchunks = TOTAL // chunk
    host_in = torch.empty(TOTAL, dtype=torch.uint8).pin_memory()
    host_out = torch.empty(TOTAL, dtype=torch.uint8).pin_memory()
    host_in.zero_()
    d_in = [torch.empty(chunk, dtype=torch.uint8, device="cuda") for _ in range(SLOTS)]
    d_out = [torch.empty(chunk, dtype=torch.uint8, device="cuda") for _ in range(SLOTS)]
    streams = [torch.cuda.Stream() for _ in range(SLOTS)]
    if kind == "u32":
        ctx = A.Context(4096, [q])
        kern_on = lambda sl, st: ctx.fwd(d_in[sl].view(torch.int32), stream=st)
        kern = lambda sl: kern_on(sl, streams[sl])
        src_of = lambda sl: d_in[sl]
    else:
        N = 16384
        tw, pre = O.tables_u64(N, q)
        d_tw = torch.from_numpy(tw.view(np.int64)).cuda()
        d_pre = torch.from_numpy(pre.view(np.int64)).cuda()
        p = A.RefPipeline()
        frames = chunk // (N * 8)
        kern_on = lambda sl, st: p.fwd_dev(N, d_in[sl].view(torch.int64), d_in[sl].view(torch.int64), d_out[sl].view(torch.int64), q,
                                           d_tw, d_pre, frames, stream=st)
        kern = lambda sl: kern_on(sl, streams[sl])
        src_of = lambda sl: d_out[sl]
    side = torch.cuda.Stream()
    if kind == "u32":
        side_buf = torch.zeros(chunk, dtype=torch.uint8, device="cuda")
        kern_side = lambda: ctx.fwd(side_buf.view(torch.int32), stream=side)
    else:
        side_in = torch.zeros(chunk, dtype=torch.uint8, device="cuda")
        side_out = torch.zeros(chunk, dtype=torch.uint8, device="cuda")
        kern_side = lambda: p.fwd_dev(N, side_in.view(torch.int64), side_in.view(torch.int64), side_out.view(torch.int64), q,
                                      d_tw, d_pre, frames, stream=side)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    best = None
    for rep in range(3):
        E = [[ev() for _ in range(4)] for _ in range(nchunks)]      # h2d start, h2d end, kernel end, d2h end
        torch.cuda.synchronize()
        t0 = ev(); t0.record()
        for st in streams:
            st.wait_event(t0)
        if mode == "engines":
            # one stream per ENGINE instead of one per slot: uploads, kernels and downloads each in their own stream, chunks
            # handed over through events; the upload stream never contains anything that waits for a kernel
            sH, sK, sD = streams[0], streams[1], streams[2]
            for st in (sH, sK, sD):
                st.wait_event(t0)
            hd = [torch.cuda.Event() for _ in range(nchunks)]
            kd = [torch.cuda.Event() for _ in range(nchunks)]
            dd = [torch.cuda.Event() for _ in range(nchunks)]
            for i in range(nchunks):
                sl = i % SLOTS
                with torch.cuda.stream(sH):
                    if i >= SLOTS:
                        sH.wait_event(dd[i - SLOTS])
                    E[i][0].record(sH)
                    d_in[sl].copy_(host_in[i * chunk:(i + 1) * chunk], non_blocking=True)
                    E[i][1].record(sH)
                    hd[i].record(sH)
                with torch.cuda.stream(sK):
                    sK.wait_event(hd[i])
                    kern_on(sl, sK)
                    E[i][2].record(sK)
                    kd[i].record(sK)
                with torch.cuda.stream(sD):
                    sD.wait_event(kd[i])
                    host_out[i * chunk:(i + 1) * chunk].copy_(src_of(sl), non_blocking=True)
                    E[i][3].record(sD)
                    dd[i].record(sD)
        if mode.startswith("ahead"):
            # the same operations, issued in a different ORDER on the host: the next chunk's H2D (and kernel) goes in before the
            # previous chunk's D2H, so that no upload is queued behind a download that is still waiting for its kernel
            look = int(mode[5:] or 1)
            for j in range(nchunks + look):
                if j < nchunks:
                    sl, st = j % SLOTS, streams[j % SLOTS]
                    with torch.cuda.stream(st):
                        E[j][0].record(st)
                        d_in[sl].copy_(host_in[j * chunk:(j + 1) * chunk], non_blocking=True)
                        E[j][1].record(st)
                        kern(sl)
                        E[j][2].record(st)
                i = j - look
                if i >= 0:
                    sl, st = i % SLOTS, streams[i % SLOTS]
                    with torch.cuda.stream(st):
                        host_out[i * chunk:(i + 1) * chunk].copy_(src_of(sl), non_blocking=True)
                        E[i][3].record(st)
        for i in range(0 if mode.startswith("ahead") or mode == "engines" else nchunks):
            sl = i % SLOTS
            st = streams[sl]
            with torch.cuda.stream(st):
                E[i][0].record(st)
                d_in[sl].copy_(host_in[i * chunk:(i + 1) * chunk], non_blocking=True)
                E[i][1].record(st)
                if with_kernel and mode == "dep":
                    kern(sl)
                elif with_kernel and mode == "spin":          # a dependent kernel that occupies ONE SM for ~60 us and touches no memory
                    torch.cuda._sleep(int(60e-6 * 1.9e9))
                E[i][2].record(st)
                host_out[i * chunk:(i + 1) * chunk].copy_(src_of(sl) if with_kernel and mode == "dep" else d_in[sl], non_blocking=True)
                E[i][3].record(st)
            if with_kernel and mode == "side":                # the real kernel, but on its own stream and buffers: no dependency on the copies
                with torch.cuda.stream(side):
                    kern_side()
        torch.cuda.synchronize()
        T = [[t0.elapsed_time(e) for e in row] for row in E]
        total = max(r[3] for r in T)
        if best is None or total < best[0]:
            best = (total, T)
    total, T = best
    h2d = [r[1] - r[0] for r in T]
    d2h = [r[3] - r[2] for r in T]
    kms = [r[2] - r[1] for r in T]
    # H2D engine: gap between the end of copy i and the end of copy i+1 minus ... report idle = start(i+1) - end(i) as seen by events
    h2d_end = sorted(r[1] for r in T)
    d2h_end = sorted(r[3] for r in T)
    med = lambda v: float(np.median(v))
    if os.environ.get("DUMP"):
        for i in range(20, 32):
            print("   chunk %2d slot %d  h2d %.3f -> %.3f  kernel end %.3f  d2h end %.3f" % (i, i % SLOTS, T[i][0], T[i][1], T[i][2], T[i][3]))
    return {"kind": kind, "chunk_MiB": chunk_mib, "slots": SLOTS, "kernel": with_kernel and mode, "total_ms": total,
            "GBps_each_way": TOTAL / total / 1e6,
            "h2d_ms_median": med(h2d), "kernel_ms_median": med(kms), "d2h_ms_median": med(d2h),
            "h2d_end_to_end_interval_median": med(np.diff(h2d_end)), "d2h_end_to_end_interval_median": med(np.diff(d2h_end)),
            "first_h2d_end": h2d_end[0], "last_h2d_end": h2d_end[-1], "last_d2h_end": d2h_end[-1]}


for kind, mib in (("u64", 16), ("u32", 32)):
    print(json.dumps(run(kind, mib, False)), flush=True)
    for mode in (os.environ.get("MODES", "dep,spin,side,engines,ahead2").split(",")):
        print(json.dumps(run(kind, mib, True, mode)), flush=True)
