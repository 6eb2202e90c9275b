"""Per-CTA phase timeline of the forward kernel (AGX_TRACE build): python profiles/trace_phases.py
Needs AGX_LIB pointing at a -DAGX_TRACE=1 build of the library.  Prints mean SM-clock durations of each phase for
the sampled CTAs (every 64th) of one 65,536-polynomial launch."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import agilex_ntt_b200 as A

n, B = 4096, 65536
ctx = A.Context(n, [1053818881])
d = torch.empty(B * n, dtype=torch.int32, device="cuda")
ctx.fill_synthetic(d, seed=1)
for _ in range(3):
    ctx.fwd(d)
torch.cuda.synchronize()
L = A.lib()
L.agx_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
cnt = (B // 64) * 16
buf = np.zeros(cnt, dtype=np.uint64)
assert L.agx_debug_trace(ctx._h, buf.ctypes.data, cnt) == 0
t = buf.reshape(-1, 16).astype(np.int64)
t = t[(t > 0).all(axis=1)]
names = ["start -> col stage 0 done (incl. load wait)", "col stage 1", "col stage 2", "col stage 3", "col stage 4", "col stage 5",
         "STS cols + barrier", "LDS rows + row stage 0", "row stage 1", "row stage 2", "row stage 3", "row stage 4", "row stage 5",
         "final reduce + STS rows + barrier", "copy-out (LDS+STG)"]
dur = np.diff(t, axis=1)
print("sampled CTAs:", len(t), " mean CTA lifetime (clk):", float((t[:, 15] - t[:, 0]).mean()))
for i, nm in enumerate(names):
    print(f"  {nm:48s} mean {dur[:, i].mean():9.0f}  p10 {np.percentile(dur[:, i], 10):8.0f}  p90 {np.percentile(dur[:, i], 90):8.0f}")
