"""Per-CTA summary of the persistent forward kernel (AGX_TRACE build): python profiles/trace_pers.py
Needs AGX_LIB pointing at a -DAGX_TRACE=1 build.  Per resident CTA: SM id, start/end (globaltimer ns), polynomials
done, total clocks, clocks waiting for the staged polynomial, clocks in the output phase."""
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
cnt = 8 * 8192
buf = np.zeros(cnt, dtype=np.uint64)
assert L.agx_debug_trace(ctx._h, buf.ctypes.data, cnt) == 0
t = buf.reshape(-1, 8).astype(np.int64)
t = t[t[:, 3] > 0]
print("CTAs:", len(t), "SMs:", len(set(t[:, 0])))
g0 = t[:, 1].min()
start, end = t[:, 1] - g0, t[:, 2] - g0
print("start ns: min %d max %d | end ns: min %d p10 %d p50 %d p90 %d max %d" % (start.min(), start.max(), end.min(),
      np.percentile(end, 10), np.percentile(end, 50), np.percentile(end, 90), end.max()))
per = t[:, 4] / t[:, 3]
print("clk per polynomial: mean %.0f min %.0f p10 %.0f p90 %.0f max %.0f" % (per.mean(), per.min(), np.percentile(per, 10), np.percentile(per, 90), per.max()))
print("wait clk per polynomial: mean %.0f p90 %.0f max %.0f" % ((t[:, 5] / t[:, 3]).mean(), np.percentile(t[:, 5] / t[:, 3], 90), (t[:, 5] / t[:, 3]).max()))
print("output clk per polynomial: mean %.0f p90 %.0f" % ((t[:, 6] / t[:, 3]).mean(), np.percentile(t[:, 6] / t[:, 3], 90)))
# per-SM: spread of CTA end times inside an SM and between SMs
sm_end = {}
for r in t:
    sm_end.setdefault(int(r[0]), []).append(int(r[2] - g0))
ends = np.array([max(v) for v in sm_end.values()])
print("per-SM last end ns: min %d p50 %d max %d ; CTAs per SM: min %d max %d" % (ends.min(), np.percentile(ends, 50), ends.max(),
      min(len(v) for v in sm_end.values()), max(len(v) for v in sm_end.values())))
inner = np.array([max(v) - min(v) for v in sm_end.values()])
print("within-SM end spread ns: mean %d max %d" % (inner.mean(), inner.max()))
