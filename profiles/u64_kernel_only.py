import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import agilex_ntt_b200 as A
from oracle import oracle as O
q = 1053818881
N, frames = int(sys.argv[1]), int(sys.argv[2])
tw, pre = O.tables_u64(N, q)
x = (np.arange(N * frames, dtype=np.uint64) * np.uint64(2654435761)) % np.uint64(q)
t = torch.empty(x.size, dtype=torch.int64).pin_memory(); xin = t.numpy().view(np.uint64); xin[:] = x
to = torch.empty(x.size, dtype=torch.int64).pin_memory(); out = to.numpy().view(np.uint64)
p = A.RefPipeline()
for it in range(3):
    p.ntt_input_kernel(xin, xin, np.array([q], dtype=np.uint64), tw, pre, frames); p.fwd_ntt_kernel(0); p.ntt_output_kernel(out, frames); p.wait()
p.close()
