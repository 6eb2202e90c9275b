#!/bin/bash
# u32 sizes 8192-32768: builds of ntt_big_kernel against each other in ONE gpurun call: bash scripts/ab_big_variants.sh "" _oldbig
# (variant "_x" = agilex-ntt_b200/lib/libagxntt_x.so, "" = the shipped library); parity of the shipped library first
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_random_gpu.py -m gpu -x -q -k "8192 or 16384 or 32768 or generic or random_parameters" 2>&1 | tail -2
for rep in 1 2; do
for v in "$@"; do
  AGX_LIB=$PWD/agilex-ntt_b200/lib/libagxntt$v.so timeout 200 python - "$v" <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import agilex_ntt_b200 as A
def time_ms(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
out = []
for n in (8192, 16384, 32768):
    c = A.Context(n, [1053818881])
    B = (1 << 28) // n
    d = torch.empty(B * n, dtype=torch.int32, device="cuda")
    c.fill_synthetic(d, seed=3)
    s0 = c.checksum(d)
    c.fwd(d); s1 = c.checksum(d); c.inv(d)
    ok = c.checksum(d) == s0
    out.append("n=%d fwd %.4f inv %.4f ms spectrum %016x %s" % (n, time_ms(lambda: c.fwd(d)), time_ms(lambda: c.inv(d)), s1, "ok" if ok else "FAILED"))
    c.close()
print("variant[%s] " % sys.argv[1] + " | ".join(out))
PY
done
done
