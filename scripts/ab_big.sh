#!/bin/bash
# n >= 8192 (u32): ntt_big_kernel (default) against the radix-2 generic kernel (AGX_GENERIC_ONLY=1) in one call; parity first
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_random_gpu.py -m gpu -x -q 2>&1 | tail -3
for g in 0 1; do
  if [ $g = 1 ]; then export AGX_GENERIC_ONLY=1; else unset AGX_GENERIC_ONLY; fi
  timeout 300 python - <<'PY'
import os, sys, json
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
for n in (8192, 16384, 32768):
    for L in (1, 3):
        primes = [1053818881, 1054015489, 1054212097][:L] if n < 32768 else [1053818881, 1054212097, 1055260673][:L]
        c = A.Context(n, primes)
        B = (1 << 28) // n // L
        d = torch.empty(B * L * n, dtype=torch.int32, device="cuda")
        c.fill_synthetic(d, seed=3)
        s0 = c.checksum(d)
        c.fwd(d); s1 = c.checksum(d); c.inv(d)
        ok = c.checksum(d) == s0
        tf, ti = time_ms(lambda: c.fwd(d)), time_ms(lambda: c.inv(d))
        T = B * L
        print(json.dumps({"generic_only": os.environ.get("AGX_GENERIC_ONLY"), "variant": c.variant(), "n": n, "L": L, "transforms": T,
                          "fwd_ms": round(tf, 3), "inv_ms": round(ti, 3), "fwd_GBps": round(T * 8 * n / tf / 1e6, 1),
                          "inv_GBps": round(T * 8 * n / ti / 1e6, 1), "round_trip": ok, "spectrum_checksum": "%016x" % s1}), flush=True)
        c.close(); del d
PY
done
