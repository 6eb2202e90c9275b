#!/bin/bash
# A/B of the polynomial product at cfg 4 (n = 2048, 131 072 products) and n = 4096 inside ONE gpurun call:
# two launches (shipped) against AGX_POLYMUL_SPLIT=1 (round 1's three launches).  Parity first.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "polymul or config4 or tensor_store" 2>&1 | tail -3
for rep in 1 2; do
for split in 0 1; do
  if [ $split = 1 ]; then export AGX_POLYMUL_SPLIT=1; else unset AGX_POLYMUL_SPLIT; fi
  timeout 300 python - <<'PY'
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch
import agilex_ntt_b200 as A
PRIME = 1053818881
def time_ms(fn, iters):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for nn, BB, L in ((2048, 131072, 1), (4096, 65536, 1), (2048, 43691, 3)):
    primes = [PRIME] if L == 1 else [1053818881, 1054015489, 1054212097]
    c = A.Context(nn, primes)
    a, b = (torch.empty(BB * L * nn, dtype=torch.int32, device="cuda") for _ in range(2))
    out = torch.empty_like(a)
    c.fill_synthetic(a, seed=1); c.fill_synthetic(b, seed=2)
    l0 = c.launch_count(); c.polymul(out, a, b); per = c.launch_count() - l0
    s = c.checksum(out)
    t = time_ms(lambda: c.polymul(out, a, b), 20)
    print(json.dumps({"split_env": os.environ.get("AGX_POLYMUL_SPLIT"), "n": nn, "L": L, "B": BB, "launches": per, "ms": round(t, 4),
                      "M_products_per_s": round(BB * L / t / 1e3, 2), "checksum": "%016x" % s}), flush=True)
    c.close(); del a, b, out
PY
done
done
