"""Generate tests/golden/ref_fwd_golden.npz by RUNNING THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
It builds oracle/_ref (the reference's src/kernel/ntt.cpp compiled against the host SYCL stand-in, see
oracle/ref_driver.cpp) and records inputs' seeds/parameters and the reference's outputs.  The fixtures travel to
the GPU box, where /root/reference does not exist.  Tables and inputs are regenerated from (N, q, psi, seed) by
the tests, so only outputs are stored.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

PRIMES = O.SEAL_PRIMES_30


def inputs(kind: str, N: int, q: int, frames: int, seed: int):
    g = np.arange(N * frames, dtype=np.uint64)
    if kind == "ramp":
        return g % np.uint64(q)
    if kind == "splitmix":
        return np.array([O.splitmix64(seed + int(i)) % q for i in g], dtype=np.uint64)
    if kind == "lazy4q":  # anywhere in [0,4q): accepted by the lazy butterfly (ntt.cpp:331-332)
        return np.array([O.splitmix64(seed + int(i)) % (4 * q) for i in g], dtype=np.uint64)
    if kind == "delta":
        x = np.zeros(N * frames, dtype=np.uint64)
        x[::N] = 1
        return x
    raise ValueError(kind)


CASES = [  # (name, N, q, kind, frames, seed)
    ("n32_ramp", 32, PRIMES[0], "ramp", 1, 0),
    ("n32_sm_f3", 32, PRIMES[1], "splitmix", 3, 42),
    ("n1024_delta", 1024, PRIMES[0], "delta", 1, 0),
    ("n1024_ramp", 1024, PRIMES[0], "ramp", 1, 0),
    ("n1024_sm42_q0", 1024, PRIMES[0], "splitmix", 1, 42),
    ("n1024_sm42_q1", 1024, PRIMES[1], "splitmix", 1, 42),
    ("n1024_sm42_q2", 1024, PRIMES[2], "splitmix", 1, 42),
    ("n1024_sm_f4", 1024, PRIMES[0], "splitmix", 4, 7),
    ("n1024_lazy", 1024, PRIMES[2], "lazy4q", 2, 99),
    ("n8192_sm", 8192, PRIMES[0], "splitmix", 1, 5),
]


def main():
    assert os.path.isfile("/root/reference/src/kernel/ntt.cpp"), "needs the reference tree"
    O.build(force=True)
    out = {}
    meta = []
    for name, N, q, kind, frames, seed in CASES:
        psi = O.min_psi(N, q)
        roots, precons = O.tables_u64(N, q, psi)
        x = inputs(kind, N, q, frames, seed)
        y = O.reference_fwd_u64(N, x, x, q, roots, precons, frames)
        out[name] = y.astype(np.uint32)  # outputs are < q < 2^30
        assert (y < q).all()
        meta.append(f"{name},{N},{q},{psi},{kind},{frames},{seed}")
    # main.cpp's dummy data (main.cpp:49-55): not a valid NTT instance, wraps mod 2^64 -> keep u64 + hashes
    N = 16384
    i = np.arange(N, dtype=np.uint64)
    y = O.reference_fwd_u64(N, i, i + 1, 65537, i + 2, i + 3, 1)
    out["main_dummy_u64"] = y
    txt = "".join("%d\n" % int(v) for v in y)
    meta.append("main_dummy_sha256_lines," + hashlib.sha256(txt.encode()).hexdigest())
    meta.append("main_dummy_sha256_le64," + hashlib.sha256(y.tobytes()).hexdigest())
    out["meta"] = np.array(meta)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_fwd_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
