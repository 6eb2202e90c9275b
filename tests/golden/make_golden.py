"""Generate tests/golden/ref_fwd_golden.npz and ref_fwd_golden_u64.npz by RUNNING THE REFERENCE'S OWN CODE.

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


# u64 cases on the moduli the reference datapath is built for (ntt.cpp:147-148, 344-363): 50-, 60-, 62- and 63-bit NTT
# primes q = 1 (mod 2^16).  Inputs lazy in [0,4q) (splitmix(seed+i) mod 4q), several frames, in2 != in (high halves come
# from in2: ntt.cpp:587-589).  N <= 8192 stores the reference's outputs, N >= 16384 their SHA-256.
# (name, N, bits, frames, seed_in, seed_in2, lazy)
U64_CASES = [
    ("u64_n1024_q50", 1024, 50, 3, 11, 12, True),
    ("u64_n1024_q60", 1024, 60, 3, 13, 14, True),
    ("u64_n1024_q62", 1024, 62, 2, 15, 16, False),   # 4q would reach 2^64: reduced inputs, lazy values inside still < 4q
    ("u64_n1024_q63_wraps", 1024, 63, 2, 17, 18, False),   # 2q < 2^64 but 4q is not: arithmetic wraps mod 2^64
    ("u64_n8192_q50", 8192, 50, 2, 21, 22, True),
    ("u64_n8192_q60", 8192, 60, 2, 23, 24, True),
    ("u64_n16384_q50", 16384, 50, 2, 31, 32, True),
    ("u64_n16384_q60", 16384, 60, 2, 33, 34, True),
    ("u64_n32768_q50", 32768, 50, 2, 41, 42, True),
    ("u64_n32768_q60", 32768, 60, 2, 43, 44, True),
]


def u64_inputs(N, q, frames, seed_in, seed_in2, lazy):
    mod = 4 * q if lazy else q
    return O.synthetic_u64(N * frames, seed_in, mod), O.synthetic_u64(N * frames, seed_in2, mod)


def make_u64():
    out, meta = {}, []
    for name, N, bits, frames, s1, s2, lazy in U64_CASES:
        q = O.U64_PRIMES[bits]
        assert O.is_prime(q) and q.bit_length() == bits and (q - 1) % (1 << 16) == 0 and q < 2**64
        psi = O.min_psi(N, q)
        assert pow(psi, N, q) == q - 1
        roots, precons = O.tables_u64(N, q, psi)
        x, x2 = u64_inputs(N, q, frames, s1, s2, lazy)
        y = O.reference_fwd_u64(N, x, x2, q, roots, precons, frames)       # the reference's own code
        if bits <= 62:
            assert (y < q).all()
        sha = hashlib.sha256(y.tobytes()).hexdigest()
        if N <= 8192:
            out[name] = y
        meta.append(f"{name},{N},{q},{psi},{frames},{s1},{s2},{int(lazy)},{sha}")
    out["meta"] = np.array(meta)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_fwd_golden_u64.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    assert os.path.isfile("/root/reference/src/kernel/ntt.cpp"), "needs the reference tree"
    O.build(force=True)
    make_u64()
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
