"""oracle.py -- Python face of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import this.
The product path (agilex-ntt_b200/) never does.

Three layers, each pinning the next:
  1. ``textbook_*``: the definition, in Python big ints:
         NTT(x)[k] = sum_j x[j] * psi^((2*bitrev(k)+1)*j) mod q
     and exact schoolbook negacyclic multiplication (what NTL's ZZ_pX mul mod X^n+1 computes; README.md:9-10).
  2. ``liboracle.so`` (ntt_oracle.c): C restatement of the reference's arithmetic (ntt.cpp:146-159, 292-300,
     331-332, 344-363, 368-369, 377-393) on u64 and on the u32 datapath, plus inverse and polymul.
  3. ``oracle/_ref/libref_ntt_<N>.so``: the reference's own ntt.cpp compiled against a host SYCL stand-in
     (oracle/ref_driver.cpp) -- available only where `make -C oracle ref` was run with /root/reference present.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SEAL_PRIMES_30 = (1053818881, 1054015489, 1054212097)
REF_SIZES = (32, 1024, 8192, 16384, 32768)              # include/kernel/ntt.h:11-20
# NTT primes for the u64 datapath the reference kernel is written for (ntt.cpp:147-148, 344-363: 64-bit modulus, 64x64
# -> high-64 Shoup product): the largest primes = 1 (mod 2^16) below 2^50, 2^60, 2^62 and 2^63.  The first three keep the
# lazy range [0,4q) inside 64 bits; with the last one the arithmetic wraps mod 2^64 (the reference wraps identically).
U64_PRIMES = {50: 1125899904679937, 60: 1152921504606584833, 62: 4611686018427322369, 63: 9223372036853661697}


def build(force: bool = False) -> str:
    """Compile the C restatement (and oracle/_ref when the reference tree is present).  Building the checker
    is not using it."""
    so = os.path.join(_HERE, "_build", "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("ntt_oracle.c", "ntt_oracle.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "all"], check=True, capture_output=True)
    if os.path.isfile("/root/reference/src/kernel/ntt.cpp"):
        have = all(os.path.exists(os.path.join(_HERE, "_ref", f"libref_ntt_{n}.so")) for n in REF_SIZES)
        if force or not have:
            subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = build()
        L = C.CDLL(so)
        u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        L.orc_min_psi.restype = C.c_uint64
        L.orc_min_psi.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_is_prime.argtypes = [C.c_uint64]
        L.orc_splitmix64.restype = C.c_uint64
        L.orc_splitmix64.argtypes = [C.c_uint64]
        L.orc_tables_u64.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, u64p, u64p]
        L.orc_tables_u32.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, u32p, u32p]
        L.orc_ref_fwd_u64.argtypes = [C.c_uint32, u64p, u64p, C.c_uint64, u64p, u64p, C.c_uint32, u64p]
        L.orc_fwd_u32_barrett.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p]
        L.orc_powmod.restype = C.c_uint64
        L.orc_powmod.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_inv_u32_barrett.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p]
        L.orc_fwd_u32_shoup.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p, u32p]
        L.orc_inv_u32_shoup.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p, u32p]
        L.orc_plan_create.restype = C.c_void_p
        L.orc_plan_create.argtypes = [C.c_uint32, C.c_uint32, u32p]
        L.orc_plan_destroy.argtypes = [C.c_void_p]
        L.orc_batch_fwd_u32.argtypes = [C.c_void_p, u32p, C.c_size_t, C.c_int, C.c_int]
        L.orc_batch_inv_u32.argtypes = [C.c_void_p, u32p, C.c_size_t, C.c_int, C.c_int]
        L.orc_batch_polymul_u32.argtypes = [C.c_void_p, u32p, u32p, u32p, C.c_size_t, C.c_int]
        L.orc_batch_ref_fwd_u64.argtypes = [C.c_uint32, u64p, C.c_uint64, u64p, u64p, C.c_size_t, C.c_int]
        L.orc_polymul_schoolbook.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p, u32p]
        L.orc_fill_synthetic.argtypes = [u32p, C.c_size_t, C.c_uint32, C.c_uint32, u32p, C.c_uint64, C.c_size_t]
        L.orc_checksum_u32.restype = C.c_uint64
        L.orc_checksum_u32.argtypes = [u32p, C.c_size_t, C.c_size_t]
        L.orc_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _p32(a: np.ndarray):
    assert a.dtype == np.uint32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def _p64(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


# ------------------------------------------------------------------------------------------- layer 1: definition

def bitrev(x: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def py_min_psi(n: int, q: int) -> int:
    """Smallest primitive 2n-th root of unity mod q, by brute force over candidates (independent of the C code)."""
    assert (q - 1) % (2 * n) == 0
    for x in range(2, q):
        c = pow(x, (q - 1) // (2 * n), q)
        if pow(c, n, q) == q - 1:
            g = c
            break
    g2, cur, best = g * g % q, g, g
    for _ in range(n):
        best = min(best, cur)
        cur = cur * g2 % q
    return best


def textbook_fwd(x, q: int, psi: int):
    """O(n^2) definition: out[k] = sum_j x[j] psi^((2 br(k)+1) j) mod q (bit-reversed output order)."""
    n = len(x)
    logn = n.bit_length() - 1
    xs = [int(v) for v in x]
    out = []
    for k in range(n):
        w = pow(psi, 2 * bitrev(k, logn) + 1, q)
        acc, p = 0, 1
        for j in range(n):
            acc += xs[j] * p
            p = p * w % q
        out.append(acc % q)
    return np.array(out, dtype=np.uint64)


def textbook_inv(X, q: int, psi: int):
    """Exact inverse of textbook_fwd: x[j] = n^-1 sum_k X[k] psi^(-(2 br(k)+1) j)."""
    n = len(X)
    logn = n.bit_length() - 1
    Xs = [int(v) for v in X]
    ipsi = pow(psi, q - 2, q)
    ninv = pow(n, q - 2, q)
    ws = [pow(ipsi, 2 * bitrev(k, logn) + 1, q) for k in range(n)]
    out = []
    for j in range(n):
        acc = 0
        for k in range(n):
            acc += Xs[k] * pow(ws[k], j, q)
        out.append(acc % q * ninv % q)
    return np.array(out, dtype=np.uint64)


def textbook_polymul(a, b, q: int):
    """Exact negacyclic product mod (X^n+1, q) with Python big ints."""
    n = len(a)
    A = [int(v) for v in a]
    Bv = [int(v) for v in b]
    c = [0] * n
    for i in range(n):
        ai = A[i]
        if ai == 0:
            continue
        for j in range(n):
            k = i + j
            if k < n:
                c[k] += ai * Bv[j]
            else:
                c[k - n] -= ai * Bv[j]
    return np.array([v % q for v in c], dtype=np.uint64)


def np_polymul_exact(a, b, q: int):
    """Exact negacyclic product via numpy object-free arithmetic (u64 products reduced per row); O(n^2)."""
    n = len(a)
    a = np.asarray(a, dtype=np.uint64)
    b = np.asarray(b, dtype=np.uint64)
    c = np.zeros(n, dtype=np.uint64)
    qq = np.uint64(q)
    for i in range(n):
        row = (a[i] * b) % qq                      # a[i]*b[j] < 2^60
        c[i:] = (c[i:] + row[: n - i]) % qq
        if i:
            c[:i] = (c[:i] + qq - row[n - i:]) % qq
    return c


def sha16(arr: np.ndarray) -> str:
    """First 16 hex digits of SHA-256 over the array bytes (little-endian), as used by SURVEY.md App. A."""
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()[:16]


# ------------------------------------------------------------------------------------------- layer 2: C oracle

def min_psi(n: int, q: int) -> int:
    return int(lib().orc_min_psi(n, q))


def splitmix64(x: int) -> int:
    return int(lib().orc_splitmix64(x & 0xFFFFFFFFFFFFFFFF))


def splitmix64_np(x: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 over a uint64 array; equals orc_splitmix64 element-wise."""
    with np.errstate(over="ignore"):
        x = x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        z = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synthetic_u64(count: int, seed: int, mod: int) -> np.ndarray:
    """count words: splitmix64(seed + i) mod `mod` (mod = q for reduced data, 4q for lazy inputs; mod < 2^64)."""
    with np.errstate(over="ignore"):
        idx = np.arange(count, dtype=np.uint64) + np.uint64(seed)
    return splitmix64_np(idx) % np.uint64(mod)


def is_prime(q: int) -> bool:
    return bool(lib().orc_is_prime(q))


def tables_u32(n: int, q: int, psi: int | None = None, inverse: bool = False):
    psi = psi or min_psi(n, q)
    r = np.empty(n, dtype=np.uint32)
    p = np.empty(n, dtype=np.uint32)
    lib().orc_tables_u32(n, q, psi, int(inverse), _p32(r), _p32(p))
    return r, p


def tables_u64(n: int, q: int, psi: int | None = None, inverse: bool = False):
    psi = psi or min_psi(n, q)
    r = np.empty(n, dtype=np.uint64)
    p = np.empty(n, dtype=np.uint64)
    lib().orc_tables_u64(n, q, psi, int(inverse), _p64(r), _p64(p))
    return r, p


def ref_fwd_u64(in1, in2, modulus: int, roots, precons, num_frames: int = 1):
    """C restatement of the reference pipeline (frames: low half from in1, high half from in2)."""
    N = len(roots)
    in1 = np.ascontiguousarray(in1, dtype=np.uint64)
    in2 = np.ascontiguousarray(in2, dtype=np.uint64)
    roots = np.ascontiguousarray(roots, dtype=np.uint64)
    precons = np.ascontiguousarray(precons, dtype=np.uint64)
    out = np.empty(N * num_frames, dtype=np.uint64)
    lib().orc_ref_fwd_u64(N, _p64(in1), _p64(in2), modulus, _p64(roots), _p64(precons), num_frames, _p64(out))
    return out


def fwd_u32_with_psi(x: np.ndarray, q: int, psi: int) -> np.ndarray:
    """Forward transform of every row of x[...][n] with the caller's root psi (tables built from it the way ntt.cpp:298-300
    consumes them), Harvey/Shoup-lazy butterflies (ntt.cpp:331-393 arithmetic at u32).  Returns a new array."""
    n = x.shape[-1]
    r, p = tables_u32(n, q, psi)
    y = np.ascontiguousarray(x, dtype=np.uint32).copy()
    for row in y.reshape(-1, n):
        lib().orc_fwd_u32_shoup(n, q, _p32(r), _p32(p), row.ctypes.data_as(C.POINTER(C.c_uint32)))
    return y


def inv_u32_with_psi(x: np.ndarray, q: int, psi: int) -> np.ndarray:
    n = x.shape[-1]
    r, p = tables_u32(n, q, psi, inverse=True)
    y = np.ascontiguousarray(x, dtype=np.uint32).copy()
    for row in y.reshape(-1, n):
        lib().orc_inv_u32_shoup(n, q, _p32(r), _p32(p), row.ctypes.data_as(C.POINTER(C.c_uint32)))
    return y


def primitive_roots_2n(n: int, q: int, count: int):
    """The `count` smallest primitive 2n-th roots of unity mod q after the minimal one (odd powers of min_psi)."""
    g = min_psi(n, q)
    roots = sorted(pow(g, e, q) for e in range(1, 2 * n, 2))
    assert roots[0] == g
    return roots[1:1 + count]


class Plan:
    """[B][L][n] u32 plan: per-limb primes, minimal psi, forward/inverse Shoup tables."""

    def __init__(self, n: int, primes):
        self.n = n
        self.primes = np.array(list(primes), dtype=np.uint32)
        self.L = len(self.primes)
        self._h = lib().orc_plan_create(n, self.L, _p32(self.primes))
        if not self._h:
            raise ValueError("invalid parameters (need prime q < 2^30 with q = 1 mod 2n)")

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:      # _LIB is gone at interpreter shutdown
            _LIB.orc_plan_destroy(self._h)
            self._h = None

    def _chk(self, data):
        assert data.dtype == np.uint32 and data.flags.c_contiguous and data.size % (self.L * self.n) == 0
        return data.size // (self.L * self.n)

    def fwd(self, data: np.ndarray, variant: str = "shoup", threads: int = 1):
        B = self._chk(data)
        lib().orc_batch_fwd_u32(self._h, _p32(data), B, int(variant == "shoup"), threads)
        return data

    def inv(self, data: np.ndarray, variant: str = "shoup", threads: int = 1):
        B = self._chk(data)
        lib().orc_batch_inv_u32(self._h, _p32(data), B, int(variant == "shoup"), threads)
        return data

    def polymul(self, a: np.ndarray, b: np.ndarray, threads: int = 1):
        B = self._chk(a)
        assert a.shape == b.shape
        c = np.empty_like(a)
        lib().orc_batch_polymul_u32(self._h, _p32(c), _p32(a), _p32(b), B, threads)
        return c

    def synthetic(self, B: int, seed: int = 42, first_poly: int = 0):
        d = np.empty((B, self.L, self.n), dtype=np.uint32)
        lib().orc_fill_synthetic(_p32(d), B, self.L, self.n, _p32(self.primes), seed, first_poly)
        return d


def polymul_schoolbook(a, b, q: int):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(b, dtype=np.uint32)
    c = np.empty_like(a)
    lib().orc_polymul_schoolbook(len(a), q, _p32(a), _p32(b), _p32(c))
    return c


def checksum_u32(data: np.ndarray, first_index: int = 0) -> int:
    data = np.ascontiguousarray(data, dtype=np.uint32)
    return int(lib().orc_checksum_u32(_p32(data.reshape(-1)), data.size, first_index))


def batch_ref_fwd_u64(N, data, modulus, roots, precons, threads=1):
    assert data.dtype == np.uint64 and data.flags.c_contiguous
    lib().orc_batch_ref_fwd_u64(N, _p64(data.reshape(-1)), modulus, _p64(roots), _p64(precons),
                                data.size // N, threads)
    return data


def max_threads() -> int:
    return int(lib().orc_max_threads())


# --------------------------------------------------------------------------- layer 3: the reference's own code

def ref_available(N: int) -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", f"libref_ntt_{N}.so"))


_REFLIBS: dict[int, C.CDLL] = {}


def reference_fwd_u64(N: int, in1, in2, modulus: int, roots, precons, num_frames: int = 1):
    """Run the reference's own fwd_ntt_kernel (compiled from /root/reference/src/kernel/ntt.cpp) on the CPU."""
    if N not in _REFLIBS:
        L = C.CDLL(os.path.join(_HERE, "_ref", f"libref_ntt_{N}.so"))
        u64p = C.POINTER(C.c_uint64)
        L.ref_fwd_run.argtypes = [u64p, u64p, C.c_uint64, u64p, u64p, C.c_uint, u64p]
        L.ref_fwd_run.restype = C.c_int
        assert L.ref_ntt_size() == N
        _REFLIBS[N] = L
    in1 = np.ascontiguousarray(in1, dtype=np.uint64)
    in2 = np.ascontiguousarray(in2, dtype=np.uint64)
    roots = np.ascontiguousarray(roots, dtype=np.uint64)
    precons = np.ascontiguousarray(precons, dtype=np.uint64)
    out = np.empty(N * num_frames, dtype=np.uint64)
    rc = _REFLIBS[N].ref_fwd_run(_p64(in1), _p64(in2), modulus, _p64(roots), _p64(precons), num_frames, _p64(out))
    if rc != 0:
        raise RuntimeError("reference pipeline stalled (pipe stand-in reported a blocking read on empty pipe)")
    return out
