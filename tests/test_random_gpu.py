"""Seeded random differential test: the CUDA path through the C ABI against the CPU oracle on parameter combinations
no hand-written case lists -- random size (tuned and generic kernels), limb count, batch, primes, per-limb roots
(minimal or not), lazy inputs -- forward, inverse, round trip and product.  Bit-exact, tolerance 0."""
import os

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

# 30-bit NTT primes of SURVEY.md App. A with the 2-adicity of q - 1 (all < 2^30, so [0,4q) fits a u32)
PRIMES = {1053818881: 20, 1054015489: 16, 1054212097: 17, 1055260673: 17, 1056178177: 18, 1056440321: 19,
          1058209793: 16, 1060175873: 16, 1060700161: 16, 1060765697: 17, 1061093377: 16, 1062469633: 18, 1062535169: 16,
          134012929: 13, 134111233: 13, 134176769: 13}


@pytest.fixture(scope="module")
def A():
    import agilex_ntt_b200 as pkg
    return pkg


@pytest.fixture(scope="module")
def torch():
    import torch as t
    assert t.cuda.is_available()
    return t


def _cases(count, seed):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(count):
        logn = int(rng.choice([3, 4, 5, 6, 7, 8, 9, 10, 10, 11, 11, 12, 12, 12, 13, 14, 15]))
        ok = [q for q, adic in PRIMES.items() if adic >= logn + 1]
        L = int(rng.integers(1, min(5, len(ok)) + 1))
        primes = [int(q) for q in rng.choice(ok, size=L, replace=False)]
        B = int(rng.integers(1, 41 if logn <= 12 else 7))
        out.append(pytest.param(logn, primes, B, int(rng.integers(0, 4)), int(rng.integers(1, 1 << 30)),
                                id=f"{i}-n{1 << logn}-L{L}-B{B}"))
    return out


@pytest.mark.parametrize("logn,primes,B,root_pick,seed", _cases(int(os.environ.get("AGX_RANDOM_CASES", "48")), int(os.environ.get("AGX_RANDOM_SEED", "20261018"))))
def test_random_parameters_against_the_oracle(A, torch, logn, primes, B, root_pick, seed):
    n, L = 1 << logn, len(primes)
    for q in primes:
        assert O.is_prime(q) and (q - 1) % (2 * n) == 0
    # root_pick 0: the library's minimal roots; otherwise the caller's (root_pick + l)-th smallest primitive 2n-th root per limb
    if root_pick == 0 or n > 8192:
        psis, c = [O.min_psi(n, q) for q in primes], A.Context(n, primes)
    else:
        psis = []
        for l, q in enumerate(primes):                      # tiny n has only n - 1 roots besides the minimal one
            r = O.primitive_roots_2n(n, q, root_pick + l + 1)
            psis.append(r[(root_pick + l) % len(r)])
        c = A.Context(n, primes, psi=psis)
    rng = np.random.default_rng(seed)
    qs = np.array(primes, dtype=np.uint64).reshape(1, L, 1)
    x = (rng.integers(0, 1 << 62, size=(B, L, n), dtype=np.uint64) % qs).astype(np.uint32)
    dev = lambda v: torch.from_numpy(np.ascontiguousarray(v).view(np.int32)).cuda()
    back = lambda t: t.cpu().numpy().view(np.uint32).reshape(B, L, n)

    want = np.stack([O.fwd_u32_with_psi(x[:, l, :], q, psis[l]) for l, q in enumerate(primes)], axis=1)
    d = dev(x)
    c.fwd(d)
    y = back(d)
    assert (y == want).all(), "forward"
    assert (y < qs).all(), "forward outputs must be fully reduced"
    c.inv(d)
    assert (back(d) == x).all(), "round trip"
    # inverse of an arbitrary (not forward-produced) reduced vector, against the oracle's inverse
    z = (rng.integers(0, 1 << 62, size=(B, L, n), dtype=np.uint64) % qs).astype(np.uint32)
    want_inv = np.stack([O.inv_u32_with_psi(z[:, l, :], q, psis[l]) for l, q in enumerate(primes)], axis=1)
    dz = dev(z)
    c.inv(dz)
    assert (back(dz) == want_inv).all(), "inverse"
    # lazy inputs in [q, 2q): same spectrum as their residues (ntt.cpp:331-332 corrects on the fly)
    lazy = (x.astype(np.uint64) + qs).astype(np.uint32)
    dl = dev(lazy)
    c.fwd(dl)
    assert (back(dl) == want).all(), "forward on lazy inputs"
    # the negacyclic product does not depend on the choice of psi: compare with the minimal-root oracle's product
    if n <= 4096:
        P = O.Plan(n, primes)
        dc = torch.empty_like(d)
        c.polymul(dc, dev(x), dev(z))
        prod = P.polymul(x.copy(), z.copy())
        assert (back(dc) == prod).all(), "product"
        zh = dev(z)                                        # the same product with z kept in evaluation form, in place on x
        c.fwd(zh)
        dx = dev(x)
        c.polymul_by_spectrum(dx, dx, zh)
        assert (back(dx) == prod).all(), "product by spectrum"
    c.close()


def _u64_cases(count, seed):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(count):
        logn = int(rng.choice([2, 3, 5, 8, 9, 10, 10, 11, 12, 13, 13, 14, 14, 15]))
        frames = int(rng.integers(1, 7))
        kind = int(rng.integers(0, 3))       # 0: NTT prime + real tables, 1: arbitrary modulus + arbitrary tables, 2: tiny modulus
        out.append(pytest.param(logn, frames, kind, int(rng.integers(1, 1 << 30)), id=f"{i}-N{1 << logn}-f{frames}-k{kind}"))
    return out


@pytest.mark.parametrize("logn,frames,kind,seed", _u64_cases(int(os.environ.get("AGX_RANDOM_CASES_U64", "40")), int(os.environ.get("AGX_RANDOM_SEED", "20261018")) + 1))
def test_random_u64_frames_against_the_restatement(A, torch, logn, frames, kind, seed):
    """The reference-shaped u64 path claims the reference's arithmetic bit for bit for ANY modulus and ANY tables
    (ntt.cpp:147-148, 331-369 wrap mod 2^64; main.cpp:46-55 itself feeds tables that are not Shoup pairs): random sizes,
    frame counts, in2 != in, separate and in-place output, on (0) NTT primes with real tables and lazy inputs, (1) arbitrary
    64-bit moduli with arbitrary 64-bit table entries and inputs, (2) a tiny modulus."""
    N = 1 << logn
    rng = np.random.default_rng(seed)
    full = lambda size: rng.integers(0, 1 << 63, size=size, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=size, dtype=np.uint64)
    if kind == 0:
        q = O.U64_PRIMES[int(rng.choice([50, 60, 62, 63]))]
        tw, pre = O.tables_u64(N, q)
        x, x2 = O.synthetic_u64(N * frames, seed, 4 * q if q < (1 << 62) else q), O.synthetic_u64(N * frames, seed + 1, q)
    elif kind == 1:
        q = int(full(1)[0]) | 1
        tw, pre, x, x2 = full(N), full(N), full(N * frames), full(N * frames)
    else:
        q = 65537
        tw, pre = full(N) % np.uint64(q), full(N)
        x, x2 = full(N * frames) % np.uint64(4 * q), full(N * frames) % np.uint64(4 * q)
    want = O.ref_fwd_u64(x, x2, q, tw, pre, frames)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
    d_in, d_in2, d_tw, d_pre = dev(x), dev(x2), dev(tw), dev(pre)
    d_out = torch.zeros_like(d_in)
    p = A.RefPipeline()
    p.fwd_dev(N, d_in, d_in2, d_out, q, d_tw, d_pre, frames)
    torch.cuda.synchronize()
    assert (d_out.cpu().numpy().view(np.uint64) == want).all(), "separate output"
    want_same = O.ref_fwd_u64(x, x, q, tw, pre, frames)
    p.fwd_dev(N, d_in, d_in, d_in, q, d_tw, d_pre, frames)
    torch.cuda.synchronize()
    assert (d_in.cpu().numpy().view(np.uint64) == want_same).all(), "in place"
    # and through the three reference-named calls on host buffers (main.cpp:60-74)
    out = np.zeros(N * frames, dtype=np.uint64)
    p.ntt_input_kernel(x, x2, np.array([q], dtype=np.uint64), tw, pre, frames)
    p.fwd_ntt_kernel(0)
    p.ntt_output_kernel(out, frames)
    p.wait()
    assert (out == want).all(), "host pipeline"
    p.close()


def _host_cases(count, seed):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(count):
        logn = int(rng.choice([10, 11, 12, 12, 13]))
        L = int(rng.integers(1, 4))
        row_bytes = (4 << logn) * L
        B = int(rng.integers(1, max(2, (110 << 20) // row_bytes)))       # up to ~110 MiB: one to three pipeline chunks, ragged
        out.append(pytest.param(logn, L, B, bool(rng.integers(0, 2)), bool(rng.integers(0, 2)), bool(rng.integers(0, 2)),
                                int(rng.integers(1, 1 << 30)), id=f"h{i}-n{1 << logn}-L{L}-B{B}"))
    return out


@pytest.mark.parametrize("logn,L,B,pin_in,pin_out,in_place,seed",
                         _host_cases(int(os.environ.get("AGX_RANDOM_CASES_HOST", "10")), int(os.environ.get("AGX_RANDOM_SEED", "20261018")) + 2))
def test_random_host_pipelines(A, torch, logn, L, B, pin_in, pin_out, in_place, seed):
    """Host-pointer entry points on random shapes: pageable and page-locked buffers in every combination (the pageable side
    goes through the staging copies of csrc/agx_copycrew.h), in place or not, batches that end inside a pipeline chunk --
    forward against the oracle, inverse back to the input, and the two products on a prefix."""
    n = 1 << logn
    primes = list(PRIMES)[:L]
    P = O.Plan(n, primes)
    c = A.Context(n, primes)
    x = P.synthetic(B, seed=seed)
    want = P.fwd(x.copy(), threads=O.max_threads())

    def buf(a, pinned):
        if not pinned:
            return a.copy(), (lambda b: b)
        t = torch.from_numpy(a.view(np.int32).copy()).pin_memory()
        return t, (lambda b: b.numpy().view(np.uint32).reshape(a.shape))

    src, view_src = buf(x, pin_in)
    if in_place:
        dst, view_dst = src, view_src
    else:
        dst, view_dst = buf(np.zeros_like(x), pin_out)
    c.fwd_host(src, dst)
    assert (view_dst(dst) == want).all(), "fwd_host"
    c.inv_host(dst)
    assert (view_dst(dst) == x).all(), "inv_host"
    m = min(B, 64)                                                       # products on a prefix (the CPU product is the slow side)
    a, b = x[:m].copy(), P.synthetic(m, seed=seed + 1)
    prod = P.polymul(a.copy(), b.copy(), threads=O.max_threads())
    oa, va = buf(a, pin_in)
    ob, vb = buf(b, pin_out)
    oc, vc = buf(np.zeros_like(a), pin_out)
    c.polymul_host(oc, oa, ob)
    assert (vc(oc) == prod).all(), "polymul_host"
    bh = P.fwd(b.copy())
    obh, _ = buf(bh, pin_in)
    c.polymul_by_spectrum_host(oa, oa, obh)                              # in place on a
    assert (va(oa) == prod).all(), "polymul_by_spectrum_host"
    c.close()
