"""GPU parity: the CUDA path, called through the C ABI (include/agxntt.h via agilex_ntt_b200.binding), against the
CPU oracle on identical seeded inputs, the golden outputs recorded from the reference's own code, SURVEY.md App. A
known answers, and -- at BASELINE.json's full sizes -- full compares plus size-independent properties.
Bit-exact everywhere (integer work): tolerance = 0."""
import hashlib

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

Q = O.SEAL_PRIMES_30


@pytest.fixture(scope="module")
def A():
    import agilex_ntt_b200 as pkg
    return pkg


@pytest.fixture(scope="module")
def torch():
    import torch as t
    assert t.cuda.is_available()
    return t


def to_dev(torch, a: np.ndarray):
    return torch.from_numpy(a.view(np.int32)).cuda()


def to_np(t) -> np.ndarray:
    return t.cpu().numpy().view(np.uint32)


_ctx_cache = {}


def ctx_for(A, n, primes):
    key = (n, tuple(primes))
    if key not in _ctx_cache:
        _ctx_cache[key] = A.Context(n, primes)
    return _ctx_cache[key]


# ------------------------------------------------------------------------------------------------ tables

@pytest.mark.parametrize("n", [8, 32, 1024, 2048, 4096, 16384, 32768])
def test_tables_match_oracle(A, n):
    c = ctx_for(A, n, Q)
    for l, q in enumerate(Q):
        assert c.psi(l) == O.min_psi(n, q)
        for inverse in (False, True):
            r, p = c.tables(l, inverse)
            ro, po = O.tables_u32(n, q, inverse=inverse)
            assert (r == ro).all() and (p == po).all()


def test_variants(A):
    assert ctx_for(A, 4096, Q).variant() == "ntt2p<12,6>"
    assert ctx_for(A, 2048, Q).variant() == "ntt2p<11,6>"
    assert ctx_for(A, 1024, Q).variant() == "ntt2p<10,5>"
    assert ctx_for(A, 32, Q).variant() == "generic"


def test_bad_parameters(A):
    with pytest.raises(A.AgxError):
        A.Context(4096, [1053818883])          # not prime
    with pytest.raises(A.AgxError):
        A.Context(4096, [12289])               # prime, but 12289 != 1 mod 8192
    with pytest.raises(A.AgxError):
        A.Context(4000, [Q[0]])                # n not a power of two
    with pytest.raises(A.AgxError):
        A.Context(4096, [])                    # no limbs


# ------------------------------------------------------------------------------- caller-supplied roots / tables

# (q, psi) pairs SURVEY.md App. A recalls from SEAL-Embedded for n = 4096; each equals the minimal root of its prime
SEAL_PSI_4096 = {134012929: 7470, 134111233: 3856, 134176769: 24149, 1053818881: 503422, 1054015489: 16768, 1054212097: 7305}


@pytest.mark.parametrize("n", [64, 1024, 2048, 4096, 8192])
def test_caller_psi_non_minimal(A, torch, n):
    """agx_create_tables: the reference's contract is that the caller supplies the roots (include/kernel/ntt.h:38-39,
    main.cpp:46-55); a scheme that fixes a non-minimal psi gets exactly that psi's transform (oracle with the same psi,
    and the textbook definition at small n), forward and inverse, per limb."""
    primes = Q if n <= 4096 else Q[:2]
    psis = [O.primitive_roots_2n(n, q, 3)[l] for l, q in enumerate(primes)]      # a different non-minimal root per limb
    c = A.Context(n, primes, psi=psis)
    P = O.Plan(n, primes)
    x = P.synthetic(9, seed=5)
    d = to_dev(torch, x)
    c.fwd(d)
    y = to_np(d).reshape(x.shape)
    for l, q in enumerate(primes):
        assert c.psi(l) == psis[l] != O.min_psi(n, q)
        want = O.fwd_u32_with_psi(x[:, l, :], q, psis[l])
        assert (y[:, l, :] == want).all()
        assert not (y[:, l, :] == P.fwd(x.copy())[:, l, :]).all()                  # and it is NOT the minimal-root transform
        r, p = c.tables(l)
        ro, po = O.tables_u32(n, q, psis[l])
        assert (r == ro).all() and (p == po).all()
        if n == 64:
            assert (O.textbook_fwd(x[0, l], q, psis[l]) == y[0, l]).all()
    c.inv(d)
    assert (to_np(d).reshape(x.shape) == x).all()
    d2 = to_dev(torch, y)
    c.inv(d2)
    for l, q in enumerate(primes):
        assert (to_np(d2).reshape(x.shape)[:, l, :] == O.inv_u32_with_psi(y[:, l, :], q, psis[l])).all()
    if n <= 4096:                                                                  # the product does not depend on psi
        a, b = P.synthetic(3, seed=1), P.synthetic(3, seed=2)
        dc = torch.empty_like(to_dev(torch, a))
        c.polymul(dc, to_dev(torch, a), to_dev(torch, b))
        assert (to_np(dc).reshape(a.shape) == P.polymul(a, b)).all()
    c.close()


def test_caller_psi_seal_pairs_and_rejects(A, torch):
    n = 4096
    for q, psi in SEAL_PSI_4096.items():
        c = A.Context(n, [q], psi=[psi])
        assert c.psi(0) == psi == O.min_psi(n, q)
        x = O.Plan(n, [q]).synthetic(2, seed=9)
        d = to_dev(torch, x)
        c.fwd(d)
        assert (to_np(d).reshape(x.shape) == O.Plan(n, [q]).fwd(x.copy())).all()
        c.close()
    q = Q[0]
    for bad in (0, 1, q - 1, q, 2, pow(O.min_psi(n, q), 2, q)):      # zero, order 1 / 2, out of range, not a 2n-th root, order n
        with pytest.raises(A.AgxError):
            A.Context(n, [q], psi=[bad])
    with pytest.raises(ValueError):
        A.Context(n, Q, psi=[1, 2])                                  # one psi per limb


@pytest.mark.parametrize("n", [256, 1024, 2048, 4096])
def test_caller_tables_in_reference_order(A, torch, n):
    """agx_set_tables: roots / precons in the order ntt.cpp:298-300 consumes (entry m + i), as host arrays -- what
    main.cpp:36-37,46-55 hands to ntt_input_kernel -- re-laid-out on the device for the kernels."""
    primes = Q[:2]
    c = A.Context(n, primes)
    P = O.Plan(n, primes)
    x = P.synthetic(6, seed=8)
    psi1 = O.primitive_roots_2n(n, primes[1], 5)[4]
    r, p = O.tables_u32(n, primes[1], psi1)
    ri, pi = O.tables_u32(n, primes[1], psi1, inverse=True)
    c.set_tables(1, r, p)                       # limb 1 forward: caller's roots and precons
    c.set_tables(1, ri, None, inverse=True)     # limb 1 inverse: precons computed by the library
    assert c.psi(1) == psi1 and c.psi(0) == O.min_psi(n, primes[0])
    d = to_dev(torch, x)
    c.fwd(d)
    y = to_np(d).reshape(x.shape)
    assert (y[:, 0, :] == P.fwd(x.copy())[:, 0, :]).all()                          # limb 0 untouched
    assert (y[:, 1, :] == O.fwd_u32_with_psi(x[:, 1, :], primes[1], psi1)).all()
    got = c.tables(1, inverse=True)
    assert (got[0] == ri).all() and (got[1] == pi).all()
    c.inv(d)
    assert (to_np(d).reshape(x.shape) == x).all()
    # inconsistent tables are refused and leave the previous ones in place
    bad = r.copy(); bad[n // 2 + 3] ^= 1
    with pytest.raises(A.AgxError):
        c.set_tables(1, bad, None)
    badp = p.copy(); badp[5] += 1
    with pytest.raises(A.AgxError):
        c.set_tables(1, r, badp)
    with pytest.raises(A.AgxError):
        c.set_tables(1, np.arange(n, dtype=np.uint32) + 2, None)                   # main.cpp:52's dummy "twiddles"
    with pytest.raises(A.AgxError):
        c.set_tables(0, r, p)                                                      # roots of another prime
    with pytest.raises(A.AgxError):
        c.set_tables(2, r, p)                                                      # limb out of range
    d = to_dev(torch, x)
    c.fwd(d)
    assert (to_np(d).reshape(x.shape) == y).all()
    c.close()


# ------------------------------------------------------------------------------- forward / inverse vs oracle

@pytest.mark.parametrize("n", [8, 32, 256, 1024, 2048, 4096, 8192, 32768])
@pytest.mark.parametrize("L", [1, 3])
def test_fwd_inv_vs_oracle(A, torch, n, L):
    primes = Q[:L]
    c = ctx_for(A, n, primes)
    P = O.Plan(n, primes)
    B = 37 if n <= 4096 else 5
    x = P.synthetic(B, seed=42)
    d = to_dev(torch, x)
    c.fwd(d)
    y = to_np(d).reshape(x.shape)
    y_ref = P.fwd(x.copy(), variant="barrett")
    assert (y == y_ref).all()
    assert (y == P.fwd(x.copy(), variant="shoup")).all()
    c.inv(d)
    assert (to_np(d).reshape(x.shape) == x).all()
    # inverse alone on oracle-produced spectra
    d2 = to_dev(torch, y_ref)
    c.inv(d2)
    assert (to_np(d2).reshape(x.shape) == P.inv(y_ref.copy(), variant="barrett")).all()


@pytest.mark.parametrize("n", [1024, 2048, 4096])
def test_survey_known_answers(A, torch, n):
    kat = {
        (Q[0], 1024): ("dadf656e0db1fcec", "87deb95a0431f1a7"), (Q[0], 2048): ("ef15ec72a1b8f75a", "4c0990d3c733f583"),
        (Q[0], 4096): ("f06769648ff96d35", "36a1101e7932a140"), (Q[1], 1024): ("f4a32865a5c55bd4", "ddd9d5aed8314f36"),
        (Q[1], 2048): ("add679a1500a8c8e", "9a0d9a9ff159beef"), (Q[1], 4096): ("43aee98a3e463189", "c01082f426ec6f16"),
        (Q[2], 1024): ("de6e87421e79aed7", "a04b585d08730040"), (Q[2], 2048): ("d80ea0fd5ba2706e", "8f486de34535d858"),
        (Q[2], 4096): ("9b587721c09a3c7a", "84177fb9278b7649"),
    }
    for q in Q:
        c = ctx_for(A, n, [q])
        ramp = np.arange(n, dtype=np.uint32)
        sm = np.array([O.splitmix64(42 + j) % q for j in range(n)], dtype=np.uint32)
        d = to_dev(torch, np.stack([ramp, sm]))
        c.fwd(d)
        out = to_np(d).reshape(2, n)
        assert O.sha16(out[0]) == kat[(q, n)][0] and O.sha16(out[1]) == kat[(q, n)][1]


def test_simple_vectors(A, torch):
    """zero -> zero, c*delta -> all c, X -> psi^(2 br(k)+1): the SEAL-Embedded-style checks."""
    for n in (1024, 2048, 4096):
        q = Q[1]
        c = ctx_for(A, n, [q])
        x = np.zeros((3, n), dtype=np.uint32)
        x[1, 0] = 7
        x[2, 1] = 1
        d = to_dev(torch, x)
        c.fwd(d)
        y = to_np(d).reshape(3, n)
        assert not y[0].any() and (y[1] == 7).all()
        psi, logn = c.psi(0), n.bit_length() - 1
        assert y[2].tolist() == [pow(psi, 2 * O.bitrev(k, logn) + 1, q) for k in range(n)]


def test_golden_from_reference_code(A, torch, golden):
    """The u32 GPU path reproduces what the reference's own ntt.cpp produced (tests/golden/make_golden.py)."""
    g, meta = golden
    checked = 0
    for name, m in meta.items():
        if name.startswith("main_dummy"):
            continue
        N, q, psi, kind, frames, seed = int(m[0]), int(m[1]), int(m[2]), m[3], int(m[4]), int(m[5])
        if kind == "lazy4q":
            continue   # [0,4q) inputs belong to the u64 reference-shaped path (test_ref_pipeline.py)
        gidx = np.arange(N * frames, dtype=np.uint64)
        if kind == "ramp":
            x = gidx % np.uint64(q)
        elif kind == "splitmix":
            x = np.array([O.splitmix64(seed + int(i)) % q for i in gidx], dtype=np.uint64)
        else:
            x = np.zeros(N * frames, dtype=np.uint64); x[::N] = 1
        c = ctx_for(A, N, [q])
        assert c.psi(0) == psi
        d = to_dev(torch, x.astype(np.uint32))
        c.fwd(d)
        assert (to_np(d) == g[name]).all(), name
        checked += 1
    assert checked >= 9


def test_edge_inputs(A, torch):
    n = 4096
    for L in (1, 3):
        primes = Q[:L]
        c = ctx_for(A, n, primes)
        P = O.Plan(n, primes)
        qv = np.array(primes, dtype=np.uint32).reshape(1, L, 1)
        # all q-1 (maximum reduced value)
        x = np.broadcast_to(qv - 1, (2, L, n)).copy()
        d = to_dev(torch, x); c.fwd(d)
        assert (to_np(d).reshape(x.shape) == P.fwd(x.copy())).all()
        # lazy inputs in [q, 2q) are accepted and mean the same residue
        base = P.synthetic(2, seed=9)
        lazy = (base + qv).astype(np.uint32)
        d = to_dev(torch, lazy); c.fwd(d)
        assert (to_np(d).reshape(base.shape) == P.fwd(base.copy())).all()
        d = to_dev(torch, lazy); c.inv(d)
        assert (to_np(d).reshape(base.shape) == P.inv(base.copy())).all()
        # empty batch is a no-op
        e = torch.empty(0, dtype=torch.int32, device="cuda")
        c.fwd(e); c.inv(e)
        # B = 1 and a ragged, non-multiple-of-anything batch
        for B in (1, 149):
            x = P.synthetic(B, seed=B)
            d = to_dev(torch, x); c.fwd(d)
            assert (to_np(d).reshape(x.shape) == P.fwd(x.copy())).all()


def test_outputs_always_reduced(A, torch):
    for n in (1024, 2048, 4096):
        c = ctx_for(A, n, Q)
        d = torch.empty(64 * 3 * n, dtype=torch.int32, device="cuda")
        c.fill_synthetic(d, seed=5)
        c.fwd(d)
        y = to_np(d).reshape(64, 3, n)
        for l, q in enumerate(Q):
            assert (y[:, l] < q).all()
        c.inv(d)
        y = to_np(d).reshape(64, 3, n)
        for l, q in enumerate(Q):
            assert (y[:, l] < q).all()


# ---------------------------------------------------------------------------------------------------- polymul

@pytest.mark.parametrize("n", [256, 1024, 2048, 4096])
def test_polymul_vs_oracle_and_schoolbook(A, torch, n):
    for L in (1, 3):
        primes = Q[:L]
        c = ctx_for(A, n, primes)
        P = O.Plan(n, primes)
        B = 9
        a, b = P.synthetic(B, seed=1), P.synthetic(B, seed=2)
        da, db = to_dev(torch, a), to_dev(torch, b)
        dc = torch.empty_like(da)
        c.polymul(dc, da, db)
        got = to_np(dc).reshape(a.shape)
        assert (got == P.polymul(a, b)).all()
        for l, q in enumerate(primes):
            assert (got[0, l] == O.polymul_schoolbook(a[0, l], b[0, l], q)).all()
        assert (to_np(da).reshape(a.shape) == a).all() and (to_np(db).reshape(a.shape) == b).all()
        if n >= 1024:   # the tuned paths allow every aliasing: c == a, c == b, and squaring a == b (== c)
            c.polymul(da, da, db)
            assert (to_np(da).reshape(a.shape) == got).all()
            da2 = to_dev(torch, a)
            c.polymul(db, da2, db)
            assert (to_np(db).reshape(a.shape) == got).all()
            sq = P.polymul(a, a)
            c.polymul(dc, da2, da2)
            assert (to_np(dc).reshape(a.shape) == sq).all()
            c.polymul(da2, da2, da2)
            assert (to_np(da2).reshape(a.shape) == sq).all()


@pytest.mark.parametrize("n", [256, 1024, 2048, 4096, 8192])
def test_polymul_by_spectrum(A, torch, n):
    """SURVEY s.8(f) rank 4: the product with one operand kept in evaluation form, c = INTT(NTT(a) .* b_hat), against the
    oracle's product and exact schoolbook, bit-identical to agx_polymul, with and without the tensor store's map (odd batch),
    every aliasing case, operands left untouched."""
    for L in (1, 3):
        primes = Q[:L]
        c = ctx_for(A, n, primes)
        P = O.Plan(n, primes)
        B = 7
        a, b = P.synthetic(B, seed=31), P.synthetic(B, seed=32)
        want = P.polymul(a, b)
        da, db = to_dev(torch, a), to_dev(torch, b)
        bh = db.clone(); c.fwd(bh)
        assert (to_np(bh).reshape(a.shape) == P.fwd(b.copy())).all()
        bh_np = to_np(bh).copy()
        dc = torch.empty_like(da)
        c.polymul_by_spectrum(dc, da, bh)
        got = to_np(dc).reshape(a.shape)
        assert (got == want).all()
        for l, q in enumerate(primes):
            assert (got[0, l] == O.polymul_schoolbook(a[0, l], b[0, l], q)).all()
        ref = torch.empty_like(da); c.polymul(ref, da, db)
        assert (to_np(ref) == to_np(dc)).all()
        assert (to_np(da).reshape(a.shape) == a).all() and (to_np(bh) == bh_np).all()
        t = da.clone(); c.polymul_by_spectrum(t, t, bh)            # c == a
        assert (to_np(t).reshape(a.shape) == want).all()
        t = bh.clone(); c.polymul_by_spectrum(t, da, t)            # c == b_hat
        assert (to_np(t).reshape(a.shape) == want).all()
        t = da.clone(); c.polymul_by_spectrum(t, t, t)             # all three: NTT(a) .* a (a read as a spectrum)
        ah = da.clone(); c.fwd(ah)
        e = torch.empty_like(da); c.elementwise("mul", e, ah, da); c.inv(e)
        assert (to_np(t) == to_np(e)).all()
        a2 = P.synthetic(B, seed=33)                               # the spectrum is reused: a second product, other a
        c.polymul_by_spectrum(dc, to_dev(torch, a2), bh)
        assert (to_np(dc).reshape(a.shape) == P.polymul(a2, b)).all()


def test_polymul_by_spectrum_host_and_errors(A, torch):
    n = 2048
    c = ctx_for(A, n, Q[:1])
    P = O.Plan(n, Q[:1])
    a, b = P.synthetic(5000, seed=41), P.synthetic(5000, seed=42)   # > one pipeline chunk (4096 polynomials)
    bh = P.fwd(b.copy(), threads=O.max_threads())
    out = np.empty_like(a)
    c.polymul_by_spectrum_host(out, a, bh)
    assert (out == P.polymul(a, b, threads=O.max_threads())).all()
    import ctypes
    L = A.lib()
    d = torch.zeros(2 * n + 4, dtype=torch.int32, device="cuda")
    ok, off = ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(d.data_ptr() + 4)
    assert L.agx_polymul_by_spectrum(c._h, ok, ok, off, 1, None) == -1      # AGX_E_INVALID: operand not 16-byte aligned
    assert L.agx_polymul_by_spectrum(c._h, off, ok, ok, 1, None) == -1
    assert L.agx_polymul_by_spectrum(c._h, ok, None, ok, 1, None) == -1
    assert L.agx_polymul_by_spectrum(c._h, ok, ok, ok, 0, None) == 0        # empty batch
    torch.cuda.synchronize()


def test_polymul_kat(A, torch):
    n, q = 2048, Q[0]
    c = ctx_for(A, n, [q])
    a = np.arange(1, n + 1, dtype=np.uint32)
    b = (2 * np.arange(n) + 1).astype(np.uint32)
    da, db = to_dev(torch, a), to_dev(torch, b)
    dc = torch.empty_like(da)
    c.polymul(dc, da, db)
    r = to_np(dc)
    assert r[:4].tolist() == [291855365, 287667213, 283483167, 279303231] and O.sha16(r) == "cc989f874ebe7af7"
    e1 = np.zeros(n, dtype=np.uint32); e1[n - 1] = 1
    e2 = np.zeros(n, dtype=np.uint32); e2[1] = 1
    c.polymul(dc, to_dev(torch, e1), to_dev(torch, e2))
    r = to_np(dc)
    assert r[0] == q - 1 and not r[1:].any()      # X^(n-1) * X = -1


# ---------------------------------------------------------------------------- synthetic fill, checksum, host API

def test_device_fill_and_checksum_match_oracle(A, torch):
    n = 2048
    c = ctx_for(A, n, Q)
    P = O.Plan(n, Q)
    d = torch.empty(11 * 3 * n, dtype=torch.int32, device="cuda")
    c.fill_synthetic(d, seed=1234, first_poly=5)
    x = P.synthetic(11, seed=1234, first_poly=5)
    assert (to_np(d).reshape(x.shape) == x).all()
    assert c.checksum(d, first_index=77) == O.checksum_u32(x, first_index=77)


@pytest.mark.parametrize("pinned", [False, True])
def test_host_entry_points(A, torch, pinned):
    n = 4096
    c = ctx_for(A, n, Q)
    P = O.Plan(n, Q)
    B = 2500                      # four chunks of the pipeline (37 MiB = 789 polys at L=3): exercises slot reuse
    x = P.synthetic(B, seed=3)
    want = P.fwd(x.copy(), threads=O.max_threads())
    if pinned:
        src = torch.from_numpy(x.view(np.int32)).pin_memory()
        dst = torch.empty_like(src).pin_memory()
        c.fwd_host(src, dst)
        assert (dst.numpy().view(np.uint32) == want).all()
        c.inv_host(dst)
        assert (dst.numpy().view(np.uint32) == x).all()
    else:
        out = np.empty_like(x)
        c.fwd_host(x, out)
        assert (out == want).all()
        c.inv_host(out)           # in place
        assert (out == x).all()


def test_polymul_host(A, torch):
    n = 2048
    c = ctx_for(A, n, Q[:1])
    P = O.Plan(n, Q[:1])
    a, b = P.synthetic(300, seed=11), P.synthetic(300, seed=12)
    out = np.empty_like(a)
    c.polymul_host(out, a, b)
    assert (out == P.polymul(a, b, threads=O.max_threads())).all()


def test_launch_count_counts(A, torch):
    c = ctx_for(A, 4096, Q[:1])
    d = torch.zeros(4 * 4096, dtype=torch.int32, device="cuda")
    before = c.launch_count()
    c.fwd(d); c.inv(d)
    assert c.launch_count() == before + 2


# ------------------------------------------------------------------- BASELINE.json configs at their full sizes

def _full_compare(A, torch, n, primes, B, seed):
    c = ctx_for(A, n, primes)
    P = O.Plan(n, primes)
    L = len(primes)
    d = torch.empty(B * L * n, dtype=torch.int32, device="cuda")
    c.fill_synthetic(d, seed=seed)
    chk_in = c.checksum(d)
    c.fwd(d)
    y = to_np(d).reshape(B, L, n)
    x = P.synthetic(B, seed=seed)
    assert O.checksum_u32(x) == chk_in
    want = P.fwd(x, threads=O.max_threads())          # in place on x
    assert (y == want).all()
    for l, q in enumerate(primes):
        assert int(y[:, l].max()) < q
    c.inv(d)
    assert c.checksum(d) == chk_in                    # round trip at full size: checksum of the whole batch
    del d
    torch.cuda.empty_cache()


def test_config2_n4096_b65536_full(A, torch):
    """configs[1]: n=4096, single prime, 65,536 polynomials (1 GiB), full bit-exact compare + round trip."""
    _full_compare(A, torch, 4096, Q[:1], 65536, 42)


def test_config3_rns3_b32768_full(A, torch):
    """configs[2]: n=4096, 3-limb RNS, batch 32,768: bit-exact vs the oracle per limb (full compare)."""
    _full_compare(A, torch, 4096, Q, 32768, 42)


def test_config4_polymul_n2048_b131072(A, torch):
    """configs[3]: negacyclic polymul n=2048, batch 131,072: 64-polynomial sample vs exact schoolbook, full batch vs
    the CPU NTT-path oracle, plus linearity (a*(b1+b2) = a*b1 + a*b2) as a size-independent property."""
    n, q, B = 2048, Q[0], 131072
    c = ctx_for(A, n, [q])
    P = O.Plan(n, [q])
    da = torch.empty(B * n, dtype=torch.int32, device="cuda")
    db = torch.empty_like(da)
    dc = torch.empty_like(da)
    c.fill_synthetic(da, seed=1234)
    c.fill_synthetic(db, seed=99)
    c.polymul(dc, da, db)
    got = to_np(dc).reshape(B, 1, n)
    a, b = P.synthetic(B, seed=1234), P.synthetic(B, seed=99)
    idx = np.linspace(0, B - 1, 64).astype(int)
    for i in idx:
        assert (got[i, 0] == O.polymul_schoolbook(a[i, 0], b[i, 0], q)).all()
    want = P.polymul(a, b, threads=O.max_threads())
    assert (got == want).all()
    del want
    # linearity on the first 4096 products
    m = 4096 * n
    b2 = torch.empty(m, dtype=torch.int32, device="cuda"); c.fill_synthetic(b2, seed=7)
    bs = ((db[:m].long() + b2.long()) % q).int()
    c1 = torch.empty(m, dtype=torch.int32, device="cuda"); c.polymul(c1, da[:m].contiguous(), bs)
    c2 = torch.empty(m, dtype=torch.int32, device="cuda"); c.polymul(c2, da[:m].contiguous(), b2)
    assert ((dc[:m].long() + c2.long()) % q == c1.long()).all()


@pytest.mark.parametrize("n", [1024, 2048])
def test_config5_sizes_roundtrip_large(A, torch, n):
    """configs[4] sweep sizes: 2^28/n polynomials (1 GiB) per GPU, round trip + sampled oracle compare."""
    B = (1 << 28) // n
    c = ctx_for(A, n, Q[:1])
    P = O.Plan(n, Q[:1])
    d = torch.empty(B * n, dtype=torch.int32, device="cuda")
    c.fill_synthetic(d, seed=1234)
    chk = c.checksum(d)
    c.fwd(d)
    lo = to_np(d[: 512 * n]).reshape(512, 1, n)
    hi = to_np(d[(B - 512) * n:]).reshape(512, 1, n)
    assert (lo == P.fwd(P.synthetic(512, seed=1234))).all()
    assert (hi == P.fwd(P.synthetic(512, seed=1234, first_poly=B - 512))).all()
    c.inv(d)
    assert c.checksum(d) == chk


# ------------------------------------------------------------------------------------- API behaviour on the GPU

def test_two_contexts_two_streams(A, torch):
    """No hidden global state: different (n, primes) contexts interleaved on different streams stay correct."""
    c1, c2 = ctx_for(A, 4096, Q[:1]), ctx_for(A, 2048, Q)
    P1, P2 = O.Plan(4096, Q[:1]), O.Plan(2048, Q)
    x1, x2 = P1.synthetic(300, seed=5), P2.synthetic(200, seed=6)
    d1, d2 = to_dev(torch, x1), to_dev(torch, x2)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(3):
        c1.fwd(d1, stream=s1); c2.fwd(d2, stream=s2)
        c1.inv(d1, stream=s1); c2.inv(d2, stream=s2)
    c1.fwd(d1, stream=s1); c2.fwd(d2, stream=s2)
    torch.cuda.synchronize()
    assert (to_np(d1).reshape(x1.shape) == P1.fwd(x1.copy(), threads=4)).all()
    assert (to_np(d2).reshape(x2.shape) == P2.fwd(x2.copy(), threads=4)).all()


def test_argument_errors(A, torch):
    import ctypes
    c = ctx_for(A, 4096, Q[:1])
    L = A.lib()
    assert L.agx_ntt_fwd(c._h, None, 4, None) == -1            # AGX_E_INVALID: null data with B > 0
    assert L.agx_ntt_fwd(c._h, None, 0, None) == 0             # empty batch is fine
    buf = torch.zeros(4096 + 4, dtype=torch.int32, device="cuda")
    assert L.agx_ntt_fwd(c._h, ctypes.c_void_p(buf.data_ptr() + 4), 1, None) == -1   # not 16-byte aligned
    assert L.agx_ntt_inv(c._h, ctypes.c_void_p(buf.data_ptr() + 8), 1, None) == -1
    assert L.agx_polymul(c._h, ctypes.c_void_p(16), None, None, 1, None) == -1
    ok, off = ctypes.c_void_p(buf.data_ptr()), ctypes.c_void_p(buf.data_ptr() + 4)
    for op in range(4):                                        # element-wise operands are read as 16-byte vectors too
        assert L.agx_elementwise(c._h, op, ok, off, ok, 1, None) == -1
        assert L.agx_elementwise(c._h, op, ok, ok, off, 1, None) == -1
        assert L.agx_elementwise(c._h, op, off, ok, ok, 1, None) == -1
    assert L.agx_elementwise(c._h, 4, ok, ok, ok, 1, None) == -1
    torch.cuda.synchronize()                                   # none of the refused calls may have poisoned the context
    assert L.agx_elementwise(c._h, 0, ok, ok, ok, 1, None) == 0
    torch.cuda.synchronize()
    assert L.agx_ntt_fwd_host(c._h, None, None, 3) == -1
    assert L.agx_ntt_fwd_host(c._h, None, None, 0) == 0
    assert L.agx_ntt_fwd(None, None, 0, None) == -1
    v = ctypes.c_uint32()
    assert L.agx_get_psi(c._h, 7, ctypes.byref(v)) == -1       # limb out of range
    with pytest.raises(ValueError):
        c.fwd(torch.zeros(4097, dtype=torch.int32, device="cuda"))      # not a whole number of polynomials
    with pytest.raises(ValueError):
        c.fwd(torch.zeros(4096, dtype=torch.int64, device="cuda"))      # wrong element size
    p = A.RefPipeline()
    assert L.agx_ntt_fwd(p._h, ctypes.c_void_p(16), 1, None) == -1      # table-less context: u32 calls refused
    p.close()


def test_elementwise_rns_ops(A, torch):
    """add / sub / mul / mac modulo each limb's prime vs numpy, then the classic use: a spectrum-domain
    multiply-accumulate equals the sum of two negacyclic products."""
    n = 2048
    c = ctx_for(A, n, Q)
    P = O.Plan(n, Q)
    B = 19
    a, b, acc = P.synthetic(B, seed=1), P.synthetic(B, seed=2), P.synthetic(B, seed=3)
    qv = np.array(Q, dtype=np.uint64).reshape(1, 3, 1)
    da, db, dacc = to_dev(torch, a), to_dev(torch, b), to_dev(torch, acc)
    out = torch.empty_like(da)
    a64, b64, acc64 = a.astype(np.uint64), b.astype(np.uint64), acc.astype(np.uint64)
    c.elementwise("add", out, da, db)
    assert (to_np(out).reshape(a.shape) == (a64 + b64) % qv).all()
    c.elementwise("sub", out, da, db)
    assert (to_np(out).reshape(a.shape) == (a64 + qv - b64) % qv).all()
    c.elementwise("mul", out, da, db)
    assert (to_np(out).reshape(a.shape) == (a64 * b64) % qv).all()
    c.elementwise("mac", dacc, da, db)
    assert (to_np(dacc).reshape(a.shape) == (acc64 + a64 * b64) % qv).all()
    c.elementwise("mul", da, da, da)                       # aliasing: in-place square
    assert (to_np(da).reshape(a.shape) == (a64 * a64) % qv).all()
    # a*b + a2*b2 through the spectrum domain
    a2, b2 = P.synthetic(B, seed=4), P.synthetic(B, seed=5)
    d = [to_dev(torch, v) for v in (a, b, a2, b2)]
    for t in d:
        c.fwd(t)
    c.elementwise("mul", d[0], d[0], d[1])
    c.elementwise("mac", d[0], d[2], d[3])
    c.inv(d[0])
    want = (P.polymul(a, b).astype(np.uint64) + P.polymul(a2, b2)) % qv
    assert (to_np(d[0]).reshape(a.shape) == want).all()


def test_five_limbs_and_27bit_primes(A, torch):
    """Limb counts that divide nothing, and smaller (27-bit) primes: the limb index is poly % L inside the kernels and
    the Barrett constants depend on the bit length of q."""
    five = (1053818881, 1054015489, 1054212097, 1055260673, 1056178177)      # SURVEY.md App. A candidates
    small = (134012929, 134111233, 134176769)                                # 27-bit, 2-adicity 13 -> n <= 4096
    for primes, n in ((five, 4096), (five, 2048), (small, 4096), (small, 1024)):
        c = ctx_for(A, n, primes)
        P = O.Plan(n, primes)
        x, y = P.synthetic(7, seed=21), P.synthetic(7, seed=22)
        d = to_dev(torch, x)
        c.fwd(d)
        assert (to_np(d).reshape(x.shape) == P.fwd(x.copy())).all()
        c.inv(d)
        assert (to_np(d).reshape(x.shape) == x).all()
        dx, dy = to_dev(torch, x), to_dev(torch, y)
        dz = torch.empty_like(dx)
        c.polymul(dz, dx, dy)
        assert (to_np(dz).reshape(x.shape) == P.polymul(x, y)).all()


@pytest.mark.parametrize("n", [1024, 2048, 4096])
def test_forward_without_tensor_store(A, torch, monkeypatch, n):
    """AGX_NO_TMA=1 at context creation selects the forward kernels that leave through the staging image and coalesced
    16-byte stores (the path taken when the driver offers no tensor-map encoder): same bits as the TMA-store kernels."""
    P = O.Plan(n, Q)
    x = P.synthetic(5, seed=77)
    want = P.fwd(x.copy())
    for no_tma in (False, True):
        if no_tma:
            monkeypatch.setenv("AGX_NO_TMA", "1")
        else:
            monkeypatch.delenv("AGX_NO_TMA", raising=False)
        c = A.Context(n, Q)
        d = torch.from_numpy(x.view(np.int32)).cuda()
        c.fwd(d)
        assert (to_np(d).reshape(x.shape) == want).all(), ("no_tma" if no_tma else "tma")
        a, b = P.synthetic(3, seed=5), P.synthetic(3, seed=6)
        da, db = torch.from_numpy(a.view(np.int32)).cuda(), torch.from_numpy(b.view(np.int32)).cuda()
        dc = torch.empty_like(da)
        c.polymul(dc, da, db)
        assert (to_np(dc).reshape(a.shape) == P.polymul(a, b)).all()
        c.close()


@pytest.mark.parametrize("n", [8192, 16384, 32768])
def test_large_sizes_register_pass_kernel_and_radix2_kernel_agree(A, torch, monkeypatch, n):
    """n >= 8192 runs ntt_big_kernel (the polynomial resident in one CTA's shared memory, passes of four stages on 16
    coefficients per thread); AGX_GENERIC_ONLY=1 at context creation selects the radix-2 kernel instead.  Both must give the
    oracle's spectra, inverses of arbitrary vectors, round trips and products, on 1 and 3 limbs, ragged batches."""
    primes = [q for q in (1053818881, 1054212097, 1055260673) if (q - 1) % (2 * n) == 0]
    for L in (1, len(primes)):
        P = O.Plan(n, primes[:L])
        x, z = P.synthetic(5, seed=31), P.synthetic(5, seed=32)
        want_f, want_i, want_p = P.fwd(x.copy()), P.inv(z.copy()), P.polymul(x.copy(), z.copy())
        for generic in (False, True):
            if generic:
                monkeypatch.setenv("AGX_GENERIC_ONLY", "1")
            else:
                monkeypatch.delenv("AGX_GENERIC_ONLY", raising=False)
            c = A.Context(n, primes[:L])
            assert c.variant() == ("generic" if generic else f"ntt_big<{n.bit_length() - 1}>")
            d = to_dev(torch, x)
            c.fwd(d)
            assert (to_np(d).reshape(x.shape) == want_f).all(), (generic, "forward")
            c.inv(d)
            assert (to_np(d).reshape(x.shape) == x).all(), (generic, "round trip")
            dz = to_dev(torch, z)
            c.inv(dz)
            assert (to_np(dz).reshape(z.shape) == want_i).all(), (generic, "inverse")
            lazy = to_dev(torch, (x.astype(np.uint64) + np.array(primes[:L], dtype=np.uint64).reshape(1, L, 1)).astype(np.uint32))
            c.fwd(lazy)
            assert (to_np(lazy).reshape(x.shape) == want_f).all(), (generic, "lazy inputs")
            dc = torch.empty_like(d)
            c.polymul(dc, to_dev(torch, x), to_dev(torch, z))
            assert (to_np(dc).reshape(x.shape) == want_p).all(), (generic, "product")
            c.close()


@pytest.mark.parametrize("n", [64, 1024, 4096, 32768])
def test_bitrev_adapter_gives_the_textbook_order(A, torch, n):
    """agx_bitrev permutes each polynomial by bit reversal in place: forward + bitrev is the natural-order spectrum
    NTT(x)[k] = sum_j x[j] psi^((2k+1)j) (checked against the big-int definition at small n), and it is an involution."""
    logn = n.bit_length() - 1
    c = A.Context(n, Q[:2])
    P = O.Plan(n, Q[:2])
    x = P.synthetic(3, seed=9)
    d = torch.from_numpy(x.view(np.int32)).cuda()
    c.fwd(d)
    y = to_np(d).reshape(x.shape).copy()
    c.bitrev(d)
    z = to_np(d).reshape(x.shape)
    perm = np.array([O.bitrev(k, logn) for k in range(n)])
    assert (z[..., perm] == y).all() and (z == y[..., perm]).all()
    if n == 64:
        for l, q in enumerate(Q[:2]):
            psi = c.psi(l)
            want = [sum(int(x[0, l, j]) * pow(psi, (2 * k + 1) * j, q) for j in range(n)) % q for k in range(n)]
            assert z[0, l].tolist() == want
    c.bitrev(d)
    assert (to_np(d).reshape(x.shape) == y).all()
    c.inv(d)
    assert (to_np(d).reshape(x.shape) == x).all()
    c.close()


def _ntt_primes(count, n, below=1 << 30):
    """The `count` largest primes q < below with q = 1 (mod 2n) (deterministic Miller-Rabin for 32-bit q)."""
    def is_prime(q):
        if q < 2:
            return False
        for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
            if q % p == 0:
                return q == p
        d, s = q - 1, 0
        while d % 2 == 0:
            d, s = d // 2, s + 1
        for a in (2, 3, 5, 7):
            x = pow(a, d, q)
            if x in (1, q - 1):
                continue
            for _ in range(s - 1):
                x = x * x % q
                if x == q - 1:
                    break
            else:
                return False
        return True
    out, q = [], (below - 1) // (2 * n) * (2 * n) + 1
    while len(out) < count:
        if is_prime(q):
            out.append(q)
        q -= 2 * n
    return tuple(out)


@pytest.mark.parametrize("n,L,B", [(4096, 64, 1), (4096, 7, 3), (2048, 64, 2), (1024, 33, 5)])
def test_many_limbs_odd_batches(A, torch, n, L, B):
    """The maximum limb count (64), limb counts and batches that divide nothing, the largest primes below 2^30:
    per-limb tables generated on the device, limb index = polynomial % L inside every kernel, odd transform counts."""
    primes = _ntt_primes(L, n)
    assert max(primes) < (1 << 30) and len(set(primes)) == L
    c = ctx_for(A, n, primes)
    P = O.Plan(n, primes)
    x, y = P.synthetic(B, seed=31), P.synthetic(B, seed=32)
    d = to_dev(torch, x)
    c.fwd(d)
    assert (to_np(d).reshape(x.shape) == P.fwd(x.copy())).all()
    c.inv(d)
    assert (to_np(d).reshape(x.shape) == x).all()
    dx, dy = to_dev(torch, x), to_dev(torch, y)
    dz = torch.empty_like(dx)
    c.polymul(dz, dx, dy)
    assert (to_np(dz).reshape(x.shape) == P.polymul(x, y)).all()
    out = np.empty_like(x)
    c.fwd_host(x.copy(), out)
    assert (out == P.fwd(x.copy())).all()


@pytest.mark.parametrize("n", [256, 8192])
def test_generic_sizes_polymul_aliasing(A, torch, n):
    """every aliasing case agx_polymul documents also holds on the generic (any-n) path."""
    primes = Q[:2]
    c = ctx_for(A, n, primes)
    P = O.Plan(n, primes)
    a, b = P.synthetic(5, seed=21), P.synthetic(5, seed=22)
    want, sq = P.polymul(a, b), P.polymul(a, a)
    da, db = to_dev(torch, a), to_dev(torch, b)
    out = torch.empty_like(da)
    c.polymul(out, da, db)
    assert (to_np(out).reshape(a.shape) == want).all()
    t = da.clone(); c.polymul(t, t, db)                    # out == a
    assert (to_np(t).reshape(a.shape) == want).all()
    t = db.clone(); c.polymul(t, da, t)                    # out == b
    assert (to_np(t).reshape(a.shape) == want).all()
    c.polymul(out, da, da)                                 # a == b
    assert (to_np(out).reshape(a.shape) == sq).all()
    t = da.clone(); c.polymul(t, t, t)                     # all three
    assert (to_np(t).reshape(a.shape) == sq).all()
    assert (to_np(da).reshape(a.shape) == a).all() and (to_np(db).reshape(a.shape) == b).all()


def test_measured_butterfly_peak_is_sane(A, torch):
    """The integer roofline bench.py reports is measured on the GPU in front of it: IMAD.HI at half rate caps the u32
    butterfly at 16 per clock per SM; the isolated stream must land under that and above 10."""
    c = ctx_for(A, 4096, Q[:1])
    v8, mhz = c.measure_butterfly_peak(0, 1024)
    v4, _ = c.measure_butterfly_peak(0, 512)
    assert 10.0 < v4 <= v8 * 1.02 and v8 <= 16.5, (v4, v8)
    assert 500 < mhz < 3000
    u64, _ = c.measure_butterfly_peak(1, 1024)
    assert 0.5 < u64 < v8
