"""Seeded random differential test: the CUDA path through the C ABI against the CPU oracle on parameter combinations
no hand-written case lists -- random size (tuned and generic kernels), limb count, batch, primes, per-limb roots
(minimal or not), lazy inputs -- forward, inverse, round trip and product.  Bit-exact, tolerance 0."""
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


@pytest.mark.parametrize("logn,primes,B,root_pick,seed", _cases(48, 20261018))
def test_random_parameters_against_the_oracle(A, torch, logn, primes, B, root_pick, seed):
    n, L = 1 << logn, len(primes)
    for q in primes:
        assert O.is_prime(q) and (q - 1) % (2 * n) == 0
    # root_pick 0: the library's minimal roots; otherwise the caller's (root_pick + l)-th smallest primitive 2n-th root per limb
    if root_pick == 0 or n > 8192:
        psis, c = [O.min_psi(n, q) for q in primes], A.Context(n, primes)
    else:
        psis = [O.primitive_roots_2n(n, q, root_pick + l + 1)[root_pick + l] for l, q in enumerate(primes)]
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
        assert (back(dc) == P.polymul(x.copy(), z.copy())).all(), "product"
    c.close()
