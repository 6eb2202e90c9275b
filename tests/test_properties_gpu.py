"""Size-independent properties of the transforms at BASELINE.json's full sizes, computed entirely on the GPU through the
C ABI and compared through device checksums / exact device compares (no CPU work proportional to the batch):
linearity of the forward transform, the convolution theorem tying agx_polymul to agx_ntt_fwd / agx_elementwise /
agx_ntt_inv, multiplication by X as a negacyclic shift, and forward-inverse-forward idempotence.  The oracle enters only
to anchor one small slice of each batch, so that the properties cannot be satisfied by a consistently wrong transform."""
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


def _np(t):
    return t.cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("n,L,B", [(4096, 1, 65536), (4096, 3, 32768 // 2), (2048, 1, 131072), (1024, 1, 262144)])
def test_forward_is_linear_at_full_size(A, torch, n, L, B):
    """NTT(a + b) = NTT(a) + NTT(b) (mod q, per limb) over a whole BASELINE-sized batch; anchored on 16 polynomials."""
    primes = Q[:L]
    c = A.Context(n, primes)
    a = torch.empty(B * L * n, dtype=torch.int32, device="cuda")
    b = torch.empty_like(a)
    s = torch.empty_like(a)
    c.fill_synthetic(a, seed=11)
    c.fill_synthetic(b, seed=12)
    c.elementwise("add", s, a, b)
    P = O.Plan(n, primes)
    want16 = P.fwd(((P.synthetic(16, seed=11).astype(np.uint64) + P.synthetic(16, seed=12)) %
                    np.array(primes, dtype=np.uint64).reshape(1, L, 1)).astype(np.uint32))
    c.fwd(a); c.fwd(b); c.fwd(s)
    c.elementwise("add", a, a, b)                       # a <- NTT(a) + NTT(b)
    assert bool((a == s).all())
    assert (_np(s[: 16 * L * n]).reshape(16, L, n) == want16).all()
    c.close()


@pytest.mark.parametrize("n,L,B", [(4096, 3, 8192), (2048, 1, 131072), (1024, 2, 65536)])
def test_convolution_theorem_at_full_size(A, torch, n, L, B):
    """agx_polymul(a, b) == INTT(NTT(a) .* NTT(b)) assembled from the separate entry points; anchored by exact schoolbook."""
    primes = Q[:L]
    c = A.Context(n, primes)
    a = torch.empty(B * L * n, dtype=torch.int32, device="cuda")
    b = torch.empty_like(a)
    c.fill_synthetic(a, seed=21)
    c.fill_synthetic(b, seed=22)
    prod = torch.empty_like(a)
    c.polymul(prod, a, b)
    P = O.Plan(n, primes)
    xa, xb = P.synthetic(2, seed=21), P.synthetic(2, seed=22)
    for i in range(2):
        for l, q in enumerate(primes):
            assert (O.polymul_schoolbook(xa[i, l], xb[i, l], q) == _np(prod[: 2 * L * n]).reshape(2, L, n)[i, l]).all()
    c.fwd(a); c.fwd(b)
    c.elementwise("mul", a, a, b)
    c.inv(a)
    assert bool((a == prod).all())
    c.close()


@pytest.mark.parametrize("n", [1024, 2048, 4096])
def test_multiplying_by_x_is_a_negacyclic_shift(A, torch, n):
    """a(X) * X mod (X^n + 1): coefficients move up by one place and the one that wraps around changes sign."""
    q = Q[1]
    c = A.Context(n, [q])
    B = 4096
    a = torch.empty(B * n, dtype=torch.int32, device="cuda")
    c.fill_synthetic(a, seed=31)
    x = torch.zeros(B * n, dtype=torch.int32, device="cuda")
    x.view(B, n)[:, 1] = 1
    out = torch.empty_like(a)
    c.polymul(out, a, x)
    av, ov = a.view(B, n).to(torch.int64), out.view(B, n).to(torch.int64)
    assert bool((ov[:, 1:] == av[:, :-1]).all())
    assert bool((ov[:, 0] == (q - av[:, -1]) % q).all())
    c.close()


def test_forward_inverse_forward(A, torch):
    """fwd o inv o fwd == fwd on lazy inputs in [0, 2q) (what the entry points accept), cfg 2's size."""
    n, B = 4096, 65536
    q = Q[0]
    c = A.Context(n, [q])
    d = torch.empty(B * n, dtype=torch.int32, device="cuda")
    c.fill_synthetic(d, seed=41)
    d += (torch.arange(B * n, device="cuda", dtype=torch.int32) & 1) * q      # every other coefficient in [q, 2q)
    c.fwd(d)
    s1 = c.checksum(d)
    assert int(d.max()) < q and int(d.min()) >= 0
    c.inv(d); c.fwd(d)
    assert c.checksum(d) == s1
    c.close()
