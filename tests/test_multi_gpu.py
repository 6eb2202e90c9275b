"""Hardware shard-and-compare: the batch is split over ranks the way bench.py splits it
(contiguous shards, replicated tables, no collective -- the contiguous version of the reference's mini-batch split,
ntt.cpp:526-536, 581, 624), every shard is transformed by its own context on its own GPU, and the concatenation must
equal the single-GPU transform of the whole batch byte for byte, the oracle's, and -- through checksums with global
indices -- what bench.py all-reduces.  Uses ALL visible GPUs; with fewer GPUs than shards the shards share GPUs (each
still has its own context, stream and device buffers), so the test never skips."""
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


@pytest.mark.parametrize("n,L,B", [(4096, 1, 203), (4096, 3, 61), (2048, 1, 77), (1024, 3, 130)])
def test_shards_equal_single_gpu_forward_and_inverse(A, torch, n, L, B):
    ngpu = torch.cuda.device_count()
    world = max(ngpu, 4)                                    # at least four shards, every visible GPU used
    primes = Q[:L]
    P = O.Plan(n, primes)
    whole = P.synthetic(B, seed=1234)                       # the global batch (bench.py's timing seed)
    want = P.fwd(whole.copy(), threads=4)

    c0 = A.Context(n, primes, device=0)
    with torch.cuda.device(0):
        d0 = torch.from_numpy(whole.view(np.int32)).cuda(0)
        c0.fwd(d0)
        single = _np(d0).reshape(whole.shape).copy()
    assert (single == want).all()
    total_sum = O.checksum_u32(want)

    parts, sums, ctxs, bufs = [], [], [], []
    for rank in range(world):
        dev = rank % ngpu
        lo, hi = A.shard_bounds(B, world, rank)
        c = A.Context(n, primes, device=dev)
        with torch.cuda.device(dev):
            d = torch.empty((hi - lo) * L * n, dtype=torch.int32, device=f"cuda:{dev}")
            c.fill_synthetic(d, seed=1234, first_poly=lo)   # shard = slice of the global synthetic batch, made on its GPU
            assert (_np(d).reshape(-1, L, n) == whole[lo:hi]).all()
            c.fwd(d)
            sums.append(c.checksum(d, first_index=lo * L * n))
            parts.append(_np(d).reshape(-1, L, n).copy())
        ctxs.append(c); bufs.append(d)
    cat = np.concatenate(parts)
    assert cat.tobytes() == single.tobytes()                # concatenated shards == single-GPU output, byte for byte
    assert A.combine_checksums(sums) == total_sum           # == what bench.py all-reduces and compares on rank 0
    assert sorted({b.device.index for b in bufs}) == list(range(ngpu))

    # and back: every shard's inverse restores its slice; device selection is left as the caller had it
    torch.cuda.set_device(0)
    for rank, (c, d) in enumerate(zip(ctxs, bufs)):
        lo, hi = A.shard_bounds(B, world, rank)
        c.inv(d, stream=torch.cuda.current_stream(d.device))
        assert torch.cuda.current_device() == 0             # the C ABI restores the caller's current device
        assert (_np(d).reshape(-1, L, n) == whole[lo:hi]).all()
        c.close()
    c0.close()


def test_device_is_restored_by_every_entry_point(A, torch):
    """the C ABI must not retarget the caller's current device (two contexts in one process)."""
    ngpu = torch.cuda.device_count()
    dev = ngpu - 1
    torch.cuda.set_device(0)
    c = A.Context(1024, Q[:1], device=dev)
    assert torch.cuda.current_device() == 0
    x = O.Plan(1024, Q[:1]).synthetic(3, seed=3)
    d = torch.from_numpy(x.view(np.int32)).cuda(dev)
    s = torch.cuda.current_stream(dev)
    c.fwd(d, stream=s); c.inv(d, stream=s); c.bitrev(d, stream=s); c.bitrev(d, stream=s)
    c.checksum(d, stream=s); c.tables(0); c.fwd_host(x.copy())
    assert torch.cuda.current_device() == 0
    import ctypes
    cur = ctypes.c_int(-1)
    rt = ctypes.CDLL("libcudart.so.12")
    assert rt.cudaGetDevice(ctypes.byref(cur)) == 0 and cur.value == 0
    assert (_np(d).reshape(x.shape) == x).all()
    c.close()
    assert torch.cuda.current_device() == 0


def test_generic_sizes_on_every_visible_gpu(A, torch):
    """One context per GPU inside one process (the C ABI's multi-GPU model when the caller is not multi-process), sizes
    served by the generic kernel: 32 KB of dynamic shared memory at n = 8192, the > 48 KB opt-in path at 16384 -- function
    attributes are per-device state."""
    for nn in (8192, 16384):
        P = O.Plan(nn, Q[:1])
        x = P.synthetic(6, seed=2)
        want = P.fwd(x.copy())
        for dev in range(torch.cuda.device_count()):
            c = A.Context(nn, Q[:1], device=dev)
            d = torch.from_numpy(x.view(np.int32)).cuda(dev)
            c.fwd(d, stream=torch.cuda.current_stream(dev))
            assert (_np(d).reshape(x.shape) == want).all()
            c.close()
