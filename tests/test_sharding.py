"""Multi-GPU partition logic on CPU: shard bounds, and a world_size-2 gloo run where each rank transforms its shard
(with the oracle standing in for the device kernels -- this tests the HOST logic: bounds, global seeding, checksum
combination) and the concatenation equals the single-rank result byte for byte."""
import os
import socket

import numpy as np
import pytest

from oracle import oracle as O


def test_shard_bounds():
    import agilex_ntt_b200 as A
    for B in (0, 1, 7, 65536, 65537):
        for w in (1, 2, 3, 4, 8):
            sh = A.all_shards(B, w)
            assert sh[0][0] == 0 and sh[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
            sizes = [b - a for a, b in sh]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        A.shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, n, q, ret):
    import torch
    import torch.distributed as dist
    import agilex_ntt_b200 as A
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = A.shard_bounds(B, world, rank)
    P = O.Plan(n, [q])
    x = P.synthetic(hi - lo, seed=42, first_poly=lo)          # global seeding: shard == slice of the whole
    y = P.fwd(x.copy())
    part = O.checksum_u32(y, first_index=lo * n)
    t = torch.tensor([part & 0xFFFFFFFF, part >> 32], dtype=torch.int64)
    parts = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(parts, t)
    # max-over-ranks timing reduction used by bench.py
    tm = torch.tensor([float(rank + 1)])
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    if rank == 0:
        ret["sum"] = A.combine_checksums(int(p[0]) | (int(p[1]) << 32) for p in parts)
        ret["tmax"] = float(tm)
    ret[f"y{rank}"] = y
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_compare():
    import torch.multiprocessing as mp
    B, n, q = 11, 1024, O.SEAL_PRIMES_30[0]
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, n, q, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    P = O.Plan(n, [q])
    whole = P.fwd(P.synthetic(B, seed=42))
    assert (np.concatenate([ret["y0"], ret["y1"]]) == whole).all()
    assert ret["sum"] == O.checksum_u32(whole)
    assert ret["tmax"] == 2.0


def test_shard_bounds_properties():
    """Property test of the partition: disjoint cover, order-preserving, balanced."""
    from hypothesis import given, settings, strategies as st
    import agilex_ntt_b200 as A

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 10**7), st.integers(1, 64))
    def prop(B, w):
        sh = A.all_shards(B, w)
        assert sum(b - a for a, b in sh) == B
        assert all(0 <= a <= b <= B for a, b in sh)
        assert all(x[1] == y[0] for x, y in zip(sh, sh[1:]))
        assert max(b - a for a, b in sh) - min(b - a for a, b in sh) <= 1
    prop()
    assert A.combine_checksums([2**64 - 1, 2]) == 1
