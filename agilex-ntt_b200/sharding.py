"""Batch partitioning across GPUs: the multi-GPU story of the hot path.

Every (polynomial, limb) transform is independent, so the batch dimension B is split contiguously over the ranks
(a polynomial's L limbs stay together), tables are replicated per GPU, and there is NO data-path collective.
This mirrors the reference's own mini-batch split over compute units (ntt.cpp:526-536: numFrames/CU plus one for
the first numFrames%CU units), made contiguous instead of round-robin so each rank's shard is one HBM range.
The only cross-rank traffic is optional: summing per-rank 64-bit checksums for the parity check.
"""
from __future__ import annotations


def shard_bounds(B: int, world: int, rank: int) -> tuple[int, int]:
    """[first, last) polynomial range of `rank`: sizes differ by at most one, earlier ranks get the extras."""
    if world <= 0 or not (0 <= rank < world) or B < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(B, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def all_shards(B: int, world: int) -> list[tuple[int, int]]:
    return [shard_bounds(B, world, r) for r in range(world)]


def combine_checksums(parts) -> int:
    """Checksums are sums mod 2^64 over globally indexed elements, so shard checksums simply add."""
    return sum(int(p) for p in parts) % (1 << 64)
