"""Chunk partition of an MSM across ranks — the rayon shape of util/msm.rs:322-336 (`chunk_size = ceil(n / threads)`,
`scalars.chunks(chunk_size).zip(bases.chunks(chunk_size))`), with one rank per GPU instead of one rayon thread."""


def chunk_bounds(n, world, rank):
    """-> (start, length) of `rank`'s contiguous slice of n terms; trailing ranks may be empty when n < world."""
    chunk = (n + world - 1) // world
    lo = min(rank * chunk, n)
    return lo, min(chunk, n - lo)
