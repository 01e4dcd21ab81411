"""CPU model of the two-level bucket sort of the MSM sort phase (snark_verifier_b200/csrc/sort2.cuh, sort mode 1).

The CUDA kernels cannot run here, so this restates their index arithmetic — partition = high bits of the bucket id, tiles of
4096 digits with per-tile partition histograms laid out [window][partition][tile], exclusive scans over tiles and partitions,
the partition-ordered (reference, low-key) stream, one block per (window, partition) for the final counting sort — in numpy and
checks the result against a plain stable sort: every bucket's run holds exactly its terms, sign bits intact, and the bucket
counts / offsets are what the single-level path (histogram atomics + scan) produces."""
import numpy as np
import pytest

TILE = 4096


def plan(c):
    nb_bits = c - 1                       # buckets per window = 2^(c-1), ids 1..NB
    lb = min(nb_bits, 8)                  # low key bits (<= 256 local buckets)
    return nb_bits, lb, 1 << (nb_bits - lb)


def two_level_sort(digits, c, rng):
    """digits: [W, nv] uint32 (|d| | sign << 31, 0 = skip).  Returns (counts [W, NB], offsets [W, NB], sorted [W, nv])."""
    W, nv = digits.shape
    nb_bits, lb, P = plan(c)
    NB = 1 << nb_bits
    T = (nv + TILE - 1) // TILE
    tilecnt = np.zeros((W, P, T), dtype=np.int64)
    # L1a: per-tile partition histogram
    for w in range(W):
        for t in range(T):
            e = digits[w, t * TILE:(t + 1) * TILE]
            d = e & 0x7FFFFFFF
            hb = (d[d != 0] - 1) >> lb
            tilecnt[w, :, t] = np.bincount(hb, minlength=P)
    # L1b / L1c: exclusive scan over tiles, then over partitions
    parttot = tilecnt.sum(axis=2)
    tilebase = np.cumsum(tilecnt, axis=2) - tilecnt
    partbase = np.zeros((W, P + 1), dtype=np.int64)
    partbase[:, 1:] = np.cumsum(parttot, axis=1)
    # L1d: scatter (reference, low key) into partition order; the order inside one (tile, partition) is whatever the atomics give
    rec = np.zeros((W, nv), dtype=np.uint32)
    key = np.zeros((W, nv), dtype=np.uint8)
    for w in range(W):
        for t in range(T):
            cur = partbase[w, :P] + tilebase[w, :, t]
            idx = np.arange(t * TILE, min((t + 1) * TILE, nv))
            for i in rng.permutation(idx):          # arbitrary arrival order
                e = int(digits[w, i])
                d = e & 0x7FFFFFFF
                if d == 0:
                    continue
                hb = (d - 1) >> lb
                pos = cur[hb]
                cur[hb] += 1
                rec[w, pos] = i | (e & 0x80000000)
                key[w, pos] = (d - 1) & ((1 << lb) - 1)
            assert np.array_equal(cur, partbase[w, :P] + tilebase[w, :, t] + tilecnt[w, :, t])
    # L2a: per-partition histogram of the low keys = the bucket counts; existing scan -> offsets
    counts = np.zeros((W, NB), dtype=np.int64)
    for w in range(W):
        for p in range(P):
            k = key[w, partbase[w, p]:partbase[w, p + 1]]
            counts[w, p << lb:(p + 1) << lb] = np.bincount(k, minlength=1 << lb)
    offsets = np.cumsum(counts, axis=1) - counts
    # L2b: one block per (window, partition) places the references
    out = np.full((W, nv), 0xFFFFFFFF, dtype=np.uint32)
    for w in range(W):
        for p in range(P):
            cur = offsets[w, p << lb:(p + 1) << lb].copy()
            for j in rng.permutation(np.arange(partbase[w, p], partbase[w, p + 1])):
                k = key[w, j]
                out[w, cur[k]] = rec[w, j]
                cur[k] += 1
    return counts, offsets, out


def check(digits, c, counts, offsets, out):
    W, nv = digits.shape
    NB = 1 << (c - 1)
    for w in range(W):
        d = digits[w] & 0x7FFFFFFF
        assert np.array_equal(counts[w], np.bincount(d[d != 0] - 1, minlength=NB))
        assert np.array_equal(offsets[w], np.cumsum(counts[w]) - counts[w])
        total = int(counts[w].sum())
        assert np.all(out[w, total:] == 0xFFFFFFFF)
        got_idx = out[w, :total] & 0x7FFFFFFF
        # every slot of bucket b's run holds a term whose digit is b + 1, with that term's sign bit
        bucket_of_slot = np.repeat(np.arange(NB), counts[w])
        assert np.array_equal(d[got_idx] - 1, bucket_of_slot)
        assert np.array_equal(out[w, :total] >> 31, digits[w, got_idx] >> 31)
        assert len(np.unique(got_idx)) == total          # a permutation of the non-zero terms


@pytest.mark.parametrize("c,nv", [(17, 3 * TILE + 77), (16, 2 * TILE), (13, TILE - 5), (10, 5000), (9, 300), (4, 1000), (2, 50)])
def test_random_digits(c, nv):
    rng = np.random.default_rng(c * 1000 + nv)
    NB = 1 << (c - 1)
    W = 3
    mag = rng.integers(0, NB + 1, size=(W, nv), dtype=np.uint32)           # 0 = skipped term
    sign = rng.integers(0, 2, size=(W, nv), dtype=np.uint32) << 31
    digits = (mag | np.where(mag != 0, sign, 0)).astype(np.uint32)
    check(digits, c, *two_level_sort(digits, c, rng))


def test_skewed_digits():
    rng = np.random.default_rng(5)
    c, nv = 16, 2 * TILE + 9
    digits = np.full((2, nv), 12345, dtype=np.uint32)                      # every term in one bucket (one partition)
    digits[1, ::3] = (1 << 15) | (1 << 31)                                  # top bucket, negative
    digits[1, 1::7] = 0
    check(digits, c, *two_level_sort(digits, c, rng))
