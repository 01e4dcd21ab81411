"""CPU model of the batched-affine bucket accumulation scheme (snark_verifier_b200/csrc/bucket_affine.cuh).

The CUDA kernel cannot run here, so this restates its control flow — the arithmetic placement of per-task scratch regions, the
level / batch loops driven by the block-wide maximum, the pair classification (chord, tangent, identity operands, opposite
points), Montgomery's trick with exclusive prefixes — with Python integers and checks it against plain point addition
(oracle/bn254_model.py).  The GPU parity suite (tests/test_gpu_parity.py, accumulate modes 2 and 3) checks the kernel itself."""
import random

from oracle import bn254_model as m

P = m.P


def classify(a, b):
    """ba_classify: (kind, d).  Points are (x, y) with (0, 0) the identity."""
    if a == (0, 0):
        return 2, None
    if b == (0, 0):
        return 3, None
    d = (b[0] - a[0]) % P
    if d != 0:
        return 0, d
    if a[1] == b[1] and a[1] != 0:
        return 1, (2 * a[1]) % P
    return 4, None


def block_reduce(lists, K, pairs_min, threads):
    """One block of `threads` tasks; returns the per-task remaining items after the affine levels (to be folded serially)."""
    cur = [list(l) for l in lists] + [[] for _ in range(threads - len(lists))]
    mmax = max(len(l) for l in cur)
    inversions = 0
    while (mmax >> 1) >= pairs_min:
        maxpairs = mmax >> 1
        nxt = [[None] * ((len(l) + 1) // 2) for l in cur]
        for cb in range(0, maxpairs, K):
            runs, prefs = [], []
            for l in cur:  # forward pass of every thread
                pairs = len(l) >> 1
                lo, hi = min(cb, pairs), min(cb + K, pairs)
                run, pref = 1, {}
                for i in range(lo, hi):
                    kind, d = classify(l[2 * i], l[2 * i + 1])
                    if kind <= 1:
                        pref[i - lo] = run
                        run = run * d % P
                runs.append(run)
                prefs.append(pref)
            total = 1
            for r in runs:
                total = total * r % P
            assert total != 0
            inversions += 1
            tinv = pow(total, P - 2, P)
            for t, l in enumerate(cur):  # what the product tree hands to thread t: 1 / runs[t]
                others = 1
                for u, r in enumerate(runs):
                    if u != t:
                        others = others * r % P
                acc = tinv * others % P
                assert acc * runs[t] % P == 1
                pairs = len(l) >> 1
                lo, hi = min(cb, pairs), min(cb + K, pairs)
                for i in range(hi - 1, lo - 1, -1):
                    a, b = l[2 * i], l[2 * i + 1]
                    kind, d = classify(a, b)
                    if kind <= 1:
                        inv = acc * prefs[t][i - lo] % P
                        acc = acc * d % P
                        assert inv * d % P == 1
                        num = (b[1] - a[1]) % P if kind == 0 else 3 * a[0] * a[0] % P
                        lam = num * inv % P
                        x3 = (lam * lam - a[0] - b[0]) % P
                        y3 = (lam * (a[0] - x3) - a[1]) % P
                        o = (x3, y3)
                    elif kind == 2:
                        o = b
                    elif kind == 3:
                        o = a
                    else:
                        o = (0, 0)
                    nxt[t][i] = o
        for t, l in enumerate(cur):
            if len(l) & 1:
                nxt[t][len(l) >> 1] = l[-1]
        cur = nxt
        assert all(x is not None for l in cur for x in l)
        mmax = (mmax >> 1) + (mmax & 1)
        assert mmax == max(len(l) for l in cur)
    return cur[:len(lists)], inversions


def fold(items):
    acc = None
    for p in items:
        acc = m.g1_add(acc, None if p == (0, 0) else p)
    return acc


def rand_point(rng, pool):
    r = rng.random()
    if r < 0.08:
        return (0, 0)
    p = rng.choice(pool)
    if rng.random() < 0.4:
        p = (p[0], (-p[1]) % P)
    return p


def test_tree_reduction_matches_plain_addition():
    rng = random.Random(7)
    pool = [m.g1_mul(m.G1_GEN, rng.randrange(1, 1 << 40)) for _ in range(6)]  # small pool: equal and opposite neighbours occur
    for trial in range(12):
        threads = 4
        lists = [[rand_point(rng, pool) for _ in range(rng.choice([1, 2, 3, 5, 8, 13, 16, 17, 31]))] for _ in range(rng.randrange(1, threads + 1))]
        rest, inversions = block_reduce(lists, K=3, pairs_min=2, threads=threads)
        for l, r in zip(lists, rest):
            assert fold(l) == fold(r)
        if max(len(l) for l in lists) >= 4:
            assert inversions > 0


def test_all_equal_and_all_opposite_lists():
    g = m.g1_mul(m.G1_GEN, 12345)
    ng = (g[0], (-g[1]) % P)
    rest, _ = block_reduce([[g] * 16, [g, ng] * 8, [(0, 0)] * 9 + [g]], K=4, pairs_min=1, threads=4)
    assert fold(rest[0]) == m.g1_mul(g, 16)
    assert fold(rest[1]) is None
    assert fold(rest[2]) == g


def region_starts(pos, slot):
    sa = (pos + slot + 1) >> 1
    sb = (sa + slot + 1) >> 1
    return sa, sb


def test_scratch_regions_are_disjoint_and_in_bounds():
    rng = random.Random(11)
    for trial in range(200):
        T = rng.choice([1, 2, 3, 7, 64])
        nb = rng.randrange(1, 40)
        counts = [rng.choice([0, 0, 1, 2, 3, 5, rng.randrange(0, 4 * T + 2)]) for _ in range(nb)]
        n = sum(counts) + rng.randrange(0, 5)  # zero digits never enter the sorted array
        cap = nb + n // T + 1
        stride_a = (n + cap) // 2 + 2
        stride_b = (stride_a + cap) // 2 + 2
        pos, slot = 0, 0
        owner_a, owner_b = {}, {}
        for c in counts:
            first = 0
            while first < c:
                mlen = min(T, c - first)
                sa, sb = region_starts(pos + first, slot)
                la = (mlen + 1) // 2
                lb = (la + 1) // 2
                for j in range(sa, sa + la):
                    assert j not in owner_a and j < stride_a
                    owner_a[j] = slot
                for j in range(sb, sb + lb):
                    assert j not in owner_b and j < stride_b
                    owner_b[j] = slot
                first += mlen
                slot += 1
            pos += c
        assert slot <= cap
