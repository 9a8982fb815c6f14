"""CPU tests of the host-side logic: tile plan coverage, key packing, instance helpers."""
import numpy as np
import pytest

from tsp_optimization_b200.engine import key_pack, key_unpack, tile_plan, tile_plan_ex
from tsp_optimization_b200.instances import (is_tour, order_to_succ, random_tours, succ_to_order,
                                             uniform_instance)


def covered_pairs(n, TI, TJ, rs, rj, tiles=None):
    """Set of (p,q) the BI kernel evaluates for the given tile ids (all by default): a pair is evaluated
    by tile (I,J) iff p in its row block, q in its column block and q >= p+2, q <= n-1, not (0, n-1)."""
    ntr = len(rj)
    cover = np.zeros((n, n), dtype=np.int32)
    ids = range(int(rs[-1])) if tiles is None else tiles
    for t in ids:
        I = int(np.searchsorted(rs, t, side="right") - 1)
        J = int(rj[I] + (t - rs[I]))
        p_lo, p_hi = I * TI, min((I + 1) * TI, n)
        q_lo, q_hi = J * TJ, min((J + 1) * TJ, n)
        for p in range(p_lo, p_hi):
            lo = max(q_lo, p + 2)
            if lo < q_hi:
                cover[p, lo:q_hi] += 1
    if n > 1:
        cover[0, n - 1] = 0 if n < 3 else cover[0, n - 1] - 1
    return cover


@pytest.mark.parametrize("n,T,R,TJ", [(52, 256, 2, 32), (299, 64, 2, 64), (700, 256, 2, 64), (1500, 64, 4, 64),
                                      (2100, 128, 8, 128), (513, 64, 8, 64), (2100, 256, 8, 128), (130, 64, 2, 32)])
def test_tile_plan_covers_every_pair_exactly_once(n, T, R, TJ):
    t, r, tj, rs, rj = tile_plan(n, R, TJ, threads=T)
    assert (t, r, tj) == (T, R, TJ)
    cover = covered_pairs(n, T * R, TJ, rs, rj)
    want = np.zeros((n, n), dtype=np.int32)
    for p in range(n):
        want[p, p + 2:] = 1
    want[0, n - 1] = 0
    assert (cover == want).all()
    assert int(want.sum()) == n * (n - 3) // 2


@pytest.mark.parametrize("n,R,TJ", [(130, 2, 32), (700, 2, 64), (1500, 4, 64), (2100, 8, 128), (513, 8, 64), (1021, 8, 88), (1275, 4, 32)])
def test_row_shuffle_tile_plan_covers_every_pair_exactly_once(n, R, TJ):
    """the row-shuffle variant (csrc/kernels_bi_scan.cuh): a warp owns 32R - 1 rows — lane L the rows [L R, L R + R) of the warp's
    range, lane 31's last row is masked and belongs to the next warp — so a 64-thread tile-row is 2 (32R - 1) positions; with
    that height the plan must still cover every non-adjacent pair exactly once, and the lanes' rows must tile the tile-row."""
    T = 64
    t, r, tj, ti, rs, rj = tile_plan_ex(n, R, TJ, threads=T, row_shuffle=1)
    assert (t, r, tj, ti) == (T, R, TJ, (T // 32) * (32 * R - 1))
    owned = []
    for warp in range(T // 32):
        for lane in range(32):
            rows = list(range(warp * (32 * R - 1) + lane * R, warp * (32 * R - 1) + lane * R + R))
            owned += rows[:-1] if lane == 31 else rows
    assert sorted(owned) == list(range(ti))          # every row of the tile-row exactly once
    cover = covered_pairs(n, ti, TJ, rs, rj)
    want = np.zeros((n, n), dtype=np.int32)
    for p in range(n):
        want[p, p + 2:] = 1
    want[0, n - 1] = 0
    assert (cover == want).all()
    # the plain plan of the same shape is a different one (taller tile-rows): the option really changes the plan
    assert tile_plan_ex(n, R, TJ, threads=T, row_shuffle=0)[3] == T * R
    assert tile_plan_ex(100000)[:4] == (64, 8, 256, 510)


def test_tile_plan_round_robin_sharding_partitions_the_tiles():
    n, T, R, TJ = 1500, 64, 4, 64
    _, _, _, rs, rj = tile_plan(n, R, TJ, threads=T)
    nt = int(rs[-1])
    total = np.zeros((n, n), dtype=np.int32)
    for world in (2, 4, 8):
        total[:] = 0
        for rank in range(world):
            total += covered_pairs(n, T * R, TJ, rs, rj, tiles=range(rank, nt, world)) + 0
        # (0, n-1) is subtracted once per rank by the helper: fix up
        total[0, n - 1] = 0
        assert total.max() == 1 and int(total.sum()) == n * (n - 3) // 2


def test_auto_tile_shape():
    """the cost model keeps R = 8 (1.125 sqrt per move) from mid-size tours upwards, fills every resident block at least
    once, and lands on the big-instance shape at n = 100 000."""
    for n, world in ((10000, 1), (20000, 1), (50000, 2), (100000, 1), (100000, 8)):
        T, R, TJ, rs, rj = tile_plan(n, world=world)
        assert R >= 8, (n, world, T, R, TJ)
    assert tile_plan(100000)[:3] == (64, 8, 256)
    assert tile_plan(52)[:2] == (64, 2)


def test_key_pack_orders_like_the_reference_scan():
    """uint64 min == (lowest delta, then lowest i, then lowest j): reference tabusearch.c:126-156."""
    rng = np.random.default_rng(0)
    keys = [(int(rng.integers(-50, 0)), int(rng.integers(0, 131072)), int(rng.integers(0, 131072))) for _ in range(2000)]
    keys += [(-7, 5, 9), (-7, 5, 8), (-7, 4, 100000), (-(1 << 27) + 1, 131071, 131071)]
    packed = [key_pack(*k) for k in keys]
    assert min(packed) == key_pack(*min(keys))
    assert sorted(keys) == [key_unpack(p) for p in sorted(packed)]
    assert all(p < (1 << 62) for p in packed)  # two spare bits: the generation of the peer-memory exchange word
    # "no move" sentinel loses against every negative delta
    assert key_pack(0, 0x1FFFF, 0x1FFFF) > max(p for p, k in zip(packed, keys) if k[0] < 0)


def test_instance_helpers():
    xy = uniform_instance(1000)
    assert xy.shape == (1000, 2) and xy.min() >= 0 and xy.max() <= 10000
    assert (xy == uniform_instance(1000)).all()
    assert (xy == np.floor(xy)).all()
    t = random_tours(50, 8, seed=3)
    for b in range(8):
        assert is_tour(t[b])
        o = succ_to_order(t[b])
        assert (order_to_succ(o) == t[b]).all()
    assert not is_tour(np.array([1, 0, 3, 2], dtype=np.int32))


def test_glibc_random_restated_equals_libc():
    """GlibcRandom == srandom()/random() of the C library (the generator behind the reference's URAND / rand_choice), and
    reference_random_population == random_generation() (reference src/genetic.c:349-364) driven by it."""
    import ctypes
    from tsp_optimization_b200.instances import GlibcRandom, reference_random_population
    libc = ctypes.CDLL("libc.so.6")
    libc.random.restype = ctypes.c_long
    for seed in (1, 123, 0, 2**31 + 5, 4242):
        libc.srandom(ctypes.c_uint(seed))
        g = GlibcRandom(seed)
        assert [libc.random() for _ in range(500)] == [g.random() for _ in range(500)], seed
    n = 97
    pop = reference_random_population(n, 3, 123)
    libc.srandom(123)
    for b in range(3):
        c = list(range(n))
        for _ in range(n):
            i1 = int(libc.random() / 2147483647.0 * n)
            i2 = int(libc.random() / 2147483647.0 * n)
            c[i1], c[i2] = c[i2], c[i1]
        assert pop[b].tolist() == c
