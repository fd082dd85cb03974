"""CPU: host-side logic — method-string parsing, bit allocation, trainer, synthetic generators, sharding helpers."""
import numpy as np
import pytest

from helpers import orc
from vaq_b200 import synth, train
from vaq_b200.sharded import make_keys_f32, make_keys_u32, shard_bounds, split_keys


def test_parse_method_string():
    p = train.parse_method_string("VAQ256m32min7max13var1,EA_TI1000")     # scripts/run_demos.sh-style strings
    assert (p["budget"], p["M"], p["min_bits"], p["max_bits"], p["var"]) == (256, 32, 7, 13, 1.0)
    assert p["methods"] == {"EA", "TI"} and p["ti_clusters"] == 1000
    p = train.parse_method_string("VAQ128m16min6max10var1,HEAP")
    assert p["methods"] == {"HEAP"} and p["budget"] == 128
    with pytest.raises(ValueError):
        train.parse_method_string("VAQ64m16min1max4var1,FAST")


def brute_force_alloc(v, budget, lo, hi):
    import itertools
    best, arg = -1, None
    kd = [max(train.next_pow2(v[i] / v[i + 1]), 0) for i in range(len(v) - 1)]
    for x in itertools.product(range(lo, hi + 1), repeat=len(v)):
        if sum(x) != budget or any(x[i] - x[i + 1] > kd[i] for i in range(len(v) - 1)):
            continue
        val = sum(a * b for a, b in zip(v, x))
        if val > best + 1e-12:
            best, arg = val, x
    return best, arg


def test_allocate_bits_is_optimal_for_the_reference_ilp():
    rng = np.random.default_rng(0)
    for _ in range(6):
        v = np.sort(rng.random(5))[::-1] + 0.05
        v /= v.sum()
        budget = int(rng.integers(12, 28))
        bits = train.allocate_bits(v, budget, 2, 7)
        best, arg = brute_force_alloc(v, budget, 2, 7)
        assert bits.sum() == budget and bits.min() >= 2 and bits.max() <= 7
        assert abs(float((v * bits).sum()) - best) < 1e-9, (bits, arg)


def test_train_shapes_and_descending_bits():
    X = synth.decaying_gaussian(3000, 32, seed=3)
    model, XP = train.train(X, 64, 8, 5, 10, kmeans_iters=3)
    assert model.L == 4 and model.bits.sum() == 64 and model.D == 32
    assert all(c.shape == (1 << int(b), 4) for c, b in zip(model.centroids, model.bits))
    assert (np.diff(model.bits) <= 0).all()           # variance-descending subspaces get at least as many bits
    np.testing.assert_allclose(XP, X @ model.eig, rtol=1e-5, atol=1e-5)
    # the trained codebooks are usable by the oracle's encode (nearest centroid, VAQ.cpp:728-748): every code in range
    om = orc.Model(model.L, model.bits, model.centroids)
    codes = orc.Port().encode(om, XP[:200])
    assert codes.shape == (200, 8) and (codes < (1 << model.bits.astype(np.int64))[None, :]).all()


def test_synthetic_generators_are_counter_based():
    bits = [9, 4, 7, 1]
    a = synth.synth_codes(bits, 1000, 5000, 42)
    b = synth.synth_codes(bits, 100, 5500, 42)
    assert np.array_equal(a[500:600], b)                  # any row slice regenerates identically
    assert (a.max(0) < (1 << np.array(bits))).all()
    cdf = synth.code_cdf(a, bits)
    c = synth.synth_codes(bits, 20000, 0, 7, cdf)
    h = np.bincount(c[:, 0], minlength=512) / 20000
    h0 = np.bincount(a[:, 0], minlength=512) / 1000
    assert np.abs(h - h0).max() < 0.01
    w = synth.synth_bitvectors(50, 10, 100, 3)
    assert w.shape == (50, 2) and (w[:, 1] >> np.uint64(36)).max() == 0
    assert np.array_equal(synth.synth_bitvectors(10, 30, 100, 3), w[20:30])


def test_shard_bounds_and_key_format():
    assert shard_bounds(10, 4) == [0, 3, 6, 9, 10]
    assert shard_bounds(8, 8)[-1] == 8 and shard_bounds(3, 8) == [0, 1, 2, 3, 3, 3, 3, 3, 3]
    d = np.array([0.0, 1.5, 1.5, 3.0e38], np.float32)
    ids = np.array([7, 2, 1, 2 ** 31 - 1], np.int32)
    keys = make_keys_f32(d, ids)
    order = np.argsort(keys)
    assert order.tolist() == [0, 2, 1, 3]                 # ascending (distance, id)
    i2, d2 = split_keys(keys)
    assert np.array_equal(i2, ids) and np.array_equal(d2, d)
    hk = make_keys_u32(np.array([3, 3, 1]), np.array([5, 4, 9]))
    assert np.argsort(hk).tolist() == [2, 1, 0]


def test_fp16_seed_bound_is_an_upper_bound():
    """Numerics of the fp16 filter kernel's bound seeding (csrc/adc_filter16_scan.cu): entries are round-toward-zero fp16
    of scale * entry, summed in half precision with round-to-nearest additions in subspace order; the seed
    (acc * (1 + (M + 3) * 2^-11) + M * 2^-24) / scale must never fall below the real distance — including subnormal and
    zero entries and M = 64, the largest model the seeding accepts.  (numpy's float16 addition is the correctly rounded
    sum, like HADD2.)"""
    rng = np.random.default_rng(5)

    def rz16(x):
        h = x.astype(np.float16)
        up = h.astype(np.float64) > x
        h[up] = np.nextafter(h[up], np.float16(-np.inf))
        return h

    worst = 0.0
    for M in (4, 9, 32, 64):
        for scale_exp in (-20, -3, 0, 7, 20):
            scale = 2.0 ** scale_exp
            for spread in (1e-9, 1e-4, 1.0):
                ent = (rng.random((4000, M)) ** 3 * spread * 200.0 / scale).astype(np.float32)      # scaled entries up to ~200
                ent[:, rng.integers(0, M)] = 0.0
                # adversarial half: entries just below the next fp16 value (round-toward-zero loses almost a whole ulp)
                h = (ent[2000:].astype(np.float64) * scale).astype(np.float16).astype(np.float64)
                ent[2000:] = (np.nextafter(np.nextafter(h.astype(np.float16), np.float16(np.inf)).astype(np.float32), np.float32(0)) / scale).astype(np.float32)
                true = ent.astype(np.float64).sum(1)
                e16 = rz16(ent.astype(np.float64) * scale)
                acc = np.zeros(ent.shape[0], np.float16)
                for s in range(M):
                    acc = (acc + e16[:, s]).astype(np.float16)
                ok = np.isfinite(acc.astype(np.float64))
                ub = (acc.astype(np.float64) * (1.0 + (M + 3) * 2.0 ** -11) + M * 2.0 ** -24) / scale
                assert (ub[ok] >= true[ok]).all(), (M, scale_exp, spread)
                nz = ok & (true > 0)
                worst = max(worst, float((true[nz] / ub[nz]).max()))
    assert worst <= 1.0


def test_fp16_filter_margins_never_drop_a_pair_under_the_bound():
    """Numerics of the fp16 lower-bound filter (csrc/adc_filter16_scan.cu, header comment): a (row, query) pair is dropped
    when its half-precision partial sum exceeds RU_fp16(thr * scale * (1 + m)), m = 2^-9 after the 4 fields of stage 1,
    2^-7 after 8 (level 1), 2^-5 after up to 32 (level 2, M <= 32), 2^-3 beyond.  A dropped pair's real partial sum — and
    with it the distance the reference computes — must exceed thr.  Checked on sums sitting right at the bound."""
    rng = np.random.default_rng(11)

    def rz16(x):
        h = x.astype(np.float16)
        up = h.astype(np.float64) > x
        h[up] = np.nextafter(h[up], np.float16(-np.inf))
        return h

    def ru16(x):
        h = x.astype(np.float16)
        dn = h.astype(np.float64) < x
        h[dn] = np.nextafter(h[dn], np.float16(np.inf))
        return h

    for nf, m in ((4, 2.0 ** -9), (8, 2.0 ** -7), (32, 2.0 ** -5), (64, 2.0 ** -3)):
        for scale_exp in (-12, 0, 9):
            scale = 2.0 ** scale_exp
            n = 20000
            ent = (rng.random((n, nf)) ** 2 * 300.0 / nf / scale).astype(np.float32)
            true = ent.astype(np.float64).sum(1)
            # thresholds around each row's own sum: the decisions that matter are the near-ties
            thr = (true * (1.0 + (rng.random(n) - 0.8) * 4.0 * m)).astype(np.float32)
            e16 = rz16(ent.astype(np.float64) * scale)
            acc = e16[:, 0].copy()
            for s in range(1, nf):
                acc = (acc + e16[:, s]).astype(np.float16)
            bound = ru16((thr.astype(np.float32) * np.float32(scale) * np.float32(1.0 + m)).astype(np.float64))
            dropped = acc.astype(np.float64) > bound.astype(np.float64)
            assert dropped.any() and (~dropped).any()
            assert (true[dropped] > thr[dropped].astype(np.float64)).all(), (nf, scale_exp)
