"""GPU parity: the CUDA path (through the C ABI) against the oracle port and the reference
fixtures.  Integer work (codes, ids) bit-exact; distances per SURVEY Appendix B."""
import numpy as np
import pytest

from helpers import RTOL, assert_knn_equiv, bitwise_equal, golden_model, load_golden, orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def port():
    return orc.Port()


def make_index(m, eig=None, codes=None):
    from vaq_b200.index import VAQIndex
    ix = VAQIndex(m.L, m.bits, m.centroids, eig=eig)
    if codes is not None:
        ix.add_codes(codes)
    return ix


def random_model(rng, M, L, bits):
    cents = [rng.standard_normal((1 << b, L)).astype(np.float32) * (1.0 + 3.0 / (1 + s)) for s, b in enumerate(bits)]
    return orc.Model(L, np.asarray(bits, np.int32), cents)


def random_codes(rng, m, n):
    return np.stack([rng.integers(0, 1 << int(b), size=n) for b in m.bits], 1).astype(np.uint16)


# ---- codes ------------------------------------------------------------------------------------

@pytest.mark.parametrize("bits", [
    [8] * 16,                                  # 128 bits, byte aligned
    [9, 9, 9, 8, 8, 7, 7, 7] * 4,              # 256 bits, straddling fields
    [13, 11, 10, 9, 7, 5, 3, 2, 1, 15, 14, 12],  # every odd width, 102 bits -> one padded word
    [15] * 68,                                 # 1020 bits, 8 words
    [1] * 5,
])
@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 70001])
def test_pack_roundtrip_bit_exact(bits, n):
    rng = np.random.default_rng(len(bits) * 1000 + n)
    m = random_model(rng, len(bits), 1, bits)
    codes = random_codes(rng, m, n)
    ix = make_index(m)
    half = n // 2
    ix.add_codes(codes[:half])      # two appends: exercises growth + unaligned row0
    ix.add_codes(codes[half:])
    assert ix.num_rows == n
    assert ix.row_bytes == 16 * ((sum(bits) + 127) // 128)
    assert np.array_equal(ix.get_codes(), codes)
    if n > 40:
        assert np.array_equal(ix.get_codes(7, 30), codes[7:37])
    ix.close()


def test_encode_bit_exact_vs_port_and_reference(port):
    for case in ("vaq_small_a", "vaq_small_b"):
        g = load_golden(case)
        m, _ = golden_model(g)
        ix = make_index(m)
        ix.encode_add(g["XP"])
        got = ix.get_codes()
        want_port = port.encode(m, g["XP"])
        assert np.array_equal(got, want_port), f"{case}: device encode != oracle port"
        # vs the compiled reference's VAQ::encode output (fixture): bit-exact
        assert np.array_equal(got, g["codes"]), f"{case}: {(got != g['codes']).sum()} codes differ from VAQ::encode"
        ix.close()


def test_synthetic_codes_match_host_generator():
    from vaq_b200 import synth
    rng = np.random.default_rng(5)
    bits = [9, 8, 7, 6, 10, 3, 12, 5]
    m = random_model(rng, len(bits), 2, bits)
    for cdf in (None, synth.code_cdf(random_codes(rng, m, 5000), bits)):
        ix = make_index(m)
        ix.set_id_base(1_000_000_007)
        ix.add_synthetic(5000, seed=99, cdf=cdf)
        ix.add_synthetic(3001, seed=99, cdf=cdf)
        got = ix.get_codes()
        want = synth.synth_codes(bits, 8001, 1_000_000_007, 99, cdf)
        assert np.array_equal(got, want)
        ix.close()


# ---- LUT ---------------------------------------------------------------------------------------

@pytest.mark.parametrize("case", ["vaq_small_a", "vaq_small_b"])
def test_lut_vs_reference(port, case):
    g = load_golden(case)
    m, _ = golden_model(g)
    ix = make_index(m)
    lut = ix.build_lut(g["Q"])
    assert bitwise_equal(lut, port.create_lut(m, g["Q"])), "device LUT != oracle port (bitwise)"
    for s in range(m.M):
        a = lut[:, m.lut_off[s]:m.lut_off[s + 1]]
        b = g["lut"][:, m.lut_off[s]:m.lut_off[s + 1]]
        if m.K[s] >= 8:
            assert bitwise_equal(a, b), f"subspace {s}: not bit-exact vs the reference's AVX2 path"
        else:
            np.testing.assert_allclose(a, b, rtol=RTOL, atol=0)
    ix.close()


@pytest.mark.parametrize("L", [1, 2, 3, 4, 6, 8, 12, 15])
def test_lut_all_sublens(port, L):
    rng = np.random.default_rng(L)
    bits = [10, 7, 4, 3, 2, 1]
    m = random_model(rng, len(bits), L, bits)
    q = rng.standard_normal((9, m.D)).astype(np.float32)
    ix = make_index(m)
    assert bitwise_equal(ix.build_lut(q), port.create_lut(m, q))
    ix.close()


# ---- search --------------------------------------------------------------------------------------

def check_search(port, m, codes, Q, k, flags, ix=None, id_base=0):
    from vaq_b200.index import PROJECTED
    own = ix is None
    if own:
        ix = make_index(m, codes=codes)
        if id_base:
            ix.set_id_base(id_base)
    from vaq_b200.index import SCAN_F32, SCAN_V1
    want_lab, want_dis = port.search_lex(m, codes, Q, k, id_base)
    # default kernel choice (fp16 lower-bound filter when it fits), the fp32 filter kernel, the lane-per-row kernel
    for extra, kern in ((0, None), (SCAN_F32, None), (SCAN_V1, 1)):
        lab, dis = ix.search(Q, k, flags | PROJECTED | extra)
        assert np.array_equal(lab, want_lab), f"ids differ from the canonical (distance, id) top-k (extra={extra:#x})"
        assert bitwise_equal(dis, want_dis), "distances are not bit-identical to the oracle's summation order"
        if kern:
            assert ix.last_config()["scan_kernel"] == kern
    if own:
        ix.close()
    return lab, dis


@pytest.mark.parametrize("case", ["vaq_small_a", "vaq_small_b"])
@pytest.mark.parametrize("mode", ["HEAP", "EA"])
def test_search_vs_reference_fixture(port, case, mode):
    from vaq_b200.index import EA, HEAP
    g = load_golden(case)
    m, _ = golden_model(g)
    k = int(g["k"])
    lab, dis = check_search(port, m, g["codes"], g["Q"], k, HEAP if mode == "HEAP" else EA)
    key = "heap" if mode == "HEAP" else "ea"
    assert_knn_equiv(lab, dis, g[f"lab_{key}"], g[f"dis_{key}"], what=f"{case}/{mode} vs reference")


def test_search_raw_queries_device_projection(port):
    from vaq_b200.index import EA
    g = load_golden("vaq_small_a")
    m, eig = golden_model(g)
    ix = make_index(m, eig=eig, codes=g["codes"])
    lab, dis = ix.search(g["Qraw"], int(g["k"]), EA)      # raw queries: (X * V) on the device
    # projection order differs from Eigen's GEMM -> tolerance-only
    np.testing.assert_allclose(dis, g["dis_ea"], rtol=2e-4)
    assert (lab == g["lab_ea"]).mean() > 0.95
    ix.close()


@pytest.mark.parametrize("n,k", [(1, 1), (5, 10), (31, 7), (32, 32), (33, 100), (1000, 100), (4097, 1), (20000, 10)])
@pytest.mark.parametrize("mode", ["HEAP", "EA"])
def test_search_ragged_sizes(port, n, k, mode):
    from vaq_b200.index import EA, HEAP
    rng = np.random.default_rng(n * 31 + k)
    bits = [9, 8, 8, 7, 6, 6, 5, 4]
    m = random_model(rng, 8, 4, bits)
    codes = random_codes(rng, m, n)
    Q = rng.standard_normal((5, m.D)).astype(np.float32) * 2
    lab, dis = check_search(port, m, codes, Q, k, HEAP if mode == "HEAP" else EA)
    if n < k:      # unfilled slots as utils/Heap.hpp:230-233,344-347 leaves them
        assert (lab[:, n:] == -1).all()
        assert (dis[:, n:] == np.finfo(np.float32).max).all()


@pytest.mark.parametrize("nq", [1, 2, 3, 4, 5, 7, 8, 9, 17])
@pytest.mark.parametrize("bits", [[6, 6, 5, 5], [9, 8, 8, 7, 6, 6, 5, 4], [11, 11, 10, 10, 9, 9, 8, 8, 7, 7, 7, 7, 6, 6, 6, 6]])
def test_search_query_tile_shapes(port, nq, bits):
    """Partial query tiles (nq not a multiple of T) and every tile width the planner can pick."""
    from vaq_b200.index import EA
    rng = np.random.default_rng(nq * 100 + len(bits))
    m = random_model(rng, len(bits), 2, bits)
    codes = random_codes(rng, m, 9000)
    Q = rng.standard_normal((nq, m.D)).astype(np.float32)
    check_search(port, m, codes, Q, 10, EA)


def test_search_multi_chunk_and_large_k(port):
    """Few queries on many rows -> several row chunks per query tile, thresholds carried between chunks."""
    from vaq_b200.index import EA, PROJECTED
    rng = np.random.default_rng(99)
    m = random_model(rng, 8, 4, [9, 8, 8, 7, 6, 6, 5, 4])
    codes = random_codes(rng, m, 300000)
    Q = rng.standard_normal((3, m.D)).astype(np.float32) * 2
    ix = make_index(m, codes=codes)
    check_search(port, m, codes, Q, 10, EA, ix=ix)
    ix.search(Q, 10, EA | PROJECTED)
    assert ix.last_config()["row_chunks"] > 1
    check_search(port, m, codes, Q, 300, EA, ix=ix)
    check_search(port, m, codes, Q[:1], 1500, EA, ix=ix)
    ix.close()


def test_search_fp16_filter_extreme_scales(port):
    """The fp16 lower-bound tables are rescaled per query: tiny and huge distance scales, zero tables."""
    from vaq_b200.index import EA, PROJECTED
    rng = np.random.default_rng(77)
    bits = [8, 8, 7, 7, 6, 6, 5, 5, 4, 4, 4, 4]
    for mult in (1e-12, 1e-4, 1.0, 3e4, 1e15):
        cents = [(rng.standard_normal((1 << b, 3)) * mult * (1 + 5.0 / (1 + s))).astype(np.float32) for s, b in enumerate(bits)]
        m = orc.Model(3, np.asarray(bits, np.int32), cents)
        codes = random_codes(rng, m, 40000)
        Q = (rng.standard_normal((9, m.D)) * mult * 2).astype(np.float32)
        Q[3] = 0
        ix = make_index(m, codes=codes)
        check_search(port, m, codes, Q, 10, EA, ix=ix)
        ix.search(Q, 10, EA | PROJECTED)
        assert ix.last_config()["scan_kernel"] == 3
        ix.close()
    # all-zero tables for one subspace and a query equal to a centroid (exact zeros in the LUT)
    cents = [rng.standard_normal((1 << b, 3)).astype(np.float32) for b in bits]
    cents[0][:] = 0
    m = orc.Model(3, np.asarray(bits, np.int32), cents)
    codes = random_codes(rng, m, 20000)
    Q = rng.standard_normal((8, m.D)).astype(np.float32)
    Q[0] = np.concatenate([c[5 % c.shape[0]] for c in cents])
    Q[:, :3] = 0
    check_search(port, m, codes, Q, 10, EA)


def test_search_seed_bound_with_subnormal_table_entries(port):
    """Adversarial case for the fp16 seed bound: 40 000 duplicates of one code tuple whose table entries, once scaled,
    fall below the fp16 subnormal step (2^-24) and round toward zero.  Every sampled row then accumulates exactly 0;
    a seed bound of ~0 would make the exact level reject the true neighbours (their distance is tiny but non-zero).
    The seed carries an absolute slack of M * 2^-24 for exactly this (adc_filter16_scan.cu, "bound seeding")."""
    from vaq_b200.index import EA, PROJECTED
    rng = np.random.default_rng(2025)
    bits = [8, 8, 7, 7, 6, 6, 5, 5]
    m = random_model(rng, len(bits), 4, bits)
    tup = np.array([rng.integers(0, 1 << b) for b in bits], np.uint16)
    centre = np.concatenate([m.centroids[s][tup[s]] for s in range(m.M)])
    codes = np.tile(tup, (40000, 1))
    codes[::997] = random_codes(rng, m, codes[::997].shape[0])           # a few unrelated rows in between
    Q = np.stack([centre + np.float32(d) for d in (1e-7, 3e-6, 1e-5, 1e-4, 0.0)] +
                 [rng.standard_normal(m.D).astype(np.float32) for _ in range(3)]).astype(np.float32)
    ix = make_index(m, codes=codes)
    lab, dis = check_search(port, m, codes, Q, 10, EA, ix=ix)
    ix.search(Q, 10, EA | PROJECTED)
    assert ix.last_config()["scan_kernel"] == 3
    dup_ids = np.nonzero((codes == tup).all(1))[0][:10]
    for q in range(4):
        assert dis[q, 0] > 0 and np.array_equal(lab[q], dup_ids)          # non-zero distance, lowest duplicate ids
    ix.close()


def test_search_duplicate_rows_tie_rule(port):
    """Exact ties: every row repeated 4x -> canonical order keeps the lowest ids."""
    from vaq_b200.index import EA
    rng = np.random.default_rng(3)
    m = random_model(rng, 8, 2, [6] * 8)
    base = random_codes(rng, m, 500)
    codes = np.concatenate([base] * 4)
    Q = rng.standard_normal((8, m.D)).astype(np.float32)
    check_search(port, m, codes, Q, 10, EA)


def test_search_many_queries_and_id_base(port):
    from vaq_b200.index import EA
    rng = np.random.default_rng(8)
    m = random_model(rng, 16, 2, [8] * 16)
    codes = random_codes(rng, m, 30000)
    Q = rng.standard_normal((700, m.D)).astype(np.float32)
    check_search(port, m, codes, Q, 10, EA, id_base=123456789)


def test_search_spill_path(port):
    """sum K_s * 4 B beyond the shared-memory budget -> trailing tables served from L2 (GIST-512 shape)."""
    from vaq_b200.index import EA, HEAP
    rng = np.random.default_rng(21)
    bits = [14, 14, 13, 13, 12, 12, 8, 8]        # 53 K + ... floats > 203 KB budget
    m = random_model(rng, 8, 2, bits)
    codes = random_codes(rng, m, 50000)
    Q = rng.standard_normal((6, m.D)).astype(np.float32)
    ix = make_index(m, codes=codes)
    from vaq_b200.index import PROJECTED
    check_search(port, m, codes, Q, 10, EA, ix=ix)
    ix.search(Q, 10, EA | PROJECTED)
    cfg = ix.last_config()
    assert cfg["spill_lut_floats"] > 0 and cfg["scan_kernel"] == 2, cfg      # tables too large for the fp16 tile of 8
    check_search(port, m, codes, Q, 10, HEAP, ix=ix)
    assert ix.last_config()["spill_lut_floats"] > 0
    ix.close()


def test_search_M_not_multiple_of_4(port):
    """The reference misreads when mHighestSubs % 4 != 0; the device path is well defined:
    same grouping with a short last group."""
    from vaq_b200.index import EA
    rng = np.random.default_rng(4)
    m = random_model(rng, 6, 3, [7, 7, 6, 6, 5, 5])
    codes = random_codes(rng, m, 3000)
    Q = rng.standard_normal((4, m.D)).astype(np.float32)
    ix = make_index(m, codes=codes)
    lab, dis = ix.search(Q, 5, EA | 0x100)
    lut = port.create_lut(m, Q)
    for q in range(4):
        d = np.zeros(3000, np.float32)
        g1 = np.zeros(3000, np.float32)
        for s in range(4):
            g1 = g1 + lut[q, m.lut_off[s] + codes[:, s]]
        g2 = np.zeros(3000, np.float32)
        for s in range(4, 6):
            g2 = g2 + lut[q, m.lut_off[s] + codes[:, s]]
        d = (d + g1) + g2
        order = np.lexsort((np.arange(3000), d))[:5]
        assert np.array_equal(lab[q], order)
        assert bitwise_equal(dis[q], d[order])
    ix.close()


def test_shard_invariance_and_key_merge(port):
    """Rule 7: splitting the rows into G shards (id_base per shard) + key merge == one index."""
    import torch
    from vaq_b200.index import EA, PROJECTED
    rng = np.random.default_rng(17)
    m = random_model(rng, 8, 4, [9, 8, 8, 7, 7, 6, 6, 5])
    n, k, nq = 40000, 10, 33
    codes = random_codes(rng, m, n)
    codes[n // 2:n // 2 + 2000] = codes[:2000]        # cross-shard exact ties
    Q = rng.standard_normal((nq, m.D)).astype(np.float32)
    want_lab, want_dis = port.search_lex(m, codes, Q, k)
    dq = torch.from_numpy(Q).cuda()
    for G in (1, 2, 4, 8):
        bounds = [(n * r) // G for r in range(G + 1)]
        keys = torch.empty((G, nq, k), dtype=torch.int64, device="cuda")
        shards = []
        for r in range(G):
            ix = make_index(m, codes=codes[bounds[r]:bounds[r + 1]])
            ix.set_id_base(bounds[r])
            ix.search_keys_device(dq.data_ptr(), nq, k, EA | PROJECTED, keys[r].data_ptr(), torch.cuda.current_stream().cuda_stream)
            shards.append(ix)
        lab = torch.empty((nq, k), dtype=torch.int32, device="cuda")
        dis = torch.empty((nq, k), dtype=torch.float32, device="cuda")
        shards[0].merge_keys_device(keys.data_ptr(), G, nq, k, 0, lab.data_ptr(), dis.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(lab.cpu().numpy(), want_lab), f"G={G}"
        assert bitwise_equal(dis.cpu().numpy(), want_dis), f"G={G}"
        for ix in shards:
            ix.close()


# ---- TI / visit, refine ---------------------------------------------------------------------------

def test_ti_visit_vs_reference_fixture(port):
    from vaq_b200.index import EA, PROJECTED, SQRT, TI
    g = load_golden("vaq_small_a")
    m, _ = golden_model(g)
    ix = make_index(m, codes=g["ti_codes_grouped"])
    ix.set_clusters(g["ti_clusters"], g["ti_start_idx"], g["ti_sizes"], g["ti_members"])
    for visit in (1.0, 0.25):
        ix.set_visit(visit)
        lab, dis = ix.search(g["Q"], int(g["k"]), TI | EA | PROJECTED | SQRT)
        tag = f"v{int(visit * 100)}"
        assert_knn_equiv(lab, dis, g[f"ti_lab_{tag}"], g[f"ti_dis_{tag}"], what=f"TI visit={visit} vs reference")
    ix.close()


def test_refine_vs_reference_fixture(port):
    for case in ("vaq_small_a", "vaq_small_b"):
        g = load_golden(case)
        m, _ = golden_model(g)
        X, Q = g["X"], g["Qraw"]
        ix = make_index(m)
        ix.set_raw_vectors(X)
        lab, dis = ix.refine(Q, g["lab_heap"], int(g["kr"]))
        assert_knn_equiv(lab, dis, g["refine_lab"], g["refine_dis"], rtol=1e-5, what=f"{case}/refine")
        ix.close()


# ---- errors ----------------------------------------------------------------------------------------

def test_error_paths():
    from vaq_b200 import _lib
    from vaq_b200.index import EA, TI, VAQIndex
    rng = np.random.default_rng(0)
    m = random_model(rng, 4, 2, [4, 4, 4, 4])
    ix = make_index(m, codes=random_codes(rng, m, 100))
    q = np.zeros((1, m.D), np.float32)
    with pytest.raises(_lib.VaqGpuError) as e:
        ix.search(q, 5, EA)                      # raw queries without eigenvectors
    assert e.value.code == _lib.VAQGPU_ESTATE
    with pytest.raises(_lib.VaqGpuError) as e:
        ix.search(q, 5, TI | 0x100)              # TI without clusters
    assert e.value.code == _lib.VAQGPU_ESTATE
    with pytest.raises(_lib.VaqGpuError) as e:
        ix.search(q, 0, EA | 0x100)
    assert e.value.code == _lib.VAQGPU_EINVAL
    with pytest.raises(_lib.VaqGpuError):
        ix.get_codes(90, 20)
    with pytest.raises(_lib.VaqGpuError):
        VAQIndex(2, [16, 4], [np.zeros((1 << 16, 2), np.float32), np.zeros((16, 2), np.float32)])
    lab, dis = ix.search(np.zeros((0, m.D), np.float32), 5, EA | 0x100)   # empty batch is a no-op
    assert lab.shape == (0, 5)
    ix.close()


# ---- full-size properties (BASELINE shapes the oracle cannot finish in seconds) -----------------------

def test_large_synthetic_index_planted_needles_and_slices(port):
    """8M on-device synthetic rows (C4/C5-style generator).  Size-independent checks: (1) every query's own
    code row, planted at a known id, comes back at rank 0 with the oracle's distance for that row; (2) the
    returned distances equal the oracle's distances of the returned rows (regenerated on the host from the
    counter-based generator); (3) no row of a 1M-row host slice beats the k-th returned distance without being
    returned; (4) results are identical for the three scan kernels."""
    from vaq_b200 import synth
    from vaq_b200.index import EA, PROJECTED, SCAN_F32, SCAN_V1
    rng = np.random.default_rng(2024)
    bits = [9, 9, 8, 8, 8, 7, 7, 7, 6, 6, 6, 6, 5, 5, 5, 5]
    m = random_model(rng, len(bits), 4, bits)
    n, nq, k, seed = 8_000_000, 24, 10, 4242
    ix = make_index(m)
    ix.reserve(n)
    ix.add_synthetic(n, seed)
    # queries = centroids of planted rows (+ noise) so that the planted row is (one of) the nearest
    planted = rng.integers(0, n, size=nq)
    pc = synth.synth_codes(bits, 1, 0, seed)  # shape check only
    assert pc.shape == (1, len(bits))
    Q = np.empty((nq, m.D), np.float32)
    for i, r in enumerate(planted):
        c = synth.synth_codes(bits, 1, int(r), seed)[0]
        Q[i] = np.concatenate([m.centroids[s][c[s]] for s in range(m.M)]) + rng.standard_normal(m.D).astype(np.float32) * 1e-3
    lab, dis = ix.search(Q, k, EA | PROJECTED)
    assert ix.last_config()["scan_kernel"] == 3
    lut = port.create_lut(m, Q)
    for i in range(nq):
        rows = synth.synth_codes(bits, 1, 0, seed)[:0]
        got_codes = np.concatenate([synth.synth_codes(bits, 1, int(r), seed) for r in lab[i]])
        d = port.adc_all(m, lut[i], got_codes)
        assert bitwise_equal(d, dis[i]), "returned distance != oracle distance of the returned row"
        assert (np.diff(dis[i]) >= 0).all()
        pr = synth.synth_codes(bits, 1, int(planted[i]), seed)
        dp = port.adc_all(m, lut[i], pr)[0]
        assert dp >= dis[i, 0] and (planted[i] in lab[i] or dp > dis[i, -1] or dp == dis[i, -1])
        assert lab[i, 0] == planted[i] or dis[i, 0] <= dp
    # (3) a host-regenerated slice cannot contain an unreturned row under the k-th distance
    lo = 3_000_000
    sl = synth.synth_codes(bits, 1_000_000, lo, seed)
    for i in range(0, nq, 4):
        d = port.adc_all(m, lut[i], sl)
        better = np.nonzero(d < dis[i, -1])[0] + lo
        assert set(better.tolist()) <= set(lab[i].tolist())
    # (4) kernel invariance
    for extra in (SCAN_F32, SCAN_V1):
        lab2, dis2 = ix.search(Q, k, EA | PROJECTED | extra)
        assert np.array_equal(lab2, lab) and bitwise_equal(dis2, dis)
    ix.close()


@pytest.mark.parametrize("M,bits_of", [(48, lambda s: 6 if s < 24 else 5),      # 264 bits -> 3 words per row (generic word path), M > 32 margins
                                       (40, lambda s: 6),                         # 240 bits, M > 32
                                       (64, lambda s: 5 if s < 32 else 4),        # 288 bits, M = 64
                                       (12, lambda s: 10 if s < 4 else 7)])       # four 10-bit leading fields (FAST1 boundary: 0,10,20,30)
def test_search_many_subspaces_fp16_path(port, M, bits_of):
    from vaq_b200.index import EA, PROJECTED
    rng = np.random.default_rng(M)
    bits = [bits_of(s) for s in range(M)]
    m = random_model(rng, M, 2, bits)
    codes = random_codes(rng, m, 60000)
    Q = rng.standard_normal((19, m.D)).astype(np.float32)
    ix = make_index(m, codes=codes)
    check_search(port, m, codes, Q, 10, EA, ix=ix)
    ix.search(Q, 10, EA | PROJECTED)
    assert ix.last_config()["scan_kernel"] == 3
    ix.close()


# ---- conflict-aware row order (csrc/layout.cu) ---------------------------------------------------------------------

def quarter_conflicts(codes, order, nf=4):
    """average over quarter-warps (8 consecutive storage rows) and stage-1 fields of the maximum multiplicity of
    `code mod 8` — the number of shared-memory wavefronts a 128-bit gather of that field costs."""
    res = (codes[order][:, :nf] & 7).astype(np.int64)
    n8 = (res.shape[0] // 8) * 8
    r = res[:n8].reshape(-1, 8, nf)
    onehot = (r[..., None] == np.arange(8)[None, None, None, :]).sum(1)          # [groups, nf, 8]
    return float(onehot.max(2).mean())


def test_conflict_aware_layout_is_invisible_and_effective(port, monkeypatch):
    from vaq_b200.index import EA, PROJECTED
    monkeypatch.setenv("VAQGPU_TUNE", "order=0")          # aligned windows only (the scan order has its own test below)
    rng = np.random.default_rng(31)
    bits = [9, 9, 9, 9, 8, 8, 7, 7, 7, 7, 6, 6]
    m = random_model(rng, len(bits), 2, bits)
    n1, n2 = 30000, 11111
    codes = random_codes(rng, m, n1 + n2)
    Q = rng.standard_normal((20, m.D)).astype(np.float32)
    ix = make_index(m, codes=codes[:n1])
    assert np.array_equal(ix.get_row_order(), np.arange(n1))                      # arrival order until the first search
    check_search(port, m, codes[:n1], Q, 10, EA, ix=ix)
    assert ix.last_config()["conflict_aware_layout"] == 1
    order = ix.get_row_order()
    assert np.array_equal(np.sort(order), np.arange(n1))                          # a permutation ...
    assert np.array_equal(order // 4096, np.arange(n1) // 4096)                   # ... inside windows of 4096 rows
    before, after = quarter_conflicts(codes[:n1], np.arange(n1)), quarter_conflicts(codes[:n1], order)
    assert before > 2.4 and after < 1.6, (before, after)
    assert np.array_equal(ix.get_codes(), codes[:n1])                             # codes read back in the original order
    assert np.array_equal(ix.get_codes(4000, 300), codes[4000:4300])
    # rows appended later: the partially filled window is planned again, everything stays consistent
    ix.add_codes(codes[n1:])
    check_search(port, m, codes, Q, 10, EA, ix=ix)
    order = ix.get_row_order()
    assert np.array_equal(np.sort(order), np.arange(n1 + n2)) and np.array_equal(order // 4096, np.arange(n1 + n2) // 4096)
    assert quarter_conflicts(codes, order) < 1.6
    assert np.array_equal(ix.get_codes(), codes)
    ix.close()


def test_scan_order_is_invisible(port, monkeypatch):
    """Indexes of 32 K .. 8 M rows are grouped by coarse cluster before the first search (scan order, vaqgpu_host.cu
    build_scan_order) and the query tiles start at their nearest cluster: ids, distances, codes read back, appended
    rows (tail, then re-clustering) and a later vaqgpu_set_clusters must all behave as if the rows had never moved."""
    from vaq_b200.index import EA, HEAP, PROJECTED, SQRT, TI
    rng = np.random.default_rng(77)
    bits = [9, 9, 9, 9, 8, 8, 7, 7, 7, 7, 6, 6]
    m = random_model(rng, len(bits), 2, bits)
    n1, n2, n3 = 40000, 3000, 20000
    codes = random_codes(rng, m, n1 + n2 + n3)
    codes[5000:5100] = codes[:100]                                               # ties across clusters' members
    Q = rng.standard_normal((41, m.D)).astype(np.float32)
    Q[:8] = np.concatenate([m.centroids[s][codes[:8, s]] for s in range(m.M)], 1)  # queries that sit on rows
    ix = make_index(m, codes=codes[:n1])
    check_search(port, m, codes[:n1], Q, 10, EA, ix=ix)
    order = ix.get_row_order()
    assert np.array_equal(np.sort(order), np.arange(n1))
    assert not np.array_equal(order // 4096, np.arange(n1) // 4096)              # rows crossed windows: grouped by cluster
    assert np.array_equal(ix.get_codes(), codes[:n1])
    assert np.array_equal(ix.get_codes(12345, 777), codes[12345:12345 + 777])
    check_search(port, m, codes[:n1], Q[:1], 10, HEAP, ix=ix)                    # one query, padded tile
    check_search(port, m, codes[:n1], Q[:9], 100, EA, ix=ix)                     # two tiles, k = 100
    monkeypatch.setenv("VAQGPU_TUNE", "chunks=3")                                # several row chunks: the start chunk rotates too
    check_search(port, m, codes[:n1], Q, 10, EA, ix=ix)
    monkeypatch.delenv("VAQGPU_TUNE")
    ix.add_codes(codes[n1:n1 + n2])                                              # tail after the clusters
    check_search(port, m, codes[:n1 + n2], Q, 10, EA, ix=ix)
    order2 = ix.get_row_order()
    assert np.array_equal(order2[:n1] < n1, np.ones(n1, bool)) and np.array_equal(np.sort(order2[n1:]), np.arange(n1, n1 + n2))
    assert np.array_equal(ix.get_codes(), codes[:n1 + n2])
    ix.add_codes(codes[n1 + n2:])                                                # outgrows the tail: clustered again
    check_search(port, m, codes, Q, 10, EA, ix=ix)
    order3 = ix.get_row_order()
    assert np.array_equal(np.sort(order3), np.arange(codes.shape[0])) and (order3[:n1] >= n1).any()
    assert np.array_equal(ix.get_codes(), codes)
    # TI clusters afterwards: arrival order again, exact TI answers
    C = 50
    sizes = np.full(C, codes.shape[0] // C, np.int64); sizes[-1] += codes.shape[0] - sizes.sum()
    start = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    clusters = rng.standard_normal((C, 8)).astype(np.float32)
    ix.set_clusters(clusters, start, sizes, np.arange(codes.shape[0], dtype=np.int32))
    assert np.array_equal(ix.get_row_order(), np.arange(codes.shape[0]))
    ix.set_visit(1.0)
    lab, dis = ix.search(Q, 10, TI | EA | PROJECTED)
    want_lab, want_dis = port.search_lex(m, codes, Q, 10)
    assert np.array_equal(lab, want_lab) and bitwise_equal(dis, want_dis)
    ix.close()


def test_layout_restored_for_ti_after_plain_searches(port):
    """An EA search re-orders the rows; vaqgpu_set_clusters afterwards must see the arrival order again."""
    from vaq_b200.index import EA, PROJECTED, SQRT, TI
    g = load_golden("vaq_small_a")
    m, _ = golden_model(g)
    grouped = g["ti_codes_grouped"]
    ix = make_index(m, codes=grouped)
    ix.search(g["Q"], 10, EA | PROJECTED)                                         # triggers the re-ordering
    assert ix.last_config()["conflict_aware_layout"] == 1
    ix.set_clusters(g["ti_clusters"], g["ti_start_idx"].astype(np.int64), g["ti_sizes"].astype(np.int64), g["ti_members"].astype(np.int32))
    assert np.array_equal(ix.get_row_order(), np.arange(grouped.shape[0]))
    ix.set_visit(0.25)
    lab, dis = ix.search(g["Q"], int(g["k"]), TI | EA | SQRT | PROJECTED)
    assert_knn_equiv(lab, dis, g["ti_lab_v25"], g["ti_dis_v25"], what="TI visit 0.25 after layout restore")
    ix.close()


# ---- TI / visit on the filter kernel ---------------------------------------------------------------------------------

def ti_visited_sets(Q, clusters, sizes, visit, k):
    """numpy restatement of the visiting rule (VAQ.cpp:799-827, 1548-1555, 1616-1618): clusters ranked by the sqrt of the
    sequential fp32 sum of squares over the first segdims dims (ties by index); the nearest floor(C*visit) are visited
    (all when visit >= 1), and the visit goes on while fewer than k rows were covered (empty clusters cover nothing)."""
    C, seg = clusters.shape
    out = []
    for q in Q:
        acc = np.zeros(C, np.float32)
        for j in range(seg):
            d = (np.float32(q[j]) - clusters[:, j]).astype(np.float32)
            acc = (acc + (d * d).astype(np.float32)).astype(np.float32)
        dist = np.sqrt(acc).astype(np.float32)
        order = np.lexsort((np.arange(C), dist))
        max_visit = C if visit >= 1.0 else int(np.float32(C) * np.float32(visit))
        vis, seen, enough = [], 0, False
        for pos, cl in enumerate(order):
            if not (pos < max_visit or not enough):
                break
            if sizes[cl] == 0:
                continue
            vis.append(int(cl)); seen += int(sizes[cl])
            if seen >= k:
                enough = True
        out.append(vis)
    return out


@pytest.mark.parametrize("visit,k", [(1.0, 10), (0.25, 10), (0.05, 10), (0.01, 200)])
def test_ti_on_filter_kernel_matches_exhaustive_over_visited_clusters(port, visit, k):
    from vaq_b200.index import EA, PROJECTED, SCAN_V1, SQRT, TI
    rng = np.random.default_rng(4242)
    bits = [9, 9, 9, 9, 8, 8, 7, 7, 6, 6, 5, 5]
    m = random_model(rng, len(bits), 2, bits)
    n, C, seg = 50000, 300, 8
    sizes = rng.integers(0, 400, size=C)
    sizes[[3, 4, 50, 299]] = 0                                   # empty clusters, also the last one
    sizes[[10, 11, 12]] = [1, 2, 31]                             # tiny ones (several clusters inside one 32-row tile)
    sizes = (sizes * (n / sizes.sum())).astype(np.int64)
    sizes[0] += n - sizes.sum()
    start = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    codes = random_codes(rng, m, n)
    codes[1000:1200] = codes[:200]                               # ties
    id_map = rng.permutation(n).astype(np.int32)
    clusters = rng.standard_normal((C, seg)).astype(np.float32)
    Q = rng.standard_normal((37, m.D)).astype(np.float32)
    ix = make_index(m, codes=codes)
    ix.set_clusters(clusters, start, sizes, id_map)
    ix.set_visit(visit)
    visited = ti_visited_sets(Q, clusters, sizes, visit, k)
    lut = port.create_lut(m, Q)
    for extra, kern in ((0, 3), (SCAN_V1, 1)):
        lab, dis = ix.search(Q, k, TI | EA | SQRT | PROJECTED | extra)
        assert ix.last_config()["scan_kernel"] == kern
        for q in range(Q.shape[0]):
            rows = np.concatenate([np.arange(start[c], start[c] + sizes[c]) for c in visited[q]]) if visited[q] else np.zeros(0, np.int64)
            d = port.adc_all(m, lut[q], codes[rows])
            order = np.lexsort((rows, d))[:k]
            want_lab = np.full(k, -1, np.int32); want_dis = np.full(k, np.finfo(np.float32).max, np.float32)
            want_lab[:order.size] = id_map[rows[order]]
            want_dis[:order.size] = np.sqrt(d[order]).astype(np.float32)
            assert np.array_equal(lab[q], want_lab), f"visit={visit} kernel={kern} query {q}"
            assert bitwise_equal(dis[q], want_dis)
    ix.close()
