"""GPU: the reference-facing classes (vaq_b200.vaq.VAQ / BitVecEngine) driven the way
examples/demo_vaq.cpp:56-345 drives the reference, checked against the compiled reference when its
library travelled with the repo (oracle/_ref), else against the oracle port."""
import numpy as np
import pytest

from helpers import assert_knn_equiv, bitwise_equal, hamming_lex, orc
from vaq_b200 import synth

pytestmark = pytest.mark.gpu


def reference_search(model, codes, Q, k, mode):
    om = orc.Model(model.L, model.bits, model.centroids)
    if orc.Ref.available():
        rv = orc.Ref().vaq(om, orc.NN_EA if mode == "EA" else orc.NN_HEAP)
        rv.set_codes(codes)
        out = rv.search(Q, k)
        rv.close()
        return out
    return orc.Port().search(om, codes, Q, k, mode)


@pytest.mark.parametrize("method", ["VAQ128m16min6max10var1,HEAP", "VAQ128m16min6max10var1,EA", "VAQ96m32min2max8var1,EA"])
def test_demo_vaq_flow_siftsmall_shape(method):
    """BASELINE config C1: siftsmall shape (10K x 128 base, the 100 shipped queries, k=100).  The reference mount
    lacks siftsmall_base.fvecs (.MISSING_LARGE_BLOBS), so the base is a seeded SIFT-like set; the queries are the
    reference's own data/siftsmall/siftsmall_query.fvecs (kept as a fixture: tests/golden/siftsmall_query.fvecs)."""
    from helpers import GOLDEN
    from vaq_b200 import io as vio
    from vaq_b200.vaq import VAQ
    X = synth.sift_like(10000, 128, seed=1)
    Qraw = vio.read_fvecs(GOLDEN / "siftsmall_query.fvecs")
    assert Qraw.shape == (100, 128)
    vaq = VAQ()
    vaq.parseMethodString(method)
    XP = vaq.train(X, kmeans_iters=4)
    vaq.encode(XP)
    codes = vaq.mCodebook
    om = orc.Model(vaq.model.L, vaq.model.bits, vaq.model.centroids)
    assert np.array_equal(codes, orc.Port().encode(om, XP))
    Q = vaq.model.project(Qraw)
    ans = vaq.search(Q, 100, projected=True)
    mode = "EA" if "EA" in method else "HEAP"
    rlab, rdis = reference_search(vaq.model, codes, Q, 100, mode)
    assert_knn_equiv(ans.labels, ans.distances, rlab, rdis, what=method)
    gt = synth.brute_force_knn(X, Qraw, 100)
    assert synth.recall_at_k(ans.labels, gt, 100) == pytest.approx(synth.recall_at_k(rlab, gt, 100), abs=2e-4)
    # raw queries -> projection on the device (tolerance-only)
    ans2 = vaq.search(Qraw, 100)
    np.testing.assert_allclose(ans2.distances, ans.distances, rtol=5e-4)
    # refine: re-rank the 100 candidates exactly, keep 10 (demo_vaq.cpp:342-345)
    ref10 = vaq.refine(Qraw, ans, X, 10)
    want_lab, want_dis = orc.Port().refine(Qraw, ans.labels, X, 10)
    assert_knn_equiv(ref10.labels, ref10.distances, want_lab, want_dis, what="refine")
    assert synth.recall_at_k(ref10.labels, gt, 10) >= synth.recall_at_k(ans.labels[:, :10], gt, 10)


def test_ti_visit_flow():
    from vaq_b200.vaq import VAQ
    X = synth.decaying_gaussian(20000, 64, seed=5)
    Qraw = synth.decaying_gaussian(40, 64, seed=6)
    vaq = VAQ()
    vaq.parseMethodString("VAQ128m16min6max10var1,EA_TI64")
    XP = vaq.train(X, kmeans_iters=3)
    vaq.encode(XP)
    codes = vaq.mCodebook
    vaq.clusterTI(True)
    Q = vaq.model.project(Qraw)
    om = orc.Model(vaq.model.L, vaq.model.bits, vaq.model.centroids)
    # visit = 1: every cluster scanned -> same neighbours as the exhaustive scan (distances sqrt, original ids)
    vaq.mVisit = 1.0
    full = vaq.search(Q, 10, projected=True)
    lab, dis = orc.Port().search_lex(om, codes, Q, 10)
    assert np.array_equal(full.labels, lab)
    np.testing.assert_allclose(full.distances, np.sqrt(dis), rtol=1e-6)
    # visit = 0.25: top-k over the rows of the nearest quarter of the clusters
    vaq.mVisit = 0.25
    part = vaq.search(Q, 10, projected=True)
    ti = vaq.ti_state
    cd = np.sqrt(((Q[:, None, :ti["clusters"].shape[1]] - ti["clusters"][None]) ** 2).sum(-1))
    for q in range(Q.shape[0]):
        vis = np.argsort(cd[q], kind="stable")[:16]
        rows = np.concatenate([ti["members"][ti["start"][c]:ti["start"][c] + ti["sizes"][c]] for c in vis])
        lut = orc.Port().create_lut(om, Q[q:q + 1])[0]
        d = orc.Port().adc_all(om, lut, codes[rows])
        order = np.lexsort((rows, d))[:10]
        assert set(part.labels[q].tolist()) == set(rows[order].tolist()), q


def test_device_cluster_ti_is_a_stable_nearest_centre_partition():
    """vaqgpu_cluster_ti (VAQ::clusterTI on the device): every row sits in the cluster of its nearest centre (distance
    between the row decoded in the leading segments and the centre), rows keep their order inside a cluster, the
    ranges partition the rows, two runs give the same bits, and the regrouped matrix holds the same rows."""
    from vaq_b200.index import VAQIndex
    rng = np.random.default_rng(12)
    bits = np.array([8, 8, 7, 7, 6, 6, 5, 5], np.int32)
    L, seg, C = 3, 4, 37
    cents = [rng.standard_normal((1 << int(b), L)).astype(np.float32) for b in bits]
    n = 30011
    codes = np.stack([rng.integers(0, 1 << int(b), size=n) for b in bits], 1).astype(np.uint16)
    outs = []
    for _ in range(2):
        ix = VAQIndex(L, bits, cents)
        ix.add_codes(codes)
        ix.cluster_ti(C, seg, 5)
        ti = ix.get_clusters()
        outs.append((ti, ix.get_codes()))
        ix.close()
    (ti, grouped), (ti2, grouped2) = outs
    for key in ("clusters", "start", "sizes", "members"):
        assert np.array_equal(ti[key], ti2[key]), f"{key} differs between two runs"
    assert np.array_equal(grouped, grouped2)
    assert ti["clusters"].shape == (C, seg * L)
    assert ti["sizes"].sum() == n and ti["start"][0] == 0 and np.array_equal(ti["start"][1:], np.cumsum(ti["sizes"])[:-1])
    assert np.array_equal(np.sort(ti["members"]), np.arange(n))
    assert np.array_equal(grouped, codes[ti["members"]])                       # the rows moved with their ids
    dec = np.concatenate([cents[s][codes[:, s]] for s in range(seg)], axis=1).astype(np.float64)
    d = ((dec[:, None, :] - ti["clusters"][None].astype(np.float64)) ** 2).sum(-1)
    for c in range(C):
        mem = ti["members"][ti["start"][c]:ti["start"][c] + ti["sizes"][c]]
        assert (np.diff(mem) > 0).all()                                        # stable: original order inside a cluster
        assert (d[mem, c] <= d[mem].min(1) * (1 + 1e-5) + 1e-6).all()          # nearest centre (fp32 vs fp64 slack)


def test_bitvecengine_flow():
    from vaq_b200.vaq import BitVecEngine
    data = synth.random_bitvectors(20000, 256, seed=9)
    q = data[100:108].copy()
    q[:, 2] ^= np.uint64(0xFF)
    e = BitVecEngine(256)
    e.loadBitV(data[:5000])
    e.appendBitV(data[5000:])
    assert e.size == 20000
    idx, dist = e.query(q, 10, BitVecEngine.Sort)
    li, ld = hamming_lex(data, q, 10)
    assert np.array_equal(idx, li) and np.array_equal(dist, ld)
    idx2, dist2 = e.queryParallel(q, 10, 4)
    assert np.array_equal(idx2, idx) and np.array_equal(dist2, dist)
