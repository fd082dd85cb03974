"""CPU: vaq_b200/io.py against the reference's own readers/writers (utils/IO.hpp, through the compiled reference)
and against the data files the reference ships (data/siftsmall)."""
from pathlib import Path

import numpy as np
import pytest

from helpers import golden_model, load_golden, orc
from vaq_b200 import io as vio

REF_DATA = orc.REFERENCE_ROOT / "data" / "siftsmall"


def test_vecs_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    f = rng.standard_normal((37, 19)).astype(np.float32)
    i = rng.integers(-5, 10 ** 6, size=(11, 100)).astype(np.int32)
    b = rng.integers(0, 256, size=(9, 128)).astype(np.uint8)
    vio.write_fvecs(tmp_path / "a.fvecs", f); vio.write_ivecs(tmp_path / "a.ivecs", i); vio.write_bvecs(tmp_path / "a.bvecs", b)
    assert np.array_equal(vio.read_fvecs(tmp_path / "a.fvecs"), f)
    assert np.array_equal(vio.read_ivecs(tmp_path / "a.ivecs"), i)
    assert np.array_equal(vio.read_bvecs(tmp_path / "a.bvecs"), b.astype(np.float32))
    assert vio.read_fvecs(tmp_path / "a.fvecs", max_rows=5).shape == (5, 19)
    raw = np.fromfile(tmp_path / "a.fvecs", np.uint8)
    assert raw.size == 37 * (4 + 19 * 4) and raw[:4].view(np.int32)[0] == 19        # int32 dim + payload per record
    f.tofile(tmp_path / "a.bin")
    assert np.array_equal(vio.read_bin(tmp_path / "a.bin", 19), f)
    (tmp_path / "bad.fvecs").write_bytes(raw[:-3].tobytes())
    with pytest.raises(ValueError):
        vio.read_fvecs(tmp_path / "bad.fvecs")
    vio.write_knn_results(tmp_path / "r.csv", np.array([[3, 1, 2], [7, 8, 9]]))
    assert (tmp_path / "r.csv").read_text().split() == ["3,1,2", "7,8,9"]


def test_model_files_round_trip(tmp_path):
    g = load_golden("vaq_small_a")
    m, _ = golden_model(g)
    vio.save_centroids(tmp_path / "c.bin", m.centroids)
    back = vio.load_centroids(tmp_path / "c.bin")
    assert len(back) == m.M and all(np.array_equal(a, b) for a, b in zip(back, m.centroids))
    vio.save_codebook(tmp_path / "cb.bin", g["codes"])
    assert np.array_equal(vio.load_codebook(tmp_path / "cb.bin"), g["codes"])


@pytest.mark.skipif(not orc.Ref.available(), reason="compiled reference not built")
def test_files_interchange_with_the_reference(tmp_path):
    """What the reference's saveCentroids/saveCodebook write loads here, and what io.py writes the reference loads."""
    ref = orc.Ref()
    g = load_golden("vaq_small_b")
    m, _ = golden_model(g)
    ref.save_codebook(tmp_path / "ref_cb.bin", g["codes"])
    assert np.array_equal(vio.load_codebook(tmp_path / "ref_cb.bin"), g["codes"])
    vio.save_codebook(tmp_path / "my_cb.bin", g["codes"])
    assert np.array_equal(ref.load_codebook(tmp_path / "my_cb.bin"), g["codes"])
    assert (tmp_path / "ref_cb.bin").read_bytes() == (tmp_path / "my_cb.bin").read_bytes()
    ref.save_centroids(tmp_path / "ref_c.bin", m)
    mine = vio.load_centroids(tmp_path / "ref_c.bin")
    assert all(np.array_equal(a, b) for a, b in zip(mine, m.centroids))
    vio.save_centroids(tmp_path / "my_c.bin", m.centroids)
    flat, M, L = ref.load_centroids_flat(tmp_path / "my_c.bin")
    assert (M, L) == (m.M, m.L) and np.array_equal(flat, m.cent_flat)
    assert (tmp_path / "ref_c.bin").read_bytes() == (tmp_path / "my_c.bin").read_bytes()
    f = np.random.default_rng(1).standard_normal((20, 16)).astype(np.float32)
    vio.write_fvecs(tmp_path / "q.fvecs", f)
    assert np.array_equal(ref.read_fvecs(tmp_path / "q.fvecs", 16, 20), f)
    iv = np.arange(300, dtype=np.int32).reshape(3, 100)
    vio.write_ivecs(tmp_path / "g.ivecs", iv)
    assert np.array_equal(ref.read_ivecs(tmp_path / "g.ivecs", 100, 10), iv)


def test_query_fixture_is_the_shipped_file():
    from helpers import GOLDEN
    q = vio.read_fvecs(GOLDEN / "siftsmall_query.fvecs")
    assert q.shape == (100, 128) and q.min() >= 0 and q.max() <= 255
    if (REF_DATA / "siftsmall_query.fvecs").exists():
        assert (GOLDEN / "siftsmall_query.fvecs").read_bytes() == (REF_DATA / "siftsmall_query.fvecs").read_bytes()


@pytest.mark.skipif(not (REF_DATA / "siftsmall_query.fvecs").exists(), reason="reference data not mounted")
def test_shipped_siftsmall_files():
    q = vio.read_fvecs(REF_DATA / "siftsmall_query.fvecs")
    gt = vio.read_ivecs(REF_DATA / "siftsmall_groundtruth.ivecs")
    assert q.shape == (100, 128) and gt.shape == (100, 100)
    assert q.min() >= 0 and q.max() <= 255 and np.array_equal(q, np.round(q))        # SIFT descriptors
    assert gt.min() >= 0 and gt.max() <= 9999
    if orc.Ref.available():
        ref = orc.Ref()
        assert np.array_equal(ref.read_fvecs(REF_DATA / "siftsmall_query.fvecs", 128, 100), q)
        assert np.array_equal(ref.read_ivecs(REF_DATA / "siftsmall_groundtruth.ivecs", 100, 200), gt)


# ---- bit-vector CSV + createBitV (utils/IO.hpp:363-397, 681-704; BitVector.hpp:46-76) -------------------------

def test_create_bitv_known_answers():
    # the reference's own uses (test/test-distancefunction.cpp, test-bitvecengine.cpp): N <= 64 scalar, lists beyond
    assert vio.create_bitv(1, 1).tolist() == [1]
    assert vio.create_bitv(32, 0xDEADBEEF).tolist() == [0xDEADBEEF]
    assert vio.create_bitv(64, 0xFFFFFFFFFFFFFFFF).tolist() == [0xFFFFFFFFFFFFFFFF]
    assert vio.create_bitv(256, [1, 2, 3, 4]).tolist() == [1, 2, 3, 4]
    with pytest.raises(ValueError):
        vio.create_bitv(256, [1, 2, 3])
    assert vio.actual_bitv_len(1) == 1 and vio.actual_bitv_len(64) == 1 and vio.actual_bitv_len(65) == 2


def test_bitvector_csv_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    for nbits in (64, 128, 256):
        bv = rng.integers(0, 2 ** 63, size=(17, nbits // 64), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(17, nbits // 64), dtype=np.uint64)
        vio.write_bitvectors_csv(tmp_path / f"b{nbits}.csv", bv, nbits)
        first = (tmp_path / f"b{nbits}.csv").read_text().splitlines()[0].split(",")
        assert len(first) == nbits and first[0] == str(int(bv[0, 0]) >> 63)      # MSB first
        assert np.array_equal(vio.read_bitvectors_csv(tmp_path / f"b{nbits}.csv", nbits), bv)


@pytest.mark.skipif(not orc.Ref.available(), reason="compiled reference not built")
@pytest.mark.parametrize("nbits", [1, 31, 64, 65, 100, 128, 200, 256])
def test_bitvector_csv_interchange_with_the_reference(tmp_path, nbits):
    """Files written by either side are byte-identical, and both readers return the same words — including the
    reference reader's behaviour on a partial last word (see vaq_b200/io.py::read_bitvectors_csv)."""
    ref = orc.Ref()
    rng = np.random.default_rng(nbits)
    w = (nbits + 63) // 64
    bv = rng.integers(0, 2 ** 63, size=(9, w), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(9, w), dtype=np.uint64)
    rem = nbits - (w - 1) * 64
    if rem < 64:        # writeToExternal prints the TOP `rem` bits of the last word
        bv[:, -1] &= ~np.uint64((1 << (64 - rem)) - 1)
    ref.write_bitv_csv(tmp_path / "ref.csv", bv, nbits)
    vio.write_bitvectors_csv(tmp_path / "mine.csv", bv, nbits)
    assert (tmp_path / "ref.csv").read_bytes() == (tmp_path / "mine.csv").read_bytes()
    got_ref = ref.read_bitv_csv(tmp_path / "mine.csv", nbits, 64)
    got_mine = vio.read_bitvectors_csv(tmp_path / "ref.csv", nbits)
    assert np.array_equal(got_ref, got_mine)
    if nbits % 64 == 0:
        assert np.array_equal(got_mine, bv)
    for raw in (0, 1, 0x8000000000000001, 0x123456789ABCDEF0):
        for n in (1, 17, 64):
            assert np.array_equal(ref.create_bitv(n, raw), vio.create_bitv(n, raw))
