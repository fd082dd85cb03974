"""GPU parity of the Hamming scan against the reference fixtures / known answers."""
import numpy as np
import pytest

from helpers import assert_hamming_equiv, hamming_lex, load_golden
from test_oracle_golden import HAMMING_KAT, KAT32, KAT64

pytestmark = pytest.mark.gpu


def make(nbits, data=None):
    from vaq_b200.index import HammingIndex
    ix = HammingIndex(nbits)
    if data is not None:
        ix.add(data)
    return ix


def test_hamming_dist_kat():
    """test/test-distancefunction.cpp:11-63,118-132 through the scan (1-row index, k=1)."""
    for nbits, a, b, want in HAMMING_KAT:
        ix = make(nbits, np.array([[a]], np.uint64))
        idx, dist = ix.query(np.array([[b]], np.uint64), 1)
        assert idx[0, 0] == 0 and dist[0, 0] == want, (nbits, hex(a), hex(b))
        ix.close()
    ix = make(256, np.array([[0, 1, 3, 7], [1, 1, 3, 7]], np.uint64))
    idx, dist = ix.query(np.array([[1, 1, 3, 7]], np.uint64), 2)
    assert idx[0].tolist() == [1, 0] and dist[0].tolist() == [0, 1]
    ix.close()


def test_bitvecengine_query_kat():
    """test/test-bitvecengine.cpp:64-79, 165-179, 246-260 (returned id order incl. ties)."""
    d1 = np.array([[1], [1], [1], [0], [0]], np.uint64)
    for nbits, data, want in ((1, d1, [0, 1, 2]), (32, KAT32, [1, 3, 4]), (64, KAT64, [1, 3, 2])):
        ix = make(nbits, data)
        idx, _ = ix.query(data[1:2], 3)
        assert idx[0].tolist() == want, nbits
        idx, _ = ix.query(data[0:1], 1)
        assert idx[0].tolist() == [0]
        ix.close()


@pytest.mark.parametrize("nbits", [256, 64, 100, 512])
def test_query_vs_reference_fixture(nbits):
    g = load_golden("hamming")
    tag = f"b{nbits}"
    data, q, k = g[f"{tag}_data"], g[f"{tag}_q"], int(g[f"{tag}_k"])
    ix = make(nbits, data)
    idx, dist = ix.query(q, k)
    li, ld = hamming_lex(data, q, k)
    assert np.array_equal(idx, li) and np.array_equal(dist, ld), "not the canonical (distance, id) top-k"
    for mname in ("heap", "sort", "heap_ea", "par"):
        assert_hamming_equiv(idx, dist, g[f"{tag}_{mname}_idx"], g[f"{tag}_{mname}_dist"], data, q, what=mname)
    ix.close()


@pytest.mark.parametrize("nbits,n,nq,k", [(256, 1, 1, 1), (256, 5, 3, 10), (256, 33, 9, 33), (128, 100000, 17, 10),
                                          (384, 5000, 8, 100), (768, 3000, 5, 3), (1024, 2000, 4, 10), (7, 300, 3, 5)])
def test_query_ragged(nbits, n, nq, k):
    from vaq_b200 import synth
    data = synth.random_bitvectors(n, nbits, seed=nbits + n)
    q = synth.random_bitvectors(nq, nbits, seed=nbits + n + 1)
    q[0] = data[n // 2]
    ix = make(nbits)
    ix.add(data[: n // 3])
    ix.add(data[n // 3:])
    idx, dist = ix.query(q, k)
    li, ld = hamming_lex(data, q, k)
    assert np.array_equal(dist, ld)
    assert np.array_equal(idx, li)
    if n < k:
        assert (idx[:, n:] == -1).all() and (dist[:, n:] == 0xFFFFFFFF).all()
    ix.close()


def test_synthetic_rows_and_shard_merge():
    import torch
    from vaq_b200 import synth
    nbits, n, nq, k, seed = 256, 60000, 21, 10, 77
    data = synth.synth_bitvectors(n, 0, nbits, seed)
    q = synth.synth_bitvectors(nq, 10 ** 9, nbits, seed)
    q[:4] = data[[5, 30000, 59999, 31]]
    li, ld = hamming_lex(data, q, k)
    dq = torch.from_numpy(q.view(np.int64)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    for G in (1, 2, 4):
        bounds = [(n * r) // G for r in range(G + 1)]
        keys = torch.empty((G, nq, k), dtype=torch.int64, device="cuda")
        shards = []
        for r in range(G):
            ix = make(nbits)
            ix.set_id_base(bounds[r])
            ix.add_synthetic(bounds[r + 1] - bounds[r], seed)     # rows regenerate from global ids
            ix.query_keys_device(dq.data_ptr(), nq, k, keys[r].data_ptr(), st)
            shards.append(ix)
        idx = torch.empty((nq, k), dtype=torch.int32, device="cuda")
        dist = torch.empty((nq, k), dtype=torch.int32, device="cuda")
        shards[0].merge_keys_device(keys.data_ptr(), G, nq, k, idx.data_ptr(), dist.data_ptr(), st)
        torch.cuda.synchronize()
        assert np.array_equal(idx.cpu().numpy(), li), f"G={G}"
        assert np.array_equal(dist.cpu().numpy().view(np.uint32), ld), f"G={G}"
        for ix in shards:
            ix.close()


def test_large_synthetic_bitvectors_planted_needles():
    """10M on-device synthetic 256-bit rows (C5-style generator): every query is a row of the index with a few
    bits flipped -> that row must come back first with exactly that distance; returned distances equal the
    popcounts of the (host-regenerated) returned rows; a host-regenerated 1M-row slice holds no unreturned row
    under the k-th distance."""
    from vaq_b200 import synth
    n, nq, k, seed, nbits = 10_000_000, 40, 10, 99, 256
    rng = np.random.default_rng(5)
    planted = rng.integers(0, n, size=nq)
    q = np.concatenate([synth.synth_bitvectors(1, int(r), nbits, seed) for r in planted])
    flips = rng.integers(0, 6, size=nq)
    for i in range(nq):
        for b in rng.choice(nbits, size=int(flips[i]), replace=False):
            q[i, b // 64] ^= np.uint64(1) << np.uint64(b % 64)
    ix = make(nbits)
    ix.add_synthetic(n, seed)
    idx, dist = ix.query(q, k)
    assert np.array_equal(idx[:, 0], planted) and np.array_equal(dist[:, 0], flips.astype(np.uint32))
    assert (np.diff(dist.astype(np.int64), axis=1) >= 0).all()
    for i in range(0, nq, 8):
        rows = np.concatenate([synth.synth_bitvectors(1, int(r), nbits, seed) for r in idx[i]])
        x = rows ^ q[i][None, :]
        pc = np.unpackbits(x.view(np.uint8).reshape(k, -1), axis=1).sum(1)
        assert np.array_equal(pc, dist[i])
    lo = 6_000_000
    sl = synth.synth_bitvectors(1_000_000, lo, nbits, seed)
    for i in range(0, nq, 8):
        x = sl ^ q[i][None, :]
        d = np.unpackbits(x.view(np.uint8).reshape(sl.shape[0], -1), axis=1).sum(1, dtype=np.int64)
        better = np.nonzero(d < int(dist[i, -1]))[0] + lo
        assert set(better.tolist()) <= set(idx[i].tolist())
    ix.close()
