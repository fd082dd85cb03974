"""GPU: row-sharded search — shard invariance of the merged answer, the peer-memory bound exchange, the single-process
multi-device handles (vaqgpu_sharded_*, hamgpu_sharded_*), and, on boxes with >= 2 GPUs, handles on two devices of one
process and the one-process-per-GPU NCCL path (SURVEY 8e, Appendix B rule 7)."""
import os
import socket
import sys

import numpy as np
import pytest

from helpers import ROOT, bitwise_equal, hamming_lex, orc

pytestmark = pytest.mark.gpu


def n_devices():
    from vaq_b200 import _lib
    return _lib.device_count()


def make_problem(seed=11, n=70001, nq=37, bits=(9, 9, 8, 8, 7, 7, 6, 6, 5, 5, 5, 5), L=2):
    rng = np.random.default_rng(seed)
    cents = [rng.standard_normal((1 << b, L)).astype(np.float32) * (1.0 + 3.0 / (1 + s)) for s, b in enumerate(bits)]
    m = orc.Model(L, np.asarray(bits, np.int32), cents)
    codes = np.stack([rng.integers(0, 1 << int(b), size=n) for b in m.bits], 1).astype(np.uint16)
    codes[n // 2:n // 2 + 200] = codes[:200]                     # ties across shard boundaries
    Q = rng.standard_normal((nq, m.D)).astype(np.float32)
    return m, codes, Q


def test_bound_exchange_between_two_shards_is_exact():
    """Two shards (both on device 0) publish bounds into each other's arrays; merged answers stay bit-identical to the
    unsharded index over several rounds with different query batches (the two halves of the bound arrays alternate)."""
    import torch
    from vaq_b200.index import EA, PROJECTED, VAQIndex
    m, codes, _ = make_problem()
    n = codes.shape[0]
    port = orc.Port()
    half = (n + 1) // 2
    shards = []
    for r, (lo, hi) in enumerate(((0, half), (half, n))):
        ix = VAQIndex(m.L, m.bits, m.centroids)
        ix.set_id_base(lo)
        ix.add_codes(codes[lo:hi])
        shards.append(ix)
    ptrs = [ix.bounds_export(64)[1] for ix in shards]
    shards[0].bounds_attach_ptr([ptrs[1]])
    shards[1].bounds_attach_ptr([ptrs[0]])
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    k = 10
    for rnd in range(5):
        rng = np.random.default_rng(100 + rnd)
        nq = (37, 64, 5, 1, 50)[rnd]
        Q = rng.standard_normal((nq, m.D)).astype(np.float32)
        dq = torch.from_numpy(Q).to(dev)
        keys = torch.empty((2, nq, k), dtype=torch.int64, device=dev)
        for r, ix in enumerate(shards):
            ix.search_keys_device(dq.data_ptr(), nq, k, EA | PROJECTED, keys[r].data_ptr(), st)
        lab = torch.empty((nq, k), dtype=torch.int32, device=dev)
        dis = torch.empty((nq, k), dtype=torch.float32, device=dev)
        shards[0].merge_keys_device(keys.data_ptr(), 2, nq, k, 0, lab.data_ptr(), dis.data_ptr(), st)
        torch.cuda.synchronize()
        want_lab, want_dis = port.search_lex(m, codes, Q, k)
        assert np.array_equal(lab.cpu().numpy(), want_lab), f"round {rnd}"
        assert bitwise_equal(dis.cpu().numpy(), want_dis), f"round {rnd}"
    for ix in shards:
        ix.close()


def test_ti_split_at_shard_boundaries_matches_one_index():
    """TI / visit with the cluster ranges split at the shard boundaries: three shards (all on device 0),
    local ranges + whole-index sizes for the visiting rule, merged answer == the unsharded TI search, bit for bit."""
    import torch
    from vaq_b200.index import EA, PROJECTED, SQRT, TI, VAQIndex
    from vaq_b200.sharded import local_cluster_ranges, shard_bounds
    m, codes, Q = make_problem(seed=77, n=40000, nq=21)
    n = codes.shape[0]
    rng = np.random.default_rng(7)
    C, seg = 90, 6
    sizes = rng.integers(0, 900, size=C).astype(np.int64)
    sizes[[5, 6, 89]] = 0
    sizes = (sizes * (n / sizes.sum())).astype(np.int64)
    sizes[0] += n - sizes.sum()
    start = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    id_map = rng.permutation(n).astype(np.int32)
    clusters = rng.standard_normal((C, seg)).astype(np.float32)
    flags = TI | EA | SQRT | PROJECTED
    one = VAQIndex(m.L, m.bits, m.centroids)
    one.add_codes(codes)
    one.set_clusters(clusters, start, sizes, id_map)
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    G = 3
    b = shard_bounds(n, G)
    shards = []
    for r in range(G):
        ix = VAQIndex(m.L, m.bits, m.centroids)
        ix.add_codes(codes[b[r]:b[r + 1]])
        ls, lz = local_cluster_ranges(start, sizes, b[r], b[r + 1])
        assert lz.sum() == b[r + 1] - b[r]
        ix.set_clusters(clusters, ls, lz, id_map[b[r]:b[r + 1]])
        ix.set_cluster_rule_sizes(sizes)
        shards.append(ix)
    dq = torch.from_numpy(Q).to(dev)
    for visit, k in ((0.25, 10), (0.02, 300), (1.0, 5)):
        one.set_visit(visit)
        want_lab, want_dis = one.search(Q, k, flags)
        keys = torch.empty((G, Q.shape[0], k), dtype=torch.int64, device=dev)
        for r, ix in enumerate(shards):
            ix.set_visit(visit)
            ix.search_keys_device(dq.data_ptr(), Q.shape[0], k, flags, keys[r].data_ptr(), st)
        lab = torch.empty((Q.shape[0], k), dtype=torch.int32, device=dev)
        dis = torch.empty((Q.shape[0], k), dtype=torch.float32, device=dev)
        shards[0].merge_keys_device(keys.data_ptr(), G, Q.shape[0], k, flags, lab.data_ptr(), dis.data_ptr(), st)
        torch.cuda.synchronize()
        # ties: the unsharded search orders equal distances by grouped row position, the merge by original id —
        # compare as sets per distance group
        got_l, got_d = lab.cpu().numpy(), dis.cpu().numpy()
        assert bitwise_equal(got_d, want_dis), f"visit={visit}"
        for q in range(Q.shape[0]):
            if not np.array_equal(got_l[q], want_lab[q]):
                for d in np.unique(want_dis[q]):
                    sel = want_dis[q] == d
                    if d != want_dis[q, -1]:
                        assert set(got_l[q][sel].tolist()) == set(want_lab[q][sel].tolist()), (visit, q)
    for ix in shards + [one]:
        ix.close()


@pytest.mark.parametrize("G", [1, 2, 4, 8])
def test_single_process_sharded_handles(G):
    from vaq_b200.index import EA, HEAP, PROJECTED, HammingShardedIndex, VAQShardedIndex
    from vaq_b200 import synth
    if G > n_devices():
        pytest.skip(f"needs {G} GPUs")
    m, codes, Q = make_problem(seed=G)
    n = codes.shape[0]
    port = orc.Port()
    sh = VAQShardedIndex(m.L, m.bits, m.centroids, n, n_gpus=G)
    sh.add_codes(codes[:1234])                                 # appended in pieces that straddle shard boundaries
    sh.add_codes(codes[1234:50000])
    sh.add_codes(codes[50000:])
    want_lab, want_dis = port.search_lex(m, codes, Q, 10)
    for flags in (EA, HEAP):
        for _ in range(3):                                     # repeated searches alternate the bound-array halves
            lab, dis = sh.search(Q, 10, flags | PROJECTED)
            assert np.array_equal(lab, want_lab) and bitwise_equal(dis, want_dis)
    Q2 = np.random.default_rng(5).standard_normal((3, m.D)).astype(np.float32)
    lab, dis = sh.search(Q2, 10, EA | PROJECTED)
    w2l, w2d = port.search_lex(m, codes, Q2, 10)
    assert np.array_equal(lab, w2l) and bitwise_equal(dis, w2d)
    sh.close()
    # synthetic rows are a function of their global id: the sharded generator equals the host generator
    sh = VAQShardedIndex(m.L, m.bits, m.centroids, 40000, n_gpus=G)
    sh.add_synthetic(40000, 77)
    hc = synth.synth_codes(m.bits, 40000, 0, 77)
    lab, dis = sh.search(Q, 10, EA | PROJECTED)
    wl, wd = port.search_lex(m, hc, Q, 10)
    assert np.array_equal(lab, wl) and bitwise_equal(dis, wd)
    sh.close()
    # Hamming
    data = synth.random_bitvectors(30011, 256, seed=G)
    hq = data[[5, 700, 29999]].copy()
    hq[:, 2] ^= np.uint64(0xFF00FF)
    hs = HammingShardedIndex(256, data.shape[0], n_gpus=G)
    hs.add(data[:999])
    hs.add(data[999:])
    idx, dist = hs.query(hq, 7)
    wi, wd = hamming_lex(data, hq, 7)
    assert np.array_equal(idx, wi) and np.array_equal(dist, wd)
    hs.close()


def test_indexes_on_two_devices_of_one_process():
    """Opt-in shared memory is a per-device function attribute: a second handle on another GPU must work."""
    from vaq_b200.index import EA, PROJECTED, HammingIndex, VAQIndex
    from vaq_b200 import synth
    if n_devices() < 2:
        pytest.skip("needs 2 GPUs")
    m, codes, Q = make_problem(seed=3, n=20000)
    port = orc.Port()
    want = port.search_lex(m, codes, Q, 10)
    for dev in (0, 1, 0):
        ix = VAQIndex(m.L, m.bits, m.centroids, device=dev)
        ix.add_codes(codes)
        lab, dis = ix.search(Q, 10, EA | PROJECTED)
        assert np.array_equal(lab, want[0]) and bitwise_equal(dis, want[1])
        ix.close()
        data = synth.random_bitvectors(5000, 256, seed=dev)
        hx = HammingIndex(256, device=dev)
        hx.add(data)
        idx, dist = hx.query(data[:4], 300)                     # k = 300: > 48 KB of lists
        wi, wd = hamming_lex(data, data[:4], 300)
        assert np.array_equal(idx, wi) and np.array_equal(dist, wd)
        hx.close()


# ---- one process per GPU over NCCL (torch.distributed), world size 2 -----------------------------------------------

def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _nccl_worker(rank, world, port, out):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from vaq_b200 import synth
    from vaq_b200.index import EA, PROJECTED
    from vaq_b200.sharded import ShardedHamming, ShardedVAQ
    m, codes, Q = make_problem(seed=21, n=90001, nq=64)
    sh = ShardedVAQ(m.L, m.bits, m.centroids, None, codes.shape[0], rank, world, rank)
    sh.add_codes_global(codes)
    ok = sh.enable_bound_exchange(Q.shape[0])
    port_ = orc.Port()
    want_lab, want_dis = port_.search_lex(m, codes, Q, 10)
    dq = torch.from_numpy(Q).to(dev)
    for _ in range(3):
        lab, dis = sh.search(dq, 10, EA | PROJECTED)
        ok = ok and np.array_equal(lab.cpu().numpy(), want_lab) and bitwise_equal(dis.cpu().numpy(), want_dis)
    data = synth.random_bitvectors(50001, 256, seed=9)
    hq = data[[1, 40000]].copy()
    hs = ShardedHamming(256, data.shape[0], rank, world, rank)
    hs.add_global(data)
    idx, hd = hs.query(torch.from_numpy(hq.view(np.int64)).to(dev), 9)
    wi, wd = hamming_lex(data, hq, 9)
    ok = ok and np.array_equal(idx.cpu().numpy(), wi) and np.array_equal(hd.cpu().numpy().view(np.uint32), wd)
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_row_sharded_world2_nccl():
    import torch.multiprocessing as mp
    if n_devices() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(280)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert dict(out) == {0: True, 1: True}
