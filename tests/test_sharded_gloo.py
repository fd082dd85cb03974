"""CPU, world_size 2 over gloo: the host side of the row-sharded search — shard bounds, id_base, the
all-gather of shard-local top-k key lists and its ordering.  Each rank computes its shard's local top-k
with the oracle (the checker stands in for the device scan here); the gathered lists, merged by the
lexicographic rule, must equal the single-index result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT, orc


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vaq_b200.sharded import allgather_keys, make_keys_f32, make_keys_u32, shard_bounds, split_keys
    rng = np.random.default_rng(1)
    bits = np.array([6, 5, 5, 4], np.int32)
    cents = [rng.standard_normal((1 << b, 2)).astype(np.float32) for b in bits]
    m = orc.Model(2, bits, cents)
    n, nq, k = 5001, 9, 10
    codes = np.stack([rng.integers(0, 1 << int(b), size=n) for b in bits], 1).astype(np.uint16)
    codes[3000:3100] = codes[:100]                           # ties across the shard boundary
    Q = rng.standard_normal((nq, m.D)).astype(np.float32)
    port_ = orc.Port()
    b = shard_bounds(n, world)
    lo, hi = b[rank], b[rank + 1]
    lab, dis = port_.search_lex(m, codes[lo:hi], Q, k, id_base=lo)          # shard-local answer, global ids
    keys = torch.from_numpy(make_keys_f32(dis, lab).view(np.int64).reshape(nq, k))
    allk = allgather_keys(keys)                                               # [G, nq, k]
    assert allk.shape == (world, nq, k)
    merged = np.sort(allk.numpy().view(np.uint64).transpose(1, 0, 2).reshape(nq, world * k), axis=1)[:, :k]
    ids, d = split_keys(merged)
    want_lab, want_dis = port_.search_lex(m, codes, Q, k)
    ok = bool(np.array_equal(ids, want_lab) and np.array_equal(d.view(np.uint32), want_dis.view(np.uint32)))
    # Hamming keys
    from vaq_b200 import synth
    from helpers import hamming_lex
    data = synth.random_bitvectors(999, 128, seed=4)
    q = data[[5, 700]].copy()
    li, ld = hamming_lex(data[b2(999, world, rank)[0]:b2(999, world, rank)[1]], q, 5, id_base=b2(999, world, rank)[0])
    hk = torch.from_numpy(make_keys_u32(ld, li).view(np.int64))
    allh = allgather_keys(hk)
    mh = np.sort(allh.numpy().view(np.uint64).transpose(1, 0, 2).reshape(2, world * 5), axis=1)[:, :5]
    hi_, hd_ = split_keys(mh, hamming=True)
    wi, wd = hamming_lex(data, q, 5)
    ok = ok and bool(np.array_equal(hi_, wi) and np.array_equal(hd_, wd))
    # bound exchange plumbing: every shard receives the IPC handles of the OTHER shards of its group, in shard order
    from vaq_b200.sharded import ShardedVAQ

    class StubIndex:
        device = 0

        def bounds_export(self, max_queries):
            return bytes([rank + 1]) * 64, 0

        def bounds_attach_ipc(self, handles):
            self.got = handles

    shv = object.__new__(ShardedVAQ)
    shv.rank, shv.world, shv.group, shv.R, shv.qgroups, shv.r, shv.qg = rank, world, None, world, 1, rank, 0
    shv.index = StubIndex()
    ok = ok and shv.enable_bound_exchange(100) and shv.index.got == [bytes([r + 1]) * 64 for r in range(world) if r != rank]
    out[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def b2(n, world, rank):
    from vaq_b200.sharded import shard_bounds
    b = shard_bounds(n, world)
    return b[rank], b[rank + 1]


@pytest.mark.timeout(120)
def test_sharded_merge_world2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(100)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert dict(out) == {0: True, 1: True}
