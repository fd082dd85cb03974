"""Row-sharded search over the GPUs of one box: one process per GPU (torch.distributed),
contiguous row blocks, local top-k per shard as sortable 64-bit keys, one all-gather over
NCCL/NVLink, device-side merge (the reference's precedent for merging partial
answers is BitVecEngine.cpp:1599-1611).

The result is the k smallest (distance, id) keys overall, hence independent of the number of
shards (bit-identical ids for G = 1, 2, 4, 8).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_rows: int, world: int) -> list[int]:
    """Row block of rank r is [bounds[r], bounds[r+1]): ceil(n/G)-sized blocks, last one short."""
    per = -(-int(n_rows) // int(world))
    return [min(r * per, n_rows) for r in range(world + 1)]


def local_cluster_ranges(start, sizes, lo: int, hi: int):
    """Intersection of the cluster row ranges [start, start+size) with the shard's rows [lo, hi), relative to lo."""
    start = np.ascontiguousarray(start, np.int64)
    sizes = np.ascontiguousarray(sizes, np.int64)
    a = np.clip(start, lo, hi)
    b = np.clip(start + sizes, lo, hi)
    return (a - lo).astype(np.int64), (b - a).astype(np.int64)


def make_keys_f32(dist: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """(float32 distance bits << 32) | uint32 id — the key format of the scan kernels (host restatement
    used by tests and by callers that post-process gathered keys)."""
    d = np.ascontiguousarray(dist, np.float32).view(np.uint32).astype(np.uint64)
    return (d << np.uint64(32)) | np.ascontiguousarray(ids).astype(np.uint32).astype(np.uint64)


def make_keys_u32(dist: np.ndarray, ids: np.ndarray) -> np.ndarray:
    d = np.ascontiguousarray(dist).astype(np.uint32).astype(np.uint64)
    return (d << np.uint64(32)) | np.ascontiguousarray(ids).astype(np.uint32).astype(np.uint64)


def split_keys(keys: np.ndarray, hamming: bool = False):
    keys = np.ascontiguousarray(keys, np.uint64)
    ids = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.int32)
    hi = (keys >> np.uint64(32)).astype(np.uint32)
    return ids, (hi if hamming else hi.view(np.float32))


def allgather_keys(local_keys, group=None):
    """[nq, k] int64 tensor per rank -> [G, nq, k] on every rank (rank order == shard order)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    local_keys = local_keys.contiguous()
    flat = torch.empty((world * local_keys.shape[0],) + tuple(local_keys.shape[1:]), dtype=local_keys.dtype,
                       device=local_keys.device)
    if world == 1:
        flat.copy_(local_keys)
    else:
        dist.all_gather_into_tensor(flat, local_keys, group=group)     # concatenation along dim 0 == [G][nq][k]
    return flat.view((world,) + tuple(local_keys.shape))


class ShardedVAQ:
    """One rank's share of a VAQ index spread over the GPUs of one box.

    ``row_shards`` = R (default: world size, i.e. pure row sharding).  The ranks form ``world // R`` replica
    groups of R ranks: inside a group the code matrix is row-sharded (rank ``r = rank % R`` holds block r of R),
    and each group answers its own contiguous slice of the query batch.  All ranks hold the same model and
    receive the same query batch; one all-gather of the shard-local top-k key lists plus a device merge per
    query slice gives every rank the full answer.  R = world is the layout BASELINE.json names; R < world trades
    HBM (the matrix is replicated world/R times) for throughput when the index is small."""

    def __init__(self, L, bits, centroids, eig, n_rows_total: int, rank: int, world: int, device: int, group=None,
                 row_shards: int | None = None):
        from .index import VAQIndex
        R = world if not row_shards else int(row_shards)
        if R < 1 or world % R:
            raise ValueError(f"row_shards={R} must divide the world size {world}")
        self.rank, self.world, self.group, self.R = rank, world, group, R
        self.qgroups = world // R
        self.r, self.qg = rank % R, rank // R
        self.bounds = shard_bounds(n_rows_total, R)
        self.lo, self.hi = self.bounds[self.r], self.bounds[self.r + 1]
        self.index = VAQIndex(L, bits, centroids, eig=eig, device=device)
        self.index.set_id_base(self.lo)
        self.index.reserve(max(1, self.hi - self.lo))

    def enable_bound_exchange(self, max_queries: int, exchange=None) -> bool:
        """Lets the shards of this rank's replica group publish their running k-th-best bounds into each other's HBM
        over NVLink (vaqgpu_bounds_*; CUDA IPC handles travel through one all-gather).  Collective over the world.
        Returns False (and changes nothing) when the group has a single shard.  ``exchange(blob) -> [blob per rank]``
        overrides the all-gather (tests)."""
        if self.R < 2:
            return False
        handle, _ = self.index.bounds_export(max_queries)
        handles = (exchange or self._allgather_bytes)(handle)
        peers = [handles[self.qg * self.R + r] for r in range(self.R) if r != self.r]
        self.index.bounds_attach_ipc(peers)
        self._barrier()
        return True

    def _barrier(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.barrier(group=self.group)

    def _allgather_bytes(self, blob: bytes) -> list[bytes]:
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", self.index.device) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
        mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
        out = torch.empty(self.world * len(blob), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(out, mine, group=self.group)
        raw = out.cpu().numpy().tobytes()
        return [raw[i * len(blob):(i + 1) * len(blob)] for i in range(self.world)]

    def add_codes_global(self, codes: np.ndarray):
        """Takes the full [N, M] uint16 code matrix (or any object sliceable by rows) and keeps this rank's block."""
        self.index.add_codes(codes[self.lo:self.hi])

    def add_synthetic(self, seed: int, cdf=None):
        self.index.add_synthetic(self.hi - self.lo, seed, cdf)

    def set_clusters_global(self, clusters, start, sizes, id_map=None):
        """TI / visit on a row-sharded index: the cluster ranges of the whole (cluster-grouped) matrix are
        split at the shard boundaries — this rank keeps the intersection of every range with its rows — while the
        visiting rule keeps counting the clusters' sizes in the whole index."""
        st, sz = local_cluster_ranges(start, sizes, self.lo, self.hi)
        im = None if id_map is None else np.ascontiguousarray(id_map, np.int32)[self.lo:self.hi]
        self.index.set_clusters(clusters, st, sz, im)
        self.index.set_cluster_rule_sizes(np.ascontiguousarray(sizes, np.int64))

    def query_slice(self, nq: int) -> tuple[int, int]:
        """Queries [a, b) this rank's replica group answers (equal slices, padded at the end)."""
        per = -(-nq // self.qgroups)
        return min(self.qg * per, nq), min((self.qg + 1) * per, nq)

    def search(self, d_queries, k: int, flags: int):
        """d_queries: CUDA float32 tensor [nq, D], identical on every rank.  Returns (labels int32 [nq,k],
        dists float32 [nq,k]) CUDA tensors, identical on every rank."""
        import torch
        nq = d_queries.shape[0]
        st = torch.cuda.current_stream().cuda_stream
        per = -(-nq // self.qgroups)
        a, b = self.query_slice(nq)
        # the key lists and their gathered copy live in buffers kept between searches (no allocator traffic per search:
        # with peer mappings in place a cudaMalloc costs milliseconds); labels / dists are fresh tensors owned by the caller
        ck = (per, k, d_queries.device)
        if getattr(self, "_buf_key", None) != ck:
            self._keys = torch.empty((per, k), dtype=torch.int64, device=d_queries.device)
            self._allk = torch.empty((self.world, per, k), dtype=torch.int64, device=d_queries.device)
            self._buf_key = ck
        keys, allk = self._keys, self._allk
        keys.fill_(-1)                                              # -1 == empty key
        if b > a:
            self.index.search_keys_device(d_queries[a:b].data_ptr(), b - a, k, flags, keys.data_ptr(), st)
        if self.world == 1:
            allk.copy_(keys.view(1, per, k))
        else:
            import torch.distributed as dist
            dist.all_gather_into_tensor(allk.view(self.world * per, k), keys, group=self.group)      # [world][per][k], rank = qg * R + r
        labels = torch.empty((self.qgroups * per, k), dtype=torch.int32, device=d_queries.device)
        dists = torch.empty((self.qgroups * per, k), dtype=torch.float32, device=d_queries.device)
        for g in range(self.qgroups):
            self.index.merge_keys_device(allk[g * self.R].data_ptr(), self.R, per, k, flags, labels[g * per].data_ptr(),
                                         dists[g * per].data_ptr(), st)
        return labels[:nq], dists[:nq]


class ShardedHamming:
    def __init__(self, nbits: int, n_rows_total: int, rank: int, world: int, device: int, group=None):
        from .index import HammingIndex
        self.rank, self.world, self.group = rank, world, group
        self.bounds = shard_bounds(n_rows_total, world)
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]
        self.index = HammingIndex(nbits, device=device)
        self.index.set_id_base(self.lo)

    def add_global(self, words: np.ndarray):
        self.index.add(words[self.lo:self.hi])

    def add_synthetic(self, seed: int):
        self.index.add_synthetic(self.hi - self.lo, seed)

    def query(self, d_queries, k: int):
        import torch
        nq = d_queries.shape[0]
        st = torch.cuda.current_stream().cuda_stream
        keys = torch.empty((nq, k), dtype=torch.int64, device=d_queries.device)
        self.index.query_keys_device(d_queries.data_ptr(), nq, k, keys.data_ptr(), st)
        allk = allgather_keys(keys, self.group)
        idx = torch.empty((nq, k), dtype=torch.int32, device=d_queries.device)
        dist = torch.empty((nq, k), dtype=torch.int32, device=d_queries.device)
        self.index.merge_keys_device(allk.data_ptr(), self.world, nq, k, idx.data_ptr(), dist.data_ptr(), st)
        return idx, dist
