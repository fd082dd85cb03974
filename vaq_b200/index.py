"""Handles over the C ABI: one ``VAQIndex`` / ``HammingIndex`` per GPU (per shard).

Host-buffer methods take/return numpy arrays and include the H2D / D2H copies; the
``*_device`` methods take raw device pointers (ints, e.g. ``torch.Tensor.data_ptr()``) and a
CUDA stream handle and only enqueue work.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import EA, HEAP, PROJECTED, SCAN_F32, SCAN_V1, SQRT, TI, ModelDesc, check  # noqa: F401 (re-exported)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _vp(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class VAQIndex:
    """Trained model + packed code matrix resident on one GPU.

    model state == the reference's public members (VAQ.hpp:57-73): ``L`` (mSubsLen),
    ``bits`` (mBitsAlloc), ``centroids[s]`` (mCentroidsPerSubs, [2^bits[s], L]) and
    ``eig`` (real(mEigenVectors), optional)."""

    def __init__(self, L: int, bits, centroids, eig=None, device: int = 0):
        self.lib = _lib.load()
        self.bits = _c(bits, np.int32)
        self.M = int(self.bits.size)
        self.L = int(L)
        self.D = self.M * self.L
        cents = [_c(c, np.float32) for c in centroids]
        if len(cents) != self.M:
            raise ValueError("one centroid block per subspace expected")
        for s, c in enumerate(cents):
            if c.shape != (1 << int(self.bits[s]), self.L):
                raise ValueError(f"centroids[{s}] has shape {c.shape}, expected {(1 << int(self.bits[s]), self.L)}")
        self.K = (1 << self.bits.astype(np.int64))
        self.lut_off = np.concatenate([[0], np.cumsum(self.K)]).astype(np.int64)
        self.lut_size = int(self.lut_off[-1])
        flat = np.concatenate([c.reshape(-1) for c in cents]).astype(np.float32)
        eig_arr = None if eig is None else _c(eig, np.float32)
        if eig_arr is not None and eig_arr.shape != (self.D, self.D):
            raise ValueError(f"eig must be [{self.D},{self.D}]")
        desc = ModelDesc(self.D, self.M, self.L, self.bits.ctypes.data_as(C.POINTER(C.c_int32)),
                         flat.ctypes.data_as(C.POINTER(C.c_float)),
                         None if eig_arr is None else eig_arr.ctypes.data_as(C.POINTER(C.c_float)))
        h = C.c_void_p()
        check(self.lib.vaqgpu_create(C.byref(desc), int(device), C.byref(h)))
        self.h = h
        self.device = int(device)
        self.has_eig = eig_arr is not None

    # -- lifetime
    def close(self):
        if getattr(self, "h", None):
            self.lib.vaqgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- rows
    def set_id_base(self, id_base: int):
        check(self.lib.vaqgpu_set_id_base(self.h, int(id_base)))

    def reserve(self, n_total: int):
        check(self.lib.vaqgpu_reserve(self.h, int(n_total)))

    def add_codes(self, codes):
        codes = _c(codes, np.uint16)
        if codes.ndim != 2 or codes.shape[1] != self.M:
            raise ValueError(f"codes must be [n, {self.M}] uint16 (mCodebook)")
        check(self.lib.vaqgpu_add_codes_u16(self.h, _vp(codes), codes.shape[0]))

    def encode_add(self, x_proj):
        x = _c(x_proj, np.float32)
        if x.ndim != 2 or x.shape[1] != self.D:
            raise ValueError(f"x_proj must be [n, {self.D}] float32 (already projected, SURVEY D4)")
        check(self.lib.vaqgpu_encode_add(self.h, _vp(x), x.shape[0]))

    def add_synthetic(self, n: int, seed: int, cdf=None):
        cdf_arr = None if cdf is None else _c(cdf, np.float32)
        if cdf_arr is not None and cdf_arr.size != self.lut_size:
            raise ValueError("cdf must have sum(2^bits) entries")
        check(self.lib.vaqgpu_add_codes_synthetic(self.h, int(n), C.c_uint64(seed), None if cdf_arr is None else _vp(cdf_arr)))

    @property
    def num_rows(self) -> int:
        n = C.c_int64()
        check(self.lib.vaqgpu_num_rows(self.h, C.byref(n)))
        return n.value

    @property
    def row_bytes(self) -> int:
        n = C.c_int32()
        check(self.lib.vaqgpu_row_bytes(self.h, C.byref(n)))
        return n.value

    def get_codes(self, row0: int = 0, n: int | None = None) -> np.ndarray:
        n = self.num_rows - row0 if n is None else n
        out = np.empty((n, self.M), np.uint16)
        check(self.lib.vaqgpu_get_codes_u16(self.h, int(row0), int(n), _vp(out)))
        return out

    def get_row_order(self, srow0: int = 0, n: int | None = None) -> np.ndarray:
        """original index of the rows stored at positions [srow0, srow0+n) (diagnostics; see csrc/layout.cu)"""
        n = self.num_rows - srow0 if n is None else n
        out = np.empty(n, np.uint32)
        check(self.lib.vaqgpu_get_row_order(self.h, int(srow0), int(n), _vp(out)))
        return out

    # -- query path
    def build_lut(self, q_proj) -> np.ndarray:
        q = _c(q_proj, np.float32).reshape(-1, self.D)
        out = np.empty((q.shape[0], self.lut_size), np.float32)
        check(self.lib.vaqgpu_build_lut(self.h, _vp(q), q.shape[0], _vp(out)))
        return out

    def search(self, queries, k: int, flags: int = HEAP | PROJECTED):
        q = _c(queries, np.float32).reshape(-1, self.D)
        labels = np.empty((q.shape[0], k), np.int32)
        dists = np.empty((q.shape[0], k), np.float32)
        check(self.lib.vaqgpu_search(self.h, _vp(q), q.shape[0], int(k), int(flags), _vp(labels), _vp(dists)))
        return labels, dists

    def search_into(self, q: np.ndarray, k: int, flags: int, labels: np.ndarray, dists: np.ndarray):
        """Same as ``search`` with caller-owned (e.g. pinned) buffers; no allocation."""
        check(self.lib.vaqgpu_search(self.h, _vp(q), q.shape[0], int(k), int(flags), _vp(labels), _vp(dists)))

    def search_device(self, d_queries: int, nq: int, k: int, flags: int, d_labels: int, d_dists: int, stream: int = 0):
        check(self.lib.vaqgpu_search_device(self.h, C.c_void_p(d_queries), nq, k, flags, C.c_void_p(d_labels),
                                            C.c_void_p(d_dists), C.c_void_p(stream)))

    def search_keys_device(self, d_queries: int, nq: int, k: int, flags: int, d_keys: int, stream: int = 0):
        check(self.lib.vaqgpu_search_keys_device(self.h, C.c_void_p(d_queries), nq, k, flags, C.c_void_p(d_keys),
                                                 C.c_void_p(stream)))

    def merge_keys_device(self, d_keys_in: int, G: int, nq: int, k: int, flags: int, d_labels: int, d_dists: int,
                          stream: int = 0):
        check(self.lib.vaqgpu_merge_keys_device(C.c_void_p(d_keys_in), G, nq, k, flags, C.c_void_p(d_labels),
                                                C.c_void_p(d_dists), C.c_void_p(stream)))

    # -- cross-shard bound exchange (row-sharded deployments)
    def bounds_export(self, max_queries: int) -> tuple[bytes, int]:
        """Allocates this shard's bound array; returns (64-byte CUDA IPC handle, device pointer)."""
        buf = (C.c_ubyte * 64)()
        ptr = C.c_void_p()
        check(self.lib.vaqgpu_bounds_export(self.h, int(max_queries), buf, C.byref(ptr)))
        return bytes(buf), int(ptr.value)

    def bounds_attach_ipc(self, handles: list[bytes]):
        blob = b"".join(handles)
        arr = (C.c_ubyte * max(1, len(blob))).from_buffer_copy(blob or b"\0")
        check(self.lib.vaqgpu_bounds_attach_ipc(self.h, len(handles), arr))

    def bounds_attach_ptr(self, ptrs: list[int]):
        arr = (C.c_void_p * max(1, len(ptrs)))(*ptrs)
        check(self.lib.vaqgpu_bounds_attach_ptr(self.h, len(ptrs), arr))

    # -- TI / visit, refine
    def set_clusters(self, clusters, start, size, id_map=None):
        cl = _c(clusters, np.float32)
        st = _c(start, np.int64)
        sz = _c(size, np.int64)
        im = None if id_map is None else _c(id_map, np.int32)
        check(self.lib.vaqgpu_set_clusters(self.h, _vp(cl), cl.shape[0], cl.shape[1], _vp(st), _vp(sz),
                                           None if im is None else _vp(im)))

    def cluster_ti(self, C_: int, n_segments: int = -1, iters: int = 10):
        """VAQ::clusterTI on the device: k-means over the decoded leading segments, rows regrouped by cluster."""
        check(self.lib.vaqgpu_cluster_ti(self.h, int(C_), int(n_segments), int(iters)))

    def get_clusters(self, with_id_map: bool = True) -> dict:
        nc, sd = C.c_int32(0), C.c_int32(0)
        check(self.lib.vaqgpu_get_clusters(self.h, C.byref(nc), C.byref(sd), None, None, None, None))
        cl = np.empty((nc.value, sd.value), np.float32)
        st = np.empty(nc.value, np.int64)
        sz = np.empty(nc.value, np.int64)
        im = np.empty(self.num_rows, np.int32) if with_id_map else None
        check(self.lib.vaqgpu_get_clusters(self.h, None, None, _vp(cl), _vp(st), _vp(sz), None if im is None else _vp(im)))
        return dict(clusters=cl, start=st, sizes=sz, members=im)

    def set_cluster_rule_sizes(self, sizes):
        sz = _c(sizes, np.int64)
        check(self.lib.vaqgpu_set_cluster_rule_sizes(self.h, _vp(sz)))

    def set_visit(self, visit: float):
        check(self.lib.vaqgpu_set_visit(self.h, float(visit)))

    def set_raw_vectors(self, xtrain):
        x = _c(xtrain, np.float32)
        check(self.lib.vaqgpu_set_raw_vectors(self.h, _vp(x), x.shape[0], x.shape[1]))

    def refine(self, queries, in_labels, k: int):
        q = _c(queries, np.float32)
        inl = _c(in_labels, np.int32).reshape(q.shape[0], -1)
        labels = np.empty((q.shape[0], k), np.int32)
        dists = np.empty((q.shape[0], k), np.float32)
        check(self.lib.vaqgpu_refine(self.h, _vp(q), q.shape[0], _vp(inl), inl.shape[1], int(k), _vp(labels), _vp(dists)))
        return labels, dists

    # -- introspection
    def last_timings(self) -> dict:
        ms = (C.c_float * 4)()
        check(self.lib.vaqgpu_last_timings(self.h, ms))
        return dict(project_ms=ms[0], lut_ms=ms[1], scan_ms=ms[2], merge_ms=ms[3])

    def last_config(self) -> dict:
        cfg = (C.c_int32 * 12)()
        check(self.lib.vaqgpu_last_config(self.h, cfg))
        keys = ["threads", "row_chunks", "smem_lut_floats", "spill_lut_floats", "smem_bytes", "row_words", "launches",
                "queries_per_launch", "queries_per_cta", "scan_kernel", "conflict_aware_layout", "layout_us"]
        return dict(zip(keys, list(cfg)))


class HammingIndex:
    """Bit-vector matrix resident on one GPU (reference BitVecEngine's ``data``,
    BitVecEngine.hpp:90; rows are ``bitv`` = ceil(nbits/64) uint64 words, BitVector.hpp:13)."""

    def __init__(self, nbits: int, device: int = 0):
        self.lib = _lib.load()
        self.nbits = int(nbits)
        self.w64 = (self.nbits + 63) // 64
        h = C.c_void_p()
        check(self.lib.hamgpu_create(self.nbits, int(device), C.byref(h)))
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.hamgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_id_base(self, id_base: int):
        check(self.lib.hamgpu_set_id_base(self.h, int(id_base)))

    def add(self, words):
        w = _c(words, np.uint64).reshape(-1, self.w64)
        check(self.lib.hamgpu_add(self.h, _vp(w), w.shape[0]))

    def add_synthetic(self, n: int, seed: int):
        check(self.lib.hamgpu_add_synthetic(self.h, int(n), C.c_uint64(seed)))

    @property
    def num_rows(self) -> int:
        n = C.c_int64()
        check(self.lib.hamgpu_num_rows(self.h, C.byref(n)))
        return n.value

    def query(self, queries, k: int):
        q = _c(queries, np.uint64).reshape(-1, self.w64)
        idx = np.empty((q.shape[0], k), np.int32)
        dist = np.empty((q.shape[0], k), np.uint32)
        check(self.lib.hamgpu_query(self.h, _vp(q), q.shape[0], int(k), _vp(idx), _vp(dist)))
        return idx, dist

    def query_into(self, q: np.ndarray, k: int, idx: np.ndarray, dist: np.ndarray):
        check(self.lib.hamgpu_query(self.h, _vp(q), q.shape[0], int(k), _vp(idx), _vp(dist)))

    def query_device(self, d_queries: int, nq: int, k: int, d_idx: int, d_dist: int, stream: int = 0):
        check(self.lib.hamgpu_query_device(self.h, C.c_void_p(d_queries), nq, k, C.c_void_p(d_idx), C.c_void_p(d_dist),
                                           C.c_void_p(stream)))

    def query_keys_device(self, d_queries: int, nq: int, k: int, d_keys: int, stream: int = 0):
        check(self.lib.hamgpu_query_keys_device(self.h, C.c_void_p(d_queries), nq, k, C.c_void_p(d_keys), C.c_void_p(stream)))

    def merge_keys_device(self, d_keys_in: int, G: int, nq: int, k: int, d_idx: int, d_dist: int, stream: int = 0):
        check(self.lib.hamgpu_merge_keys_device(C.c_void_p(d_keys_in), G, nq, k, C.c_void_p(d_idx), C.c_void_p(d_dist),
                                                C.c_void_p(stream)))

    def last_timings(self) -> dict:
        ms = (C.c_float * 2)()
        check(self.lib.hamgpu_last_timings(self.h, ms))
        return dict(scan_ms=ms[0], merge_ms=ms[1])

    def last_config(self) -> dict:
        cfg = (C.c_int32 * 8)()
        check(self.lib.hamgpu_last_config(self.h, cfg))
        keys = ["threads", "splits", "queries_per_cta", "_", "smem_bytes", "row_words", "launches", "queries_per_launch"]
        return dict(zip(keys, list(cfg)))


class VAQShardedIndex:
    """One host process, ``n_gpus`` devices: the row-sharded index behind ``vaqgpu_sharded_*`` (contiguous row blocks,
    peer-memory bound exchange, one ncclAllGather of the shard-local key lists, device merge)."""

    def __init__(self, L: int, bits, centroids, n_rows_total: int, eig=None, n_gpus: int = 1, dev_ids=None):
        self.lib = _lib.load()
        self.bits = _c(bits, np.int32)
        self.M, self.L = int(self.bits.size), int(L)
        self.D = self.M * self.L
        flat = np.concatenate([_c(c, np.float32).reshape(-1) for c in centroids]).astype(np.float32)
        eig_arr = None if eig is None else _c(eig, np.float32)
        desc = ModelDesc(self.D, self.M, self.L, self.bits.ctypes.data_as(C.POINTER(C.c_int32)),
                         flat.ctypes.data_as(C.POINTER(C.c_float)),
                         None if eig_arr is None else eig_arr.ctypes.data_as(C.POINTER(C.c_float)))
        ids = None if dev_ids is None else (C.c_int * n_gpus)(*[int(d) for d in dev_ids])
        h = C.c_void_p()
        check(self.lib.vaqgpu_sharded_create(C.byref(desc), int(n_gpus), ids, int(n_rows_total), C.byref(h)))
        self.h, self.n_gpus = h, int(n_gpus)

    def close(self):
        if getattr(self, "h", None):
            self.lib.vaqgpu_sharded_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_codes(self, codes):
        codes = _c(codes, np.uint16)
        check(self.lib.vaqgpu_sharded_add_codes_u16(self.h, _vp(codes), codes.shape[0]))

    def encode_add(self, x_proj):
        x = _c(x_proj, np.float32)
        check(self.lib.vaqgpu_sharded_encode_add(self.h, _vp(x), x.shape[0]))

    def add_synthetic(self, n: int, seed: int, cdf=None):
        cdf_arr = None if cdf is None else _c(cdf, np.float32)
        check(self.lib.vaqgpu_sharded_add_codes_synthetic(self.h, int(n), C.c_uint64(seed), None if cdf_arr is None else _vp(cdf_arr)))

    def search(self, queries, k: int, flags: int = EA | PROJECTED):
        q = _c(queries, np.float32).reshape(-1, self.D)
        labels = np.empty((q.shape[0], k), np.int32)
        dists = np.empty((q.shape[0], k), np.float32)
        check(self.lib.vaqgpu_sharded_search(self.h, _vp(q), q.shape[0], int(k), int(flags), _vp(labels), _vp(dists)))
        return labels, dists

    def search_into(self, q: np.ndarray, k: int, flags: int, labels: np.ndarray, dists: np.ndarray):
        check(self.lib.vaqgpu_sharded_search(self.h, _vp(q), q.shape[0], int(k), int(flags), _vp(labels), _vp(dists)))


class HammingShardedIndex:
    """One host process, ``n_gpus`` devices: ``hamgpu_sharded_*``."""

    def __init__(self, nbits: int, n_rows_total: int, n_gpus: int = 1, dev_ids=None):
        self.lib = _lib.load()
        self.nbits, self.w64 = int(nbits), (int(nbits) + 63) // 64
        ids = None if dev_ids is None else (C.c_int * n_gpus)(*[int(d) for d in dev_ids])
        h = C.c_void_p()
        check(self.lib.hamgpu_sharded_create(self.nbits, int(n_gpus), ids, int(n_rows_total), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.hamgpu_sharded_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add(self, words):
        w = _c(words, np.uint64).reshape(-1, self.w64)
        check(self.lib.hamgpu_sharded_add(self.h, _vp(w), w.shape[0]))

    def add_synthetic(self, n: int, seed: int):
        check(self.lib.hamgpu_sharded_add_synthetic(self.h, int(n), C.c_uint64(seed)))

    def query(self, queries, k: int):
        q = _c(queries, np.uint64).reshape(-1, self.w64)
        idx = np.empty((q.shape[0], k), np.int32)
        dist = np.empty((q.shape[0], k), np.uint32)
        check(self.lib.hamgpu_sharded_query(self.h, _vp(q), q.shape[0], int(k), _vp(idx), _vp(dist)))
        return idx, dist
