"""TEST INFRASTRUCTURE — ctypes/numpy front-end of the CPU oracle.

Two checkers live behind this module:

* ``Port``  — oracle/libvaq_oracle.so, the plain-C restatement (oracle/vaq_oracle.c).
  Travels to the GPU box; pinned against the reference by tests/test_oracle_golden.py
  and the committed tests/golden fixtures.
* ``Ref``   — oracle/_ref/libvaq_ref.so, the UNMODIFIED reference compiled from
  /root/reference by oracle/Makefile (present whenever build() ran in a container
  that has the reference mounted; the built .so travels with the repo snapshot).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this
module.  vaq_b200/ never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
PORT_SO = HERE / "libvaq_oracle.so"
REF_SO = HERE / "_ref" / "libvaq_ref.so"
REFERENCE_ROOT = Path(os.environ.get("VAQ_REFERENCE_ROOT", "/root/reference"))

# VAQ::NNMethod bit flags, reference bitvecengine/VAQ.hpp:38-49
NN_SORT, NN_EA, NN_TI, NN_HEAP = 0x01, 0x02, 0x04, 0x80
# BitVecEngine::QueryMethod, reference BitVecEngine.hpp:82-84
QM_HEAP, QM_SORT, QM_HEAP_EA, QM_SORT_EA = 0, 1, 2, 3


def build(ref: bool | None = None) -> None:
    """Compile the checkers (``make -C oracle``).  ``ref`` defaults to "reference present"."""
    if ref is None:
        ref = (REFERENCE_ROOT / "bitvecengine" / "VAQ.cpp").exists()
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", str(HERE), f"REF={REFERENCE_ROOT}", "-j8", *targets], check=True)


DEMO_CHECK_SRC = HERE.parent / "tests" / "cpp" / "demo_sequence_check.cpp"
DEMO_CHECK_BIN = HERE.parent / "tests" / "cpp" / "_bin" / "demo_sequence_check"


def build_demo_check() -> Path | None:
    """tests/cpp/demo_sequence_check.cpp — the reference demo's query phase (examples/demo_vaq.cpp:337-345, verbatim)
    compiled against include/vaq_gpu.hpp with the reference's own headers (VAQ.hpp, utils/Experiment.hpp, vendored
    Eigen) and linked with the compiled reference + libvaqgpu.so.  Needs the reference tree (headers are never copied):
    returns None when it is not mounted; the binary is git-ignored and travels to the GPU box like the other built files."""
    ref = REFERENCE_ROOT
    if not ((ref / "bitvecengine" / "VAQ.hpp").exists() and (ref / "external" / "eigen" / "Eigen" / "Core").exists() and REF_SO.exists()):
        return DEMO_CHECK_BIN if DEMO_CHECK_BIN.exists() else None
    root = HERE.parent
    if DEMO_CHECK_BIN.exists() and DEMO_CHECK_BIN.stat().st_mtime > max(DEMO_CHECK_SRC.stat().st_mtime, (root / "include" / "vaq_gpu.hpp").stat().st_mtime,
                                                                       (root / "include" / "vaqgpu.h").stat().st_mtime):
        return DEMO_CHECK_BIN
    DEMO_CHECK_BIN.parent.mkdir(exist_ok=True)
    subprocess.run(["g++", "-std=c++14", "-O1", "-mavx2", "-mfma", "-fopenmp", "-w", "-I", str(root / "include"), "-I", str(HERE / "stubs"),
                    "-I", str(ref / "external" / "eigen"), "-I", str(ref / "bitvecengine"), str(DEMO_CHECK_SRC), "-o", str(DEMO_CHECK_BIN),
                    "-L", str(root / "vaq_b200"), "-lvaqgpu", f"-Wl,-rpath,{root / 'vaq_b200'}",
                    "-L", str(HERE / "_ref"), "-lvaq_ref", f"-Wl,-rpath,{HERE / '_ref'}", "-Wl,--allow-multiple-definition"], check=True)
    return DEMO_CHECK_BIN


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Model:
    """Plain description of a trained VAQ model (what the reference keeps in its public
    members, VAQ.hpp:57-73): M subspaces of L dims, bits[s], centroids[s] = [2^bits[s], L]."""

    def __init__(self, L: int, bits, centroids):
        self.L = int(L)
        self.bits = np.ascontiguousarray(bits, dtype=np.int32)
        self.M = int(self.bits.size)
        self.D = self.M * self.L
        self.centroids = [_f32(c) for c in centroids]
        for s, c in enumerate(self.centroids):
            assert c.shape == (1 << int(self.bits[s]), self.L), (s, c.shape)
        self.cent_flat = np.concatenate([c.reshape(-1) for c in self.centroids]).astype(np.float32)
        self.K = (1 << self.bits.astype(np.int64)).astype(np.int64)
        self.lut_off = np.concatenate([[0], np.cumsum(self.K)]).astype(np.int32)
        self.lut_size = int(self.lut_off[-1])
        self.max_bits = int(self.bits.max())


class Port:
    """Plain-C restatement (oracle/vaq_oracle.c)."""

    def __init__(self):
        if not PORT_SO.exists():
            build(ref=False)
        self.lib = C.CDLL(str(PORT_SO))
        self.lib.orc_hamming_dist.restype = C.c_uint32

    def create_lut(self, m: Model, q_proj) -> np.ndarray:
        q = _f32(q_proj).reshape(-1, m.D)
        out = np.empty((q.shape[0], m.lut_size), np.float32)
        for i in range(q.shape[0]):
            self.lib.orc_create_lut(m.M, m.L, _ptr(m.bits, C.c_int), _ptr(m.cent_flat, C.c_float),
                                    _ptr(q[i], C.c_float), _ptr(out[i], C.c_float))
        return out

    def search(self, m: Model, codes, q_proj, k: int, mode: str = "HEAP", nthreads: int = 0):
        codes = np.ascontiguousarray(codes, dtype=np.uint16)
        q = _f32(q_proj).reshape(-1, m.D)
        nq = q.shape[0]
        labels = np.empty((nq, k), np.int32)
        dists = np.empty((nq, k), np.float32)
        self.lib.orc_search(m.M, m.L, _ptr(m.bits, C.c_int), _ptr(m.cent_flat, C.c_float),
                            _ptr(codes, C.c_uint16), C.c_long(codes.shape[0]), _ptr(q, C.c_float), nq, k,
                            1 if mode == "EA" else 0, nthreads or os.cpu_count(),
                            _ptr(labels, C.c_int), _ptr(dists, C.c_float))
        return labels, dists

    def adc_all(self, m: Model, lut_row, codes) -> np.ndarray:
        codes = np.ascontiguousarray(codes, dtype=np.uint16)
        lut_row = _f32(lut_row)
        out = np.empty(codes.shape[0], np.float32)
        self.lib.orc_adc_all(m.M, _ptr(m.lut_off, C.c_int), _ptr(lut_row, C.c_float), _ptr(codes, C.c_uint16),
                             C.c_long(codes.shape[0]), _ptr(out, C.c_float))
        return out

    def topk_lex(self, d, k: int, id_base: int = 0):
        d = _f32(d)
        ids = np.empty(k, np.int32)
        dis = np.empty(k, np.float32)
        self.lib.orc_topk_lex_f32(_ptr(d, C.c_float), C.c_long(d.size), id_base, k, _ptr(ids, C.c_int), _ptr(dis, C.c_float))
        return ids, dis

    def search_lex(self, m: Model, codes, q_proj, k: int, id_base: int = 0):
        """Canonical result: k lexicographically smallest (ADC distance, id) per query."""
        q = _f32(q_proj).reshape(-1, m.D)
        luts = self.create_lut(m, q)
        labels = np.empty((q.shape[0], k), np.int32)
        dists = np.empty((q.shape[0], k), np.float32)
        for i in range(q.shape[0]):
            labels[i], dists[i] = self.topk_lex(self.adc_all(m, luts[i], codes), k, id_base)
        return labels, dists

    def search_ti(self, m: Model, ti: dict, q_proj, k: int, visit: float = 1.0, use_ea: bool = True):
        q = _f32(q_proj).reshape(-1, m.D)
        nq = q.shape[0]
        luts = self.create_lut(m, q)
        labels = np.empty((nq, k), np.int32)
        dists = np.empty((nq, k), np.float32)
        codes = np.ascontiguousarray(ti["codes_grouped"], np.uint16)
        cl = _f32(ti["clusters"]); st = np.ascontiguousarray(ti["start_idx"], np.int32)
        sz = np.ascontiguousarray(ti["sizes"], np.int32); mem = np.ascontiguousarray(ti["members"], np.int32)
        c2c = _f32(ti["code_to_cc"])
        pruned = C.c_long(0)
        for i in range(nq):
            self.lib.orc_search_ti(m.M, _ptr(m.lut_off, C.c_int), _ptr(luts[i], C.c_float), _ptr(codes, C.c_uint16),
                                   int(cl.shape[0]), int(cl.shape[1]), _ptr(cl, C.c_float), _ptr(st, C.c_int),
                                   _ptr(sz, C.c_int), _ptr(mem, C.c_int), _ptr(c2c, C.c_float), _ptr(q[i], C.c_float),
                                   C.c_float(visit), int(use_ea), k, _ptr(labels[i], C.c_int), _ptr(dists[i], C.c_float),
                                   C.byref(pruned))
        return labels, dists

    def refine(self, queries, in_labels, xtrain, k: int):
        q = _f32(queries); x = _f32(xtrain)
        inl = np.ascontiguousarray(in_labels, np.int32).reshape(q.shape[0], -1)
        labels = np.empty((q.shape[0], k), np.int32)
        dists = np.empty((q.shape[0], k), np.float32)
        self.lib.orc_refine(_ptr(q, C.c_float), q.shape[0], q.shape[1], _ptr(inl, C.c_int), inl.shape[1],
                            _ptr(x, C.c_float), k, _ptr(labels, C.c_int), _ptr(dists, C.c_float))
        return labels, dists

    def encode(self, m: Model, x_proj, with_margin: bool = False):
        x = _f32(x_proj).reshape(-1, m.D)
        codes = np.empty((x.shape[0], m.M), np.uint16)
        margin = np.empty((x.shape[0], m.M), np.float32) if with_margin else None
        self.lib.orc_encode(m.M, m.L, _ptr(m.bits, C.c_int), _ptr(m.cent_flat, C.c_float), _ptr(x, C.c_float),
                            C.c_long(x.shape[0]), _ptr(codes, C.c_uint16),
                            _ptr(margin, C.c_float) if with_margin else None)
        return (codes, margin) if with_margin else codes

    def hamming_dist(self, a, b) -> int:
        a = np.ascontiguousarray(a, np.uint64); b = np.ascontiguousarray(b, np.uint64)
        return int(self.lib.orc_hamming_dist(_ptr(a, C.c_uint64), _ptr(b, C.c_uint64), a.size))

    def bve_query(self, data, queries, k: int, method: int = QM_SORT):
        data = np.ascontiguousarray(data, np.uint64); queries = np.ascontiguousarray(queries, np.uint64)
        n, w = data.shape
        nq = queries.shape[0]
        idx = np.empty((nq, k), np.int32)
        dist = np.empty((nq, k), np.uint32)
        self.lib.orc_bve_query(_ptr(data, C.c_uint64), C.c_long(n), w, _ptr(queries, C.c_uint64), nq, k, method,
                               _ptr(idx, C.c_int), _ptr(dist, C.c_uint32))
        return idx, dist

    def nproc(self) -> int:
        return int(self.lib.orc_nproc())


class Ref:
    """The unmodified reference, compiled (oracle/_ref/libvaq_ref.so)."""

    @staticmethod
    def available() -> bool:
        return REF_SO.exists()

    def __init__(self):
        if not REF_SO.exists():
            raise FileNotFoundError(f"{REF_SO} not built (needs /root/reference at build time)")
        self.lib = C.CDLL(str(REF_SO))
        self.lib.ref_vaq_create.restype = C.c_void_p
        self.lib.ref_bve_create.restype = C.c_void_p
        self.lib.ref_vaq_num_codes.restype = C.c_long
        self.lib.ref_hamming_dist.restype = C.c_uint32
        self.lib.ref_hamming_dist_sub.restype = C.c_uint32

    # -- VAQ
    def vaq(self, m: Model, methods: int = NN_HEAP, eig_real=None) -> "RefVAQ":
        return RefVAQ(self, m, methods, eig_real)

    # -- BitVecEngine
    def bve_query(self, nbits: int, data, queries, k: int, method: int = QM_SORT, threads: int = 0):
        data = np.ascontiguousarray(data, np.uint64); queries = np.ascontiguousarray(queries, np.uint64)
        h = C.c_void_p(self.lib.ref_bve_create(nbits, _ptr(data, C.c_uint64), C.c_long(data.shape[0])))
        nq = queries.shape[0]
        idx = np.empty((nq, k), np.int32)
        dist = np.empty((nq, k), np.uint32)
        self.lib.ref_bve_query(h, _ptr(queries, C.c_uint64), nq, k, method, threads, _ptr(idx, C.c_int), _ptr(dist, C.c_uint32))
        self.lib.ref_bve_destroy(h)
        return idx, dist

    def hamming_dist(self, a, b) -> int:
        a = np.ascontiguousarray(a, np.uint64); b = np.ascontiguousarray(b, np.uint64)
        return int(self.lib.ref_hamming_dist(_ptr(a, C.c_uint64), _ptr(b, C.c_uint64), a.size))

    def hamming_dist_sub(self, a, b, sublen: int, subidx: int) -> int:
        a = np.ascontiguousarray(a, np.uint64); b = np.ascontiguousarray(b, np.uint64)
        return int(self.lib.ref_hamming_dist_sub(_ptr(a, C.c_uint64), _ptr(b, C.c_uint64), a.size, sublen, subidx))

    def generate_dummy(self, nbits: int, size: int, seed: int) -> np.ndarray:
        w = (nbits + 63) // 64
        out = np.empty((size, w), np.uint64)
        self.lib.ref_generate_dummy(nbits, size, seed, _ptr(out, C.c_uint64))
        return out

    def create_bitv(self, nbits: int, raw: int) -> np.ndarray:
        out = np.zeros((nbits + 63) // 64, np.uint64)
        self.lib.ref_create_bitv(nbits, C.c_uint64(raw), _ptr(out, C.c_uint64))
        return out

    def nproc(self) -> int:
        return int(self.lib.ref_nproc())

    # -- on-disk formats as written / read by the reference's own utils/IO.hpp
    def save_codebook(self, path: str, codes) -> None:
        codes = np.ascontiguousarray(codes, np.uint16)
        self.lib.ref_save_codebook(str(path).encode(), _ptr(codes, C.c_uint16), C.c_long(codes.shape[0]), codes.shape[1])

    def load_codebook(self, path: str, cap: int = 1 << 24) -> np.ndarray:
        buf = np.empty(cap, np.uint16)
        M = C.c_int(0)
        self.lib.ref_load_codebook.restype = C.c_long
        n = self.lib.ref_load_codebook(str(path).encode(), _ptr(buf, C.c_uint16), C.c_long(cap), C.byref(M))
        return buf[: n * M.value].reshape(n, M.value).copy()

    def save_centroids(self, path: str, m: "Model") -> None:
        self.lib.ref_save_centroids(str(path).encode(), m.M, m.L, _ptr(m.bits, C.c_int), _ptr(m.cent_flat, C.c_float))

    def load_centroids_flat(self, path: str, cap: int = 1 << 24):
        buf = np.empty(cap, np.float32)
        M, L = C.c_int(0), C.c_int(0)
        self.lib.ref_load_centroids.restype = C.c_long
        n = self.lib.ref_load_centroids(str(path).encode(), _ptr(buf, C.c_float), C.c_long(cap), C.byref(M), C.byref(L))
        return buf[:n].copy(), M.value, L.value

    def read_bitv_csv(self, path: str, cols: int, cap_rows: int) -> np.ndarray:
        w = (cols + 63) // 64
        out = np.zeros((cap_rows, w), np.uint64)
        self.lib.ref_read_bitv_csv.restype = C.c_long
        n = self.lib.ref_read_bitv_csv(str(path).encode(), cols, _ptr(out, C.c_uint64), C.c_long(cap_rows))
        return out[:n]

    def write_bitv_csv(self, path: str, words, nbits: int) -> None:
        words = np.ascontiguousarray(words, np.uint64)
        self.lib.ref_write_bitv_csv(str(path).encode(), _ptr(words, C.c_uint64), C.c_long(words.shape[0]), nbits)

    def read_fvecs(self, path: str, dim: int, rows: int) -> np.ndarray:
        out = np.zeros((rows, dim), np.float32)
        self.lib.ref_read_fvecs(str(path).encode(), dim, _ptr(out, C.c_float), C.c_long(rows))
        return out

    def read_ivecs(self, path: str, dim: int, cap_rows: int) -> np.ndarray:
        out = np.zeros((cap_rows, dim), np.int32)
        self.lib.ref_read_ivecs.restype = C.c_long
        n = self.lib.ref_read_ivecs(str(path).encode(), dim, _ptr(out, C.c_int), C.c_long(cap_rows))
        return out[:n]


class RefVAQ:
    def __init__(self, ref: Ref, m: Model, methods: int, eig_real):
        self.ref, self.m = ref, m
        self.lib = ref.lib
        eig = None if eig_real is None else _f32(eig_real)
        self.h = C.c_void_p(self.lib.ref_vaq_create(m.D, m.M, m.L, _ptr(m.bits, C.c_int), _ptr(m.cent_flat, C.c_float),
                                                    None if eig is None else _ptr(eig, C.c_float), m.max_bits, methods))

    def close(self):
        if self.h:
            self.lib.ref_vaq_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_methods(self, methods: int):
        self.lib.ref_vaq_set_methods(self.h, methods)

    def set_visit(self, visit: float):
        self.lib.ref_vaq_set_visit(self.h, C.c_float(visit))

    def set_codes(self, codes):
        codes = np.ascontiguousarray(codes, np.uint16)
        assert codes.shape[1] == self.m.M
        self.lib.ref_vaq_set_codes(self.h, _ptr(codes, C.c_uint16), C.c_long(codes.shape[0]))

    def encode(self, x_proj, nthreads: int = 0) -> np.ndarray:
        x = _f32(x_proj).reshape(-1, self.m.D)
        self.lib.ref_vaq_encode(self.h, _ptr(x, C.c_float), C.c_long(x.shape[0]), self.m.D, nthreads)
        return self.get_codes()

    def get_codes(self) -> np.ndarray:
        n = self.lib.ref_vaq_num_codes(self.h)
        out = np.empty((n, self.m.M), np.uint16)
        self.lib.ref_vaq_get_codes(self.h, _ptr(out, C.c_uint16))
        return out

    def create_lut(self, q_proj) -> np.ndarray:
        """Compact [nq, sum K_s] LUTs gathered from the reference's col-major [2^maxbits, M] table."""
        m = self.m
        q = _f32(q_proj).reshape(-1, m.D)
        rows = 1 << m.max_bits
        full = np.empty((m.M, rows), np.float32)   # col-major [rows x M] == row-major [M x rows]
        out = np.empty((q.shape[0], m.lut_size), np.float32)
        for i in range(q.shape[0]):
            self.lib.ref_vaq_create_lut(self.h, _ptr(q[i], C.c_float), _ptr(full, C.c_float))
            for s in range(m.M):
                out[i, m.lut_off[s]:m.lut_off[s + 1]] = full[s, :m.K[s]]
        return out

    def search(self, queries, k: int, nthreads: int = 1):
        q = _f32(queries).reshape(-1, self.m.D)
        nq = q.shape[0]
        labels = np.empty((nq, k), np.int32)
        dists = np.empty((nq, k), np.float32)
        self.lib.ref_vaq_search(self.h, _ptr(q, C.c_float), nq, self.m.D, k, nthreads, _ptr(labels, C.c_int), _ptr(dists, C.c_float))
        return labels, dists

    def refine(self, queries, in_labels, xtrain, k: int):
        q = _f32(queries); x = _f32(xtrain)
        inl = np.ascontiguousarray(in_labels, np.int32).reshape(q.shape[0], -1)
        labels = np.empty((q.shape[0], k), np.int32)
        dists = np.empty((q.shape[0], k), np.float32)
        self.lib.ref_vaq_refine(self.h, _ptr(q, C.c_float), q.shape[0], q.shape[1], _ptr(inl, C.c_int), inl.shape[1],
                                _ptr(x, C.c_float), C.c_long(x.shape[0]), k, _ptr(labels, C.c_int), _ptr(dists, C.c_float))
        return labels, dists

    def set_ti(self, clusters, start, sizes, members, code_to_cc):
        """clusters [C, segdims]; start/sizes [C]; members: ids in regrouped order (far -> near inside a cluster);
        code_to_cc indexed by original id.  The codebook must already hold the regrouped rows (set_codes)."""
        cl = _f32(clusters); st = np.ascontiguousarray(start, np.int32); sz = np.ascontiguousarray(sizes, np.int32)
        mem = np.ascontiguousarray(members, np.int32); c2c = _f32(code_to_cc)
        self.lib.ref_vaq_set_ti(self.h, _ptr(cl, C.c_float), int(cl.shape[0]), int(cl.shape[1]), _ptr(st, C.c_int),
                                    _ptr(sz, C.c_int), _ptr(mem, C.c_int), _ptr(c2c, C.c_float), C.c_long(c2c.size))

    def cluster_ti(self, n_clusters: int, n_segments: int = -1, use_kmeans: bool = False, seed: int = 1) -> dict:
        self.lib.ref_vaq_cluster_ti(self.h, n_clusters, n_segments, int(use_kmeans), seed)
        segdims = self.lib.ref_vaq_ti_segdims(self.h)
        n = self.lib.ref_vaq_num_codes(self.h)
        clusters = np.empty((n_clusters, segdims), np.float32)
        start = np.empty(n_clusters, np.int32); sizes = np.empty(n_clusters, np.int32)
        members = np.empty(n, np.int32); c2c = np.empty(n, np.float32)
        self.lib.ref_vaq_get_ti(self.h, _ptr(clusters, C.c_float), _ptr(start, C.c_int), _ptr(sizes, C.c_int),
                                _ptr(members, C.c_int), _ptr(c2c, C.c_float))
        return dict(clusters=clusters, start_idx=start, sizes=sizes, members=members, code_to_cc=c2c,
                    codes_grouped=self.get_codes())
