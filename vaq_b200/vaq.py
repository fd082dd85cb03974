"""Python mirror of the reference's two index classes on top of the C ABI: ``VAQ``
(bitvecengine/VAQ.hpp:36-114) and ``BitVecEngine`` (BitVecEngine.hpp:86-106, 1121, 1218) — same method
names, argument meaning and result layout, so the parity tests read like the reference's own callers
(examples/demo_vaq.cpp:56-345).  Training stays on the host (vaq_b200/train.py restates VAQ::train);
encode / search / refine / TI-visit run on the GPU.  There is no CPU search path here.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import train as host_train
from .index import EA, HEAP, PROJECTED, SQRT, TI, HammingIndex, VAQIndex


@dataclass
class LabelDistVecF:
    """utils/Types.hpp:98-104 — query-major [nq, k] labels and distances, ascending distance."""
    labels: np.ndarray
    distances: np.ndarray


class VAQ:
    # VAQ::NNMethod, VAQ.hpp:38-49
    NN_SORT, NN_EA, NN_TI, NN_HEAP = 0x01, 0x02, 0x04, 0x80

    def __init__(self, device: int = 0):
        self.device = device
        self.mBitBudget = 256
        self.mSubspaceNum = 32
        self.mMinBitsPerSubs = 1
        self.mMaxBitsPerSubs = 13
        self.mPercentVarExplained = 1.0
        self.mMethods = self.NN_HEAP
        self.mVisit = 1.0
        self.mTIClusterNum = 0
        self.mTISegmentNum = -1
        self.model: host_train.VAQModel | None = None
        self.index: VAQIndex | None = None
        self._ti = None
        self._raw_key, self._raw_keepalive = None, None

    # VAQ::parseMethodString, VAQ.cpp:1189-1267
    def parseMethodString(self, s: str) -> None:
        p = host_train.parse_method_string(s)
        if p["budget"] is None:
            raise ValueError(f"bad method string {s!r}")
        self.mBitBudget, self.mSubspaceNum = p["budget"], p["M"]
        self.mMinBitsPerSubs, self.mMaxBitsPerSubs = p["min_bits"], p["max_bits"]
        self.mPercentVarExplained = p["var"]
        if p["var"] != 1.0:
            raise ValueError("var<1 truncates mHighestSubs; the reference's scans misread unless it stays a multiple "
                             "of 4 (SURVEY D2) — only var1 is supported")
        m = 0
        for name, bit in (("SORT", self.NN_SORT), ("EA", self.NN_EA), ("TI", self.NN_TI), ("HEAP", self.NN_HEAP)):
            if name in p["methods"]:
                m |= bit
        self.mMethods = m or self.NN_HEAP
        if p["ti_clusters"]:
            self.mTIClusterNum = p["ti_clusters"]
            self.mTISegmentNum = p["ti_segments"]

    # VAQ::train, VAQ.cpp:11-661 (host).  Returns the projected rows: the reference projects its argument
    # in place (VAQ.cpp:294) and encode() expects projected rows.
    def train(self, XTrain: np.ndarray, verbose: bool = False, **kw) -> np.ndarray:
        self.model, XP = host_train.train(XTrain, self.mBitBudget, self.mSubspaceNum, self.mMinBitsPerSubs,
                                          self.mMaxBitsPerSubs, **kw)
        self._new_index()
        return XP

    def load_model(self, model: host_train.VAQModel) -> None:
        self.model = model
        self._new_index()

    def _new_index(self):
        if self.index is not None:
            self.index.close()
        m = self.model
        self.index = VAQIndex(m.L, m.bits, m.centroids, eig=m.eig, device=self.device)
        self._ti = None
        self.reset_raw_cache()

    def reset_raw_cache(self) -> None:
        """Forget which raw vectors are resident on the device (refine() uploads them again)."""
        self._raw_key, self._raw_keepalive = None, None

    # VAQ::encode, VAQ.cpp:663-774 (device; rows already projected)
    def encode(self, XTrainProjected: np.ndarray) -> None:
        self.index.encode_add(XTrainProjected)

    @property
    def mCodebook(self) -> np.ndarray:
        """[N, M] uint16, VAQ.hpp:72 (unpacked from the device layout)."""
        return self.index.get_codes()

    def set_codebook(self, codes: np.ndarray) -> None:
        self.index.add_codes(codes)

    # VAQ::clusterTI, VAQ.cpp:878-999, on the device (csrc/cluster_ti.cu): k-means over the rows decoded in their leading
    # mTISegmentNum segments, rows regrouped by cluster, cluster ranges + original ids kept in HBM.  (The reference also
    # sorts each cluster far->near for its break test, VAQ.cpp:968-982; the device scan is exhaustive inside a visited
    # cluster, so that order is irrelevant.)
    def clusterTI(self, useKMeans: bool = True, verbose: bool = False, iters: int = 10) -> None:
        C = self.mTIClusterNum
        if C <= 0:
            raise ValueError("set mTIClusterNum (method string ...,EA_TI<c>) first")
        self.index.cluster_ti(C, self.mTISegmentNum, iters if useKMeans else 0)
        self._ti = True

    # clusterTI's outputs as the reference holds them (VAQ.hpp:77-84), read back from the device
    @property
    def ti_state(self) -> dict:
        return self.index.get_clusters()

    # VAQ::search, VAQ.cpp:776-847
    def search(self, XTest: np.ndarray, k: int, verbose: bool = False, projected: bool = False) -> LabelDistVecF:
        flags = PROJECTED if (projected or self.model.eig is None) else 0
        if self.mMethods & self.NN_TI:
            self.index.set_visit(self.mVisit)
            flags |= TI | EA | SQRT
        elif self.mMethods & self.NN_EA:
            flags |= EA
        else:
            flags |= HEAP
        X = np.ascontiguousarray(XTest, np.float32)
        if X.shape[1] < self.model.D:
            X = np.pad(X, ((0, 0), (0, self.model.D - X.shape[1])))
        lab, dis = self.index.search(X, k, flags)
        return LabelDistVecF(lab, dis)

    # VAQ::refine, VAQ.cpp:849-876
    def refine(self, XTest: np.ndarray, answersIn: LabelDistVecF, XTrain: np.ndarray, k: int) -> LabelDistVecF:
        # raw rows are uploaded once per (buffer, shape); a new index, a different array or reset_raw_cache() re-uploads.
        # (An array mutated in place keeps its key: call reset_raw_cache() after changing XTrain's contents.)
        X = np.ascontiguousarray(XTrain, np.float32)
        key = (X.ctypes.data, X.shape, X.strides)
        if self._raw_key != key:
            self.index.set_raw_vectors(X)
            self._raw_key, self._raw_keepalive = key, X          # the address stays ours while it is the cache key
        lab, dis = self.index.refine(XTest, answersIn.labels, k)
        return LabelDistVecF(lab, dis)


class BitVecEngine:
    # BitVecEngine::QueryMethod, BitVecEngine.hpp:82-84
    Heap, Sort, HeapEarlyAbandon, SortEarlyAbandon = 0, 1, 2, 3

    def __init__(self, N: int, device: int = 0):
        self.N = int(N)
        self.actBitVLen = (self.N + 63) // 64
        self.index = HammingIndex(self.N, device=device)

    # loadBitV / appendBitV, BitVecEngine.cpp:12, 1630 (rows are [n, actBitVLen] uint64 words)
    def loadBitV(self, bv) -> None:
        if self.index.num_rows:
            dev = self.index.device
            self.index.close()
            self.index = HammingIndex(self.N, device=dev)
        self.appendBitV(bv)

    def appendBitV(self, bv) -> None:
        bv = np.ascontiguousarray(bv, np.uint64).reshape(-1, self.actBitVLen)
        if bv.shape[0]:
            self.index.add(bv)

    @property
    def size(self) -> int:
        return self.index.num_rows

    # query / queryParallel, BitVecEngine.cpp:509-519, 1264-1304 -> (idx [nq,k] int32, dist [nq,k] uint32)
    def query(self, queries, k: int, method: int = 1):
        return self.index.query(queries, k)

    def queryParallel(self, queries, k: int, thread: int = 1):
        return self.index.query(queries, k)
