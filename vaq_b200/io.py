"""On-disk formats of the reference (bitvecengine/utils/IO.hpp), so indexes and data sets written by the
reference load here and vice versa.  Host-side, numpy only.

* fvecs / ivecs / bvecs  — per record an int32 dimension followed by that many float32 / int32 / uint8
  (readFVecsFromExternal :126-161, readIVecsFromExternal :334-361, readBVecsFromExternal :198-233)
* raw bin               — row-major float32, no header (readFromExternalBin :261-288)
* centroids             — size_t n; then per subspace size_t rows, size_t cols, float32[rows*cols] row-major
  (saveCentroids :736-754, loadCentroids :522-549)
* codebook              — size_t rows, size_t cols, uint16[rows*cols] row-major (saveCodebook :757-772,
  loadCodebook :552-571)
* kNN results           — one text line of comma-separated labels per query (writeKNNResults :720-734)

size_t is 8 bytes (the reference is built for x86-64).
"""
from __future__ import annotations

import os

import numpy as np


def _read_vecs(path, dtype, max_rows: int = -1) -> np.ndarray:
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size == 0:
        return np.empty((0, 0), dtype)
    dim = int(raw[:4].view(np.int32)[0])
    item = np.dtype(dtype).itemsize
    rec = 4 + dim * item
    if dim <= 0 or raw.size % rec:
        raise ValueError(f"{path}: not a well-formed vecs file (dim={dim}, {raw.size} bytes)")
    n = raw.size // rec
    recs = raw.reshape(n, rec)
    if not (recs[:, :4].view(np.int32)[:, 0] == dim).all():
        raise ValueError(f"{path}: records with differing dimensions")
    out = np.ascontiguousarray(recs[:, 4:]).view(dtype).reshape(n, dim)
    return out[:max_rows] if max_rows >= 0 else out


def read_fvecs(path, max_rows: int = -1) -> np.ndarray:
    return _read_vecs(path, np.float32, max_rows)


def read_ivecs(path, max_rows: int = -1) -> np.ndarray:
    return _read_vecs(path, np.int32, max_rows)


def read_bvecs(path, max_rows: int = -1) -> np.ndarray:
    """uint8 components, returned as float32 like the reference's reader (it fills a float matrix)."""
    return _read_vecs(path, np.uint8, max_rows).astype(np.float32)


def _write_vecs(path, a: np.ndarray, dtype) -> None:
    a = np.ascontiguousarray(a, dtype)
    n, dim = a.shape
    rec = np.empty((n, 4 + dim * a.itemsize), np.uint8)
    rec[:, :4] = np.frombuffer(np.int32(dim).tobytes(), np.uint8)
    rec[:, 4:] = a.view(np.uint8).reshape(n, -1)
    rec.tofile(path)


def write_fvecs(path, a) -> None:
    _write_vecs(path, a, np.float32)


def write_ivecs(path, a) -> None:
    _write_vecs(path, a, np.int32)


def write_bvecs(path, a) -> None:
    _write_vecs(path, a, np.uint8)


def read_bin(path, dim: int, max_rows: int = -1) -> np.ndarray:
    a = np.fromfile(path, dtype=np.float32, count=-1 if max_rows < 0 else max_rows * dim)
    return a[: (a.size // dim) * dim].reshape(-1, dim)


def save_centroids(path, centroids) -> None:
    with open(path, "wb") as f:
        f.write(np.uint64(len(centroids)).tobytes())
        for c in centroids:
            c = np.ascontiguousarray(c, np.float32)
            f.write(np.array(c.shape, np.uint64).tobytes())
            f.write(c.tobytes())


def load_centroids(path) -> list[np.ndarray]:
    raw = memoryview(open(path, "rb").read())
    n = int(np.frombuffer(raw[:8], np.uint64)[0])
    off, out = 8, []
    for _ in range(n):
        rows, cols = (int(x) for x in np.frombuffer(raw[off:off + 16], np.uint64))
        off += 16
        out.append(np.frombuffer(raw[off:off + rows * cols * 4], np.float32).reshape(rows, cols).copy())
        off += rows * cols * 4
    return out


def save_codebook(path, codes) -> None:
    codes = np.ascontiguousarray(codes, np.uint16)
    with open(path, "wb") as f:
        f.write(np.array(codes.shape, np.uint64).tobytes())
        f.write(codes.tobytes())


def load_codebook(path) -> np.ndarray:
    with open(path, "rb") as f:
        rows, cols = (int(x) for x in np.frombuffer(f.read(16), np.uint64))
        a = np.fromfile(f, dtype=np.uint16, count=rows * cols)
    if a.size != rows * cols:
        raise ValueError(f"{path}: truncated codebook ({a.size} of {rows * cols} codes)")
    return a.reshape(rows, cols)


def write_knn_results(path, labels) -> None:
    labels = np.asarray(labels)
    with open(path, "w") as f:
        for row in labels:
            f.write(",".join(str(int(x)) for x in row) + os.linesep)
