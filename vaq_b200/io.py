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
* bit vectors (CSV)     — one text line of 0/1 columns per vector, packed MSB-first into uint64 words
  (readFromExternal :363-397, writeToExternal :681-704); `create_bitv` restates createBitV (BitVector.hpp:46-76)

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


# ---- bit vectors (BitVecEngine inputs) ---------------------------------------------------------------------------

def actual_bitv_len(N: int) -> int:
    """actualBitVLen, BitVector.hpp:36-38"""
    return (int(N) + 63) // 64


def create_bitv(N: int, raw) -> np.ndarray:
    """createBitV, BitVector.hpp:46-76 -> [ceil(N/64)] uint64.

    ``raw`` an int: the scalar overload (:46-61) — N <= 64 stores ``raw`` as the single word.  For N > 64 the
    reference shifts a 64-bit value by multiples of 64 (undefined in C++; x86-64 takes the count modulo 64), so
    every leading word is ``raw`` itself and the last one is ``raw`` masked to the remaining bits; restated as the
    compiled reference behaves.  ``raw`` a sequence: the initializer-list overload (:63-76), one value per word."""
    n = actual_bitv_len(N)
    if isinstance(raw, (int, np.integer)):
        raw = int(raw) & 0xFFFFFFFFFFFFFFFF
        v = np.zeros(n, np.uint64)
        if N <= 64:
            v[0] = raw
        else:
            v[:-1] = raw                                    # raw >> (64 * j), j >= 1: count taken modulo 64
            v[-1] = raw & ((1 << (N - (n - 1) * 64)) - 1)    # raw & LSB(rest)
        return v
    words = np.asarray(list(raw), dtype=np.uint64)
    if words.size != n:
        raise ValueError(f"createBitV({N}, list): expected {n} words, got {words.size}")
    return words.copy()


def read_bitvectors_csv(path, cols: int, delim: str = ",") -> np.ndarray:
    """readFromExternal(filepath, bitvectors&, cols, delim), utils/IO.hpp:363-397 -> [n, ceil(cols/64)] uint64.

    Column c of a line is bit 63 - (c % 64) of word c / 64 (MSB first); reading stops at the first empty line.
    When ``cols`` is not a multiple of 64 the reference shifts the accumulated value once too often before
    left-aligning it (:377-389), which pushes the first column of the last, partial word out of the register; that
    behaviour is reproduced (checked against the compiled reference in tests/test_io_formats.py), so only files
    whose ``cols`` is a multiple of 64 round-trip losslessly — as in the reference."""
    cols = int(cols)
    w = actual_bitv_len(cols)
    rows = []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line:
                break
            fields = line.split(delim)[:cols]
            bits = np.array([int(x) for x in fields], dtype=np.uint64)
            v = np.zeros(w, np.uint64)
            full = bits.size // 64
            for j in range(full):
                word = np.uint64(0)
                for b in bits[j * 64:(j + 1) * 64]:
                    word = (word << np.uint64(1)) | b          # vTemp |= bit ... vTemp <<= 1 between bits
                v[j] = word
            r = bits.size - full * 64
            if r:
                acc = 0
                for b in bits[full * 64:]:
                    acc = ((acc | int(b)) << 1) & 0xFFFFFFFFFFFFFFFF      # the shift after the last bit is the defect
                v[full] = (acc << (64 - r)) & 0xFFFFFFFFFFFFFFFF
            rows.append(v)
    return np.stack(rows) if rows else np.empty((0, w), np.uint64)


def write_bitvectors_csv(path, bv, N: int) -> None:
    """writeToExternal(filepath, const bitvectors&, N), utils/IO.hpp:681-704: per word the top min(64, N % 64 for a
    partial last word) bits, MSB first, comma separated, one vector per line."""
    bv = np.ascontiguousarray(bv, np.uint64).reshape(-1, actual_bitv_len(N))
    with open(path, "w") as f:
        for row in bv:
            parts = []
            for i, val in enumerate(row):
                max_bin = 64 if (i + 1) * 64 <= N else N % 64
                v = int(val)
                parts.extend(str((v >> (63 - b)) & 1) for b in range(max_bin))
            f.write(",".join(parts) + "\n")
