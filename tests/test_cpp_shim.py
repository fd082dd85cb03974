"""The C++ host shim (include/vaq_gpu.hpp) compiles against the C ABI and behaves like the reference
classes: CPU run -> clean 'no CUDA device' error (no fallback); GPU run -> results equal the oracle."""
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from helpers import ROOT, assert_knn_equiv, bitwise_equal, golden_model, load_golden, orc

SRC = ROOT / "tests" / "cpp" / "shim_check.cpp"


def build_binary(tmp_path: Path) -> Path:
    from vaq_b200 import build as vb
    vb.build()
    exe = tmp_path / "shim_check"
    cxx = shutil.which("g++") or "g++"
    subprocess.run([cxx, "-std=c++14", "-O1", "-Wall", "-Wextra", "-I", str(ROOT / "include"), str(SRC), "-o", str(exe),
                    "-L", str(ROOT / "vaq_b200"), "-lvaqgpu", f"-Wl,-rpath,{ROOT / 'vaq_b200'}"], check=True)
    return exe


def write_input(path: Path, g, m, with_eig: bool):
    eig = g["eig"] if with_eig else None
    q = g["Qraw"] if with_eig else g["Q"]
    hdr = np.array([m.L, m.M, q.shape[0], int(g["k"]), int(with_eig), g["codes"].shape[0]], np.int32)
    with open(path, "wb") as f:
        f.write(hdr.tobytes()); f.write(m.bits.astype(np.int32).tobytes()); f.write(m.cent_flat.tobytes())
        if with_eig:
            f.write(np.ascontiguousarray(eig, np.float32).tobytes())
        f.write(np.ascontiguousarray(g["codes"], np.uint16).tobytes()); f.write(np.ascontiguousarray(q, np.float32).tobytes())


def test_shim_compiles_and_fails_loudly_without_gpu(tmp_path):
    from vaq_b200 import _lib
    exe = build_binary(tmp_path)
    g = load_golden("vaq_small_a")
    m, _ = golden_model(g)
    write_input(tmp_path / "in.bin", g, m, False)
    r = subprocess.run([str(exe), str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    if _lib.device_count() == 0:
        assert r.returncode == 3 and "no CUDA device" in r.stderr, (r.returncode, r.stderr)
    else:
        assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_shim_search_matches_oracle(tmp_path):
    exe = build_binary(tmp_path)
    g = load_golden("vaq_small_a")
    m, _ = golden_model(g)
    write_input(tmp_path / "in.bin", g, m, False)
    r = subprocess.run([str(exe), str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    nq, k = g["Q"].shape[0], int(g["k"])
    raw = np.fromfile(tmp_path / "out.bin", np.uint8)
    n = nq * k * 4
    ea_lab = raw[:n].view(np.int32).reshape(nq, k); ea_dis = raw[n:2 * n].view(np.float32).reshape(nq, k)
    hp_lab = raw[2 * n:3 * n].view(np.int32).reshape(nq, k); hp_dis = raw[3 * n:4 * n].view(np.float32).reshape(nq, k)
    want_lab, want_dis = orc.Port().search_lex(m, g["codes"], g["Q"], k)
    for lab, dis in ((ea_lab, ea_dis), (hp_lab, hp_dis)):
        assert np.array_equal(lab, want_lab) and bitwise_equal(dis, want_dis)
    assert_knn_equiv(ea_lab, ea_dis, g["lab_ea"], g["dis_ea"], what="shim EA vs reference")
    assert_knn_equiv(hp_lab, hp_dis, g["lab_heap"], g["dis_heap"], what="shim HEAP vs reference")
