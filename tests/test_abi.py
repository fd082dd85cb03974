"""CPU: the C-ABI library loads, exports every symbol include/vaqgpu.h declares, and fails loudly
(no fallback) when there is no GPU."""
import ctypes
import re

import numpy as np
import pytest

from helpers import ROOT


def declared_symbols():
    text = (ROOT / "include" / "vaqgpu.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:vaqgpu|hamgpu)_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    from vaq_b200 import _lib, build
    build.build()
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 35
    raw = ctypes.CDLL(str(_lib.LIB_PATH))
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in vaqgpu.h but not exported"
        assert s in _lib.PROTOTYPES, f"{s} has no ctypes prototype in vaq_b200/_lib.py"
    assert set(_lib.PROTOTYPES) == set(syms)
    assert lib.vaqgpu_last_error() is not None


def test_argument_validation_needs_no_device():
    from vaq_b200 import _lib
    from vaq_b200.index import VAQIndex
    with pytest.raises(_lib.VaqGpuError) as e:
        VAQIndex(2, [0, 4], [np.zeros((1, 2), np.float32), np.zeros((16, 2), np.float32)])
    assert e.value.code == _lib.VAQGPU_EINVAL and "bits[0]" in str(e.value)
    with pytest.raises(ValueError):
        VAQIndex(2, [4, 4], [np.zeros((16, 2), np.float32)])


def test_no_cpu_fallback():
    from vaq_b200 import _lib
    from vaq_b200.index import HammingIndex, VAQIndex
    if _lib.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(_lib.VaqGpuError) as e:
        VAQIndex(2, [4, 4], [np.zeros((16, 2), np.float32)] * 2)
    assert e.value.code == _lib.VAQGPU_ECUDA and "no CPU path" in str(e.value)
    with pytest.raises(_lib.VaqGpuError) as e:
        HammingIndex(256)
    assert e.value.code == _lib.VAQGPU_ECUDA


def test_product_never_imports_the_oracle():
    for p in (ROOT / "vaq_b200").rglob("*.py"):
        assert "oracle" not in p.read_text().replace("the oracle", ""), f"{p} mentions the oracle package"
    for p in (ROOT / "vaq_b200" / "csrc").glob("*.cu*"):
        assert "oracle/" not in p.read_text()
