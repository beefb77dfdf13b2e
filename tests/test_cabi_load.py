"""The C-ABI library loads on a CPU-only box and exports every symbol include/kmgpu.h declares.
No compute call is made here (there is no GPU and no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "kmgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kmgpu_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from khmer_b200 import cabi
    assert sorted(cabi.SYMBOLS) == _declared()


def test_library_exports_every_declared_symbol():
    from khmer_b200 import cabi
    assert os.path.exists(cabi.LIB_PATH), "libkmgpu.so not built (run __graft_entry__.build())"
    L = ctypes.CDLL(cabi.LIB_PATH)
    for name in _declared():
        assert hasattr(L, name), name
    L.kmgpu_abi_version.restype = ctypes.c_int
    assert L.kmgpu_abi_version() == 2


def test_no_device_fails_loudly():
    """Without a GPU every compute entry refuses: there is no CPU fallback behind the ABI."""
    from khmer_b200 import cabi
    if cabi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(cabi.KmgpuError) as e:
        cabi.Sketch(cabi.BYTE, cabi.TWOBIT, 20, [1009, 1013])
    assert e.value.code == 2  # KMGPU_ENODEV


def test_product_does_not_touch_the_oracle():
    """Nothing under khmer_b200/, bench.py's product arm aside, may reference oracle/."""
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "khmer_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cc", ".hh", ".h", ".cpp")):
                txt = open(os.path.join(dp, fn), errors="replace").read()
                if "oracle" in txt.lower() and "khmer_oracle" in txt or "oracle_lib" in txt or "_ref/libkhmer_ref" in txt:
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad
