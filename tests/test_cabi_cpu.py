"""CPU-side checks of the C-ABI boundary: the library loads without a GPU and
exports every symbol include/isp_b200.h declares; the ctypes table matches the
header; the product package never imports the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "isp_b200.h")
LIB = os.path.join(ROOT, "isegprobe_b200", "libisp_b200.so")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(isp_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(LIB)


def test_library_exports_every_declared_symbol(built):
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(built, n), f"{n} declared in include/isp_b200.h but not exported"


def test_ctypes_table_covers_header(built):
    from isegprobe_b200 import _lib
    meta = {"isp_version", "isp_last_error", "isp_launch_count"}
    assert set(_declared()) - meta == set(_lib.SIGNATURES), "ctypes SIGNATURES and header disagree"
    # arity check against the header prototypes
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, args in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\((.*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        assert len([a for a in m.group(1).split(",") if a.strip()]) == len(args), name


def test_version_and_error_string(built):
    built.isp_version.restype = ctypes.c_int
    built.isp_last_error.restype = ctypes.c_char_p
    assert built.isp_version() == 1
    assert isinstance(built.isp_last_error(), bytes)


def test_argument_validation_without_gpu(built):
    """Shape/pointer validation happens before any CUDA call, so it is testable on CPU."""
    from isegprobe_b200 import _lib
    L = _lib.lib()
    assert L.isp_distmaps_fwd(None, None, 1, 1, 8, 8, 5.0, 1.0, 1, None) == -1
    assert b"null" in L.isp_last_error()
    assert L.isp_adaptive_conv_fwd(ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), 1, 8, 8, 48, 49, None) == -4
    assert b"64" in L.isp_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "isegprobe_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), fn
                assert "/root/reference" not in txt, fn


def test_cpu_tensors_are_rejected():
    import torch
    from isegprobe_b200 import DistMaps
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DistMaps(5, use_disks=True)(torch.zeros(1, 3, 8, 8), torch.zeros(1, 2, 3))
