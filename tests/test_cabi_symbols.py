"""The C-ABI library loads and exports every symbol include/xpcs_b200.h declares (no GPU
compute is attempted here)."""
import ctypes as C
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "xpcs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xpcs_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(pkg):
    lib = pkg.cabi.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "library does not export %s" % n
    assert sorted(pkg.cabi.SYMBOLS) == names, "cabi.py and the header disagree"


def test_struct_sizes_match(pkg):
    assert C.sizeof(pkg.XpcsParams) == 88 and C.sizeof(pkg.XpcsInfo) == 64 and C.sizeof(pkg.XpcsShardPlan) == 32  # sizeof() in C
    assert pkg.cabi.load().xpcs_compiled_arch() == 100
    assert pkg.cabi.load().xpcs_abi_version() == 1


def test_create_fails_loudly_without_gpu(pkg):
    """No CPU fallback: on a box without a usable sm_100 device, creating a handle raises."""
    import numpy as np
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    dq = np.ones((8, 8), np.int32)
    try:
        pkg.Correlator(dq, dq, 64)
    except pkg.XpcsError as e:
        assert e.code == -2 and "no CPU path" in str(e)
    else:
        raise AssertionError("Correlator was created without a GPU")


def test_bad_params_rejected(pkg):
    lib = pkg.cabi.load()
    h = C.c_void_p()
    p = pkg.XpcsParams()
    p.struct_size = 4
    assert lib.xpcs_create(C.byref(p), 0, C.byref(h)) == -1
    assert b"struct_size" in lib.xpcs_last_error(None)
