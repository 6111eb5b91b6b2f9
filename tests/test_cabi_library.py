"""CPU-only checks of the C-ABI boundary: the library builds/loads and exports every
symbol include/ganecdotes_b200.h declares; the product fails loudly without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ganecdotes_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ganecdotes_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(build.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    # and the ctypes binding covers exactly the declared surface
    assert sorted(_lib.EXPORTED_SYMBOLS) == names
    assert lib.gx_version() == _lib.GX_ABI_VERSION
    header = open(os.path.join(ROOT, "include", "ganecdotes_b200.h")).read()
    assert f"#define GX_ABI_VERSION {_lib.GX_ABI_VERSION}" in header
    for which, st in enumerate((_lib.gx_conv_desc, _lib.gx_gemm_desc, _lib.gx_gather_desc, _lib.gx_ll_desc)):
        assert lib.gx_abi_sizeof(which) == ctypes.sizeof(st)      # the handshake load() performs
    lib.gx_error_string.restype = ctypes.c_char_p
    assert b"argument" in lib.gx_error_string(-1)


def test_argument_validation_without_gpu():
    """entry points validate arguments before touching the device"""
    from ganecdotes_b200 import _lib
    lib = _lib.load(require_device=False)
    assert lib.gx_upfirdn2d(None, None, None, 1, 4, 4, 1, 4, 4, 1, 1, 1, 1, 0, 0, 0, 0, None) == -1
    assert lib.gx_sinkhorn_pass(None, 10, 7, 7, 1.0, 1, None, None, None, None, 10, 0, None, None, None) == -1
    d = _lib.gx_gemm_desc()
    assert lib.gx_gemm(ctypes.byref(d), None) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a CPU-only box")
def test_no_cpu_fallback():
    from ganecdotes_b200 import _lib
    from ganecdotes_b200.stylegan2.op import upfirdn2d, fused_leaky_relu
    with pytest.raises(RuntimeError):
        _lib.load()
    with pytest.raises(RuntimeError):
        upfirdn2d(torch.randn(1, 1, 4, 4), torch.ones(2, 2))
    with pytest.raises(RuntimeError):
        fused_leaky_relu(torch.randn(1, 4, 4, 4), torch.zeros(4))


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under ganecdotes_b200/ may reference it"""
    pkg = os.path.join(ROOT, "ganecdotes_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("the oracle", ""), os.path.join(dp, f)
