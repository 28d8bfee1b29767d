"""The C-ABI shared library loads and exports every symbol include/svd_xtend_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "svd_xtend_b200.h")


@pytest.fixture(scope="module")
def lib():
    from svd_xtend_b200 import build
    build.build()
    from svd_xtend_b200 import _lib
    return _lib.load()


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svdx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_covers_header():
    from svd_xtend_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared()


def test_struct_layout_matches(lib):
    from svd_xtend_b200._lib import SvdxAttn, SvdxTapGemm
    assert lib.svdx_struct_size(0) == ctypes.sizeof(SvdxTapGemm)
    assert lib.svdx_struct_size(1) == ctypes.sizeof(SvdxAttn)
    assert lib.svdx_struct_size(7) == -1


def test_bad_arguments_fail_loudly_without_gpu(lib):
    from svd_xtend_b200._lib import SvdxError, SvdxTapGemm, check
    assert lib.svdx_tapgemm(None, None) == -1
    d = SvdxTapGemm()
    assert lib.svdx_tapgemm(ctypes.byref(d), None) == -1
    assert b"null" in lib.svdx_last_error()
    assert lib.svdx_attention_fwd(None, None) == -1
    assert lib.svdx_layernorm_fwd(None, 0, 0, 0, None, None, 0.0, None, 0, None, None, None, 0, None, 0, None) == -1
    assert lib.svdx_prep_weight(None, 0, None, 0, 0, 0, 0, 0, None) == -1
    with pytest.raises(SvdxError):
        check(lib.svdx_colsum(None, 0, 0, 0, None, 0, None), "colsum")
