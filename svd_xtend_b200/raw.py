"""Thin tensor-level wrappers over the C ABI (no autograd): pointer/shape marshalling only."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import A_CONV2D, A_ROWS, OUT_BF16, OUT_F32, OUT_F32_ATOMIC, SvdxAttn, SvdxTapGemm, check, load

bf16 = torch.bfloat16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _rowmajor(t: torch.Tensor, name: str) -> int:
    """leading dimension (elements) of a 2-D-like view whose last dim is contiguous."""
    if t.stride(-1) != 1:
        raise ValueError(f"{name}: last dimension must be contiguous")
    return t.stride(-2)


def pick_block_n(n_out: int, mn_major: bool = False) -> int:
    if mn_major:
        for bn in (256, 192, 128, 64):
            if n_out % bn == 0:
                return bn
        return 256 if n_out > 256 else ((n_out + 63) // 64) * 64
    for bn in (256, 160, 128, 96, 64, 32):
        if n_out % bn == 0:
            return bn
    return 256 if n_out > 256 else ((n_out + 31) // 32) * 32


def tapgemm(
    a: torch.Tensor,
    b: torch.Tensor,
    out: torch.Tensor,
    *,
    M: int,
    N: int,
    K: int,
    mode: int = A_ROWS,
    taps: Sequence[Sequence[int]] = ((0, 0, 0),),
    rows_per_group: Optional[int] = None,
    groups: int = 1,
    conv_whn: Optional[Sequence[int]] = None,
    lda: Optional[int] = None,
    ldb: Optional[int] = None,
    ldo: Optional[int] = None,
    a_mn: bool = False,
    b_mn: bool = False,
    block_n: Optional[int] = None,
    split_k: int = 1,
    out_dtype: Optional[int] = None,
    geglu: bool = False,
    bias: Optional[torch.Tensor] = None,
    rowbias: Optional[torch.Tensor] = None,
    rowbias_div: int = 1,
    res1: Optional[torch.Tensor] = None,
    res2: Optional[torch.Tensor] = None,
    scales: Optional[torch.Tensor] = None,
    pre: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """Launch svdx_tapgemm on the current stream. All tensors are CUDA; a/b/res/pre are bf16."""
    d = SvdxTapGemm()
    assert a.dtype == bf16 and b.dtype == bf16
    d.a = a.data_ptr()
    d.lda = lda if lda is not None else _rowmajor(a, "a")
    d.a_mode = mode
    d.a_major_mn = int(a_mn)
    d.rows_per_group = rows_per_group if rows_per_group is not None else M
    d.groups = groups
    if conv_whn is not None:
        d.W, d.H, d.nimg = conv_whn
    d.num_taps = len(taps)
    for i, t in enumerate(taps):
        d.tap_d0[i], d.tap_d1[i], d.tap_d2[i] = int(t[0]), int(t[1]), int(t[2])
    d.b = b.data_ptr()
    d.ldb = ldb if ldb is not None else _rowmajor(b, "b")
    d.b_major_mn = int(b_mn)
    d.M, d.N, d.K = M, N, K
    n_out = N // 2 if geglu else N
    if block_n is None:
        block_n = 2 * pick_block_n(n_out) if geglu else pick_block_n(n_out, b_mn)
        if geglu and block_n > 256:
            block_n = 256 if n_out % 128 == 0 else 128
    d.block_n = block_n
    d.split_k = split_k
    d.out = out.data_ptr()
    d.ldo = ldo if ldo is not None else _rowmajor(out, "out")
    if out_dtype is None:
        out_dtype = OUT_BF16 if out.dtype == bf16 else OUT_F32
    d.out_dtype = out_dtype
    d.geglu = int(geglu)
    if bias is not None:
        assert bias.dtype == torch.float32
        d.bias = bias.data_ptr()
    if rowbias is not None:
        assert rowbias.dtype == torch.float32
        d.rowbias = rowbias.data_ptr()
        d.rowbias_div = rowbias_div
        d.ldrb = _rowmajor(rowbias, "rowbias")
    if res1 is not None:
        assert res1.dtype == bf16
        d.res1 = res1.data_ptr()
        d.ldr1 = _rowmajor(res1, "res1")
    if res2 is not None:
        assert res2.dtype == bf16
        d.res2 = res2.data_ptr()
        d.ldr2 = _rowmajor(res2, "res2")
    if scales is not None:
        assert scales.dtype == torch.float32 and scales.numel() >= 3
        d.scales = scales.data_ptr()
    if pre is not None:
        assert pre.dtype == bf16
        d.pre = pre.data_ptr()
        d.ldpre = _rowmajor(pre, "pre")
    check(load().svdx_tapgemm(C.byref(d), _stream()), "svdx_tapgemm")
    return out


CONV3x3_TAPS = tuple((kw - 1, kh - 1, 0) for kh in range(3) for kw in range(3))
