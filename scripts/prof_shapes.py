"""Launches ONE instance of every hot kernel at its BASELINE config-2 shape (40x64 latents, 14 frames), in a fixed order, for

    ncu --set full --clock-control none --import-source on -k regex:'tapgemm|attn_|gn_|ln_|adamw|geglu|softmax' -o gpurun_out/prof_rN python scripts/prof_shapes.py

Each kernel prints a tag line to stdout BEFORE its launch; scripts/summarize_ncu_full.py pairs the tags with the captured launches
in order. Numbers printed by a run under ncu are never bench values."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svd_xtend_b200 import raw  # noqa: E402

DEV = "cuda:0"
bf16 = torch.bfloat16
TAGS = []


def tag(name, flops=0.0, nbytes=0.0, launches=1):
    TAGS.append(dict(name=name, flops=flops, bytes=nbytes, launches=launches))
    print(f"TAG {name}", flush=True)


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=DEV) * scale).to(bf16)


def main():
    T, H, W, C = 14, 40, 64, 320
    M = T * H * W
    x = rnd(M, C)
    torch.cuda.synchronize()
    # ---- tapgemm family
    w9 = rnd(C, 9 * C, scale=(9 * C) ** -0.5)
    out = torch.empty(M, C, device=DEV, dtype=bf16)
    bias = torch.randn(C, device=DEV)
    sums = torch.zeros(T, 2, C, device=DEV)
    tag("conv3x3 L0 320->320 (+bias, fused GN stats) CTA-pair kernel, 256x320 tile", 2.0 * M * C * 9 * C)
    raw.tapgemm(x, w9, out, M=M, N=C, K=C, mode=raw.A_CONV2D, taps=raw.CONV3x3_TAPS, conv_whn=(W, H, T), bias=bias, gn_sum=sums, gn_rows=H * W)
    res = rnd(M, C)
    tag("conv3x3 L0 320->320 (+bias +residual) CTA-pair kernel, 256x320 tile", 2.0 * M * C * 9 * C)
    raw.tapgemm(x, w9, out, M=M, N=C, K=C, mode=raw.A_CONV2D, taps=raw.CONV3x3_TAPS, conv_whn=(W, H, T), bias=bias, res1=res)
    w3 = rnd(C, 3 * C, scale=(3 * C) ** -0.5)
    tag("temporal conv (3,1,1) L0 320->320", 2.0 * M * C * 3 * C)
    raw.tapgemm(x, w3, out, M=M, N=C, K=C, taps=[(-H * W, 0, 0), (0, 0, 0), (H * W, 0, 0)], rows_per_group=M, groups=1, bias=bias)
    wqkv = rnd(3 * C, C, scale=C ** -0.5)
    oqkv = torch.empty(M, 3 * C, device=DEV, dtype=bf16)
    tag("linear q|k|v L0 N=960 K=320", 2.0 * M * 3 * C * C)
    raw.tapgemm(x, wqkv, oqkv, M=M, N=3 * C, K=C)
    wp = rnd(C, C, scale=C ** -0.5)
    tag("linear proj L0 N=320 K=320 (+bias +residual)", 2.0 * M * C * C)
    raw.tapgemm(x, wp, out, M=M, N=C, K=C, bias=bias, res1=res)
    wff = rnd(8 * C, C, scale=C ** -0.5)
    bff = torch.randn(8 * C, device=DEV)
    off = torch.empty(M, 4 * C, device=DEV, dtype=bf16)
    pre = torch.empty(M, 8 * C, device=DEV, dtype=bf16)
    tag("GEGLU projection L0 N=2560 K=320 (fused GEGLU + saved pre-activation)", 2.0 * M * 8 * C * C)
    raw.tapgemm(x, wff, off, M=M, N=8 * C, K=C, bias=bff, geglu=True, pre=pre)
    wo = rnd(C, 4 * C, scale=(4 * C) ** -0.5)
    tag("linear ff.out L0 N=320 K=1280 (+bias +residual)", 2.0 * M * C * 4 * C)
    raw.tapgemm(off, wo, out, M=M, N=C, K=4 * C, bias=bias, res1=res)
    dpre = rnd(M, 8 * C)
    gw = torch.zeros(8 * C, C, device=DEV)
    bn, split = raw.wgrad_plan(8 * C, C, M)
    tag(f"weight gradient dW[2560,320] += dy^T x, tokens 35840 (1-CTA kernel, bn {bn} split {split})", 2.0 * M * 8 * C * C)
    raw.tapgemm(dpre, x, gw, M=8 * C, N=C, K=M, a_mn=True, b_mn=True, split_k=split, out_dtype=raw.OUT_F32_ATOMIC, block_n=bn, lda=8 * C, ldb=C)
    gw2 = torch.zeros(C, 4 * C, device=DEV)
    bn, split = raw.wgrad_plan(C, 4 * C, M)
    tag(f"weight gradient dW[320,1280] += dy^T x, tokens 35840 (bn {bn} split {split})", 2.0 * M * C * 4 * C)
    raw.tapgemm(x, off, gw2, M=C, N=4 * C, K=M, a_mn=True, b_mn=True, split_k=split, out_dtype=raw.OUT_F32_ATOMIC, block_n=bn, lda=C, ldb=4 * C)
    wt = rnd(8 * C, C, scale=C ** -0.5)     # dgrad operand [K_out = 8C rows? no: N = C outputs, K = 8C]
    wdg = rnd(C, 8 * C, scale=(8 * C) ** -0.5)
    tag("dgrad of the GEGLU projection L0 N=320 K=2560", 2.0 * M * C * 8 * C)
    raw.tapgemm(dpre, wdg, out, M=M, N=C, K=8 * C)
    # L2-level conv (10x16 latents, C = 1280)
    M2, C2 = T * 10 * 16, 1280
    x2 = rnd(M2, C2)
    w92 = rnd(C2, 9 * C2, scale=(9 * C2) ** -0.5)
    o2 = torch.empty(M2, C2, device=DEV, dtype=bf16)
    b2 = torch.randn(C2, device=DEV)
    tag("conv3x3 L2 1280->1280 (10x16 latents)", 2.0 * M2 * C2 * 9 * C2)
    raw.tapgemm(x2, w92, o2, M=M2, N=C2, K=C2, mode=raw.A_CONV2D, taps=raw.CONV3x3_TAPS, conv_whn=(16, 10, T), bias=b2)
    # ---- attention
    heads = 5
    lse = torch.empty(M, heads, device=DEV)
    delta = torch.empty(M, heads, device=DEV)
    o = torch.empty(M, C, device=DEV, dtype=bf16)
    dqkv = torch.empty(M, 3 * C, device=DEV, dtype=bf16)
    q, k, v = oqkv[:, :C], oqkv[:, C:2 * C], oqkv[:, 2 * C:]
    sp = dict(heads=heads, S=H * W, nseq=T, inner=1, outer_stride=H * W, inner_stride=0, tok_stride=1)
    fl = 4.0 * T * heads * (H * W) ** 2 * 64
    tag("attention forward, spatial L0 (S=2560, 5 heads, 14 frames)", fl)
    raw.attention_fwd(q, k, v, o, lse=lse, **sp)
    tag("attention backward, spatial L0 (delta + dQ + dK/dV)", 2.5 * fl, launches=3)
    raw.attention_bwd(q, k, v, o, res, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], lse, delta, **sp)
    tp = dict(heads=heads, S=T, nseq=H * W, inner=H * W, outer_stride=T * H * W, inner_stride=1, tok_stride=H * W)
    tag("attention forward, temporal L0 (S=14, 2560 pixels x 5 heads) short-sequence kernel", 0.0, 4.0 * M * C * 2)
    raw.attention_fwd(q, k, v, o, lse=lse, **tp)
    tag("attention backward, temporal L0 short-sequence kernel", 0.0, 8.0 * M * C * 2)
    raw.attention_bwd(q, k, v, o, res, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], lse, delta, **tp)
    # ---- norms / elementwise
    gamma, beta = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    y = torch.empty(M, C, device=DEV, dtype=bf16)
    tag("GroupNorm+SiLU apply from fused channel sums L0", 0.0, 4.0 * M * C)
    mean, rstd = raw.groupnorm_apply_fused(out, None, T, H * W, 1e-5, sums, None, gamma, beta, True, y)
    ws = torch.zeros(2 * T * 32, device=DEV)
    tag("GroupNorm+SiLU backward L0 (partial sums + apply, + residual gradient)", 0.0, 8.0 * M * C, launches=2)
    raw.groupnorm_bwd(out, None, res, T, H * W, mean, rstd, gamma, beta, True, y, None, ws=ws, dres=x)
    ab = torch.empty(T, 2, C, device=DEV)
    raw.groupnorm_apply(out, None, T, H * W, mean, rstd, gamma, beta, True, y, ab=ab)      # (captured too: plain apply)
    TAGS.insert(len(TAGS), dict(name="GroupNorm+SiLU apply L0 (stand-alone statistics given)", flops=0.0, bytes=4.0 * M * C, launches=1))
    print("TAG GroupNorm+SiLU apply L0 (stand-alone statistics given)", flush=True)
    gsum = torch.zeros(T, 2, C, device=DEV)
    dyb = torch.empty(M, C, device=DEV, dtype=bf16)
    tag("conv3x3 dgrad L0 320->320 + GroupNorm-backward sums in the epilogue (256x320 tile)", 2.0 * M * C * 9 * C)
    raw.tapgemm(x, w9, dyb, M=M, N=C, K=C, mode=raw.A_CONV2D, taps=raw.CONV3x3_TAPS, conv_whn=(W, H, T),
                gnb=dict(x=out, x2=None, ab=ab, rows=H * W, silu=True, sum=gsum))
    tag("GroupNorm+SiLU backward L0 from the epilogue sums (one launch, + residual gradient)", 0.0, 8.0 * M * C)
    raw.groupnorm_bwd_fused(out, None, dyb, T, H * W, mean, rstd, gamma, beta, True, gsum, y, None, dres=x)
    tag("LayerNorm forward L0", 0.0, 4.0 * M * C)
    lm, lr = raw.layernorm_fwd(x, gamma, beta, 1e-5, y)
    tag("LayerNorm backward L0 (+ residual gradient)", 0.0, 8.0 * M * C)
    raw.layernorm_bwd(x, res, gamma, lm, lr, y, out)
    bg = torch.zeros(8 * C, device=DEV)
    tag("GEGLU backward L0 (+ fused bias gradient)", 0.0, 10.0 * M * 4 * C)
    raw.geglu_bwd(pre, off, dpre, bias_grad=bg)
    n = 397_620_480 // 4     # a quarter of the as-scripted trainable set (keeps the profile run short)
    p_ = torch.zeros(n, device=DEV)
    g_ = torch.randn(n, device=DEV)
    m_, v_ = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    sh = torch.empty(n, device=DEV, dtype=bf16)
    state = torch.tensor([1e-5, 0.9, 0.999, 1e-8, 1e-2, 0.0, 1.0, 1.0], device=DEV)
    tag("AdamW + bf16 shadow (99.4 M parameters)", 0.0, 30.0 * n, launches=2)
    raw.adamw_graph(p_, g_, m_, v_, state, 1.0, shadow=sh)
    torch.cuda.synchronize()
    import json
    json.dump(TAGS, open(os.path.join(ROOT, "gpurun_out", "prof_tags.json"), "w"))


if __name__ == "__main__":
    main()
