"""Micro-benchmarks of single kernels at the shapes of the BASELINE config-2 step, timed the way they run in the product:
inside a CUDA graph (REPS launches captured back to back, replayed, CUDA events), on buffers rotated through > 126 MB so that
consecutive launches do not hit in L2.

    python scripts/kbench.py gn ln attn wgrad ...      (sections; default: all)
Environment knobs of the library (SVDX_GN_CTAS_PER_SM, ...) are read once per process: sweep them from the shell."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svd_xtend_b200 import raw  # noqa: E402

DEV = "cuda:0"
bf16 = torch.bfloat16
F32 = torch.float32


def graph_time(fn_list, reps=20):
    """fn_list: callables (one per rotated buffer set); returns us per call inside a captured graph"""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for f in fn_list:
            f()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    n = 0
    with torch.cuda.graph(g):
        for _ in range(reps):
            for f in fn_list:
                f()
                n += 1
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (3 * n)


def rot(nbytes_per_set):
    """number of rotated buffer sets so that one pass touches > 160 MB (KBENCH_WARM=1: one set, everything stays in L2 —
    the other bracket of what a kernel sees inside the step, where its input was just written by the producer)"""
    if os.environ.get("KBENCH_WARM") == "1":
        return 1
    return max(2, int(160e6 // max(nbytes_per_set, 1)) + 1)


LEVELS = [(14, 2560, 320), (14, 640, 640), (14, 160, 1280), (14, 40, 1280)]      # (frames, HW, C) of config 2


def sec_gn():
    print("# GroupNorm kernels (per-frame statistics), us per launch in a graph; GB/s on algorithmic bytes")
    for T, HW, C in LEVELS:
        M = T * HW
        k = rot(M * C * 2 * 3)
        xs = [torch.randn(M, C, device=DEV).to(bf16) for _ in range(k)]
        dys = [torch.randn(M, C, device=DEV).to(bf16) for _ in range(k)]
        ys = [torch.empty(M, C, device=DEV, dtype=bf16) for _ in range(k)]
        gamma, beta = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
        st = [raw.groupnorm_stats(x, None, T, HW, 1e-5) for x in xs]
        csum = [torch.stack([x.float().view(T, HW, C).sum(1), (x.float() ** 2).view(T, HW, C).sum(1)], 1).contiguous() for x in xs]
        ws = torch.zeros(2 * T * 32, device=DEV)
        t_stats = graph_time([lambda x=x: raw.groupnorm_stats(x, None, T, HW, 1e-5) for x in xs])
        t_apply = graph_time([lambda x=x, y=y, s=s: raw.groupnorm_apply(x, None, T, HW, s[0], s[1], gamma, beta, True, y) for x, y, s in zip(xs, ys, st)])
        t_fused = graph_time([lambda x=x, y=y, c=c: raw.groupnorm_apply_fused(x, None, T, HW, 1e-5, c, None, gamma, beta, True, y) for x, y, c in zip(xs, ys, csum)])
        t_bwd = graph_time([lambda x=x, y=y, d=d, s=s: raw.groupnorm_bwd(x, None, d, T, HW, s[0], s[1], gamma, beta, True, y, None) for x, y, d, s in zip(xs, ys, dys, st)])
        by = M * C * 2
        print(f"  M={M:6d} C={C:5d}: stats(memset+partial+finalize) {t_stats:7.1f} us | apply {t_apply:6.1f} us ({2 * by / t_apply / 1e3:6.0f} GB/s) | "
              f"apply_fused {t_fused:6.1f} us ({2 * by / t_fused / 1e3:6.0f} GB/s) | bwd(memset+partial+apply) {t_bwd:6.1f} us ({3 * by / t_bwd / 1e3:6.0f} GB/s)", flush=True)


def sec_ln():
    print("# LayerNorm kernels, us per launch in a graph")
    for T, HW, C in LEVELS:
        M = T * HW
        k = rot(M * C * 2 * 3)
        xs = [torch.randn(M, C, device=DEV).to(bf16) for _ in range(k)]
        dys = [torch.randn(M, C, device=DEV).to(bf16) for _ in range(k)]
        ys = [torch.empty(M, C, device=DEV, dtype=bf16) for _ in range(k)]
        gamma, beta = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
        st = [raw.layernorm_fwd(x, gamma, beta, 1e-5, y) for x, y in zip(xs, ys)]
        t_f = graph_time([lambda x=x, y=y: raw.layernorm_fwd(x, gamma, beta, 1e-5, y) for x, y in zip(xs, ys)])
        t_b = graph_time([lambda x=x, y=y, d=d, s=s: raw.layernorm_bwd(x, d, gamma, s[0], s[1], y, d) for x, y, d, s in zip(xs, ys, dys, st)])
        dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
        t_bg = graph_time([lambda x=x, y=y, d=d, s=s: raw.layernorm_bwd(x, d, gamma, s[0], s[1], y, d, dg, db) for x, y, d, s in zip(xs, ys, dys, st)])
        by = M * C * 2
        print(f"  M={M:6d} C={C:5d}: fwd {t_f:6.1f} us ({2 * by / t_f / 1e3:6.0f} GB/s) | bwd(+dres) {t_b:6.1f} us ({4 * by / t_b / 1e3:6.0f} GB/s) | "
              f"bwd(+dres,+dgamma) {t_bg:6.1f} us", flush=True)


def sec_attn():
    print("# attention kernels, us per launch in a graph (fwd | bwd = delta + dQ + dK/dV)")
    for name, T, HW, heads, temporal in [("spatial L0", 14, 2560, 5, False), ("spatial L1", 14, 640, 10, False), ("spatial L2", 14, 160, 20, False),
                                         ("spatial L3", 14, 40, 20, False), ("temporal L0", 14, 2560, 5, True), ("temporal L1", 14, 640, 10, True),
                                         ("temporal L2", 14, 160, 20, True), ("temporal L3", 14, 40, 20, True)]:
        C = heads * 64
        M = T * HW
        k = rot(M * 3 * C * 2 * 2)
        qkvs = [torch.randn(M, 3 * C, device=DEV).to(bf16) for _ in range(k)]
        outs = [torch.empty(M, C, device=DEV, dtype=bf16) for _ in range(k)]
        dqkv = [torch.empty(M, 3 * C, device=DEV, dtype=bf16) for _ in range(k)]
        lse = torch.empty(M, heads, device=DEV)
        delta = torch.empty(M, heads, device=DEV)
        geo = (dict(heads=heads, S=T, nseq=HW, inner=HW, outer_stride=T * HW, inner_stride=1, tok_stride=HW) if temporal
               else dict(heads=heads, S=HW, nseq=T, inner=1, outer_stride=HW, inner_stride=0, tok_stride=1))

        def f(q, o):
            raw.attention_fwd(q[:, :C], q[:, C:2 * C], q[:, 2 * C:], o, lse=lse, **geo)

        def b(q, o, d):
            raw.attention_bwd(q[:, :C], q[:, C:2 * C], q[:, 2 * C:], o, o, d[:, :C], d[:, C:2 * C], d[:, 2 * C:], lse, delta, **geo)

        t_f = graph_time([lambda q=q, o=o: f(q, o) for q, o in zip(qkvs, outs)])
        t_b = graph_time([lambda q=q, o=o, d=d: b(q, o, d) for q, o, d in zip(qkvs, outs, dqkv)])
        S = T if temporal else HW
        fl = 4.0 * (M // S) * heads * S * S * 64
        print(f"  {name:12s} S={S:5d} seqs={M // S:6d} heads={heads:2d}: fwd {t_f:7.1f} us ({fl / t_f / 1e6:6.0f} TF/s) | bwd {t_b:7.1f} us ({2.5 * fl / t_b / 1e6:6.0f} TF/s) "
              f"| bytes fwd {4 * M * C * 2 / t_f / 1e3:6.0f} GB/s", flush=True)


def sec_wgrad():
    print("# weight-gradient GEMMs dW[O,K] += dy[M,O]^T x[M,K] (MN-major operands), us per launch in a graph")
    for M, O, K in [(35840, 2560, 320), (35840, 320, 1280), (35840, 960, 320), (35840, 320, 320), (8960, 5120, 640), (8960, 640, 2560),
                    (8960, 1920, 640), (8960, 640, 640), (2240, 10240, 1280), (2240, 1280, 5120), (2240, 3840, 1280), (2240, 1280, 1280)]:
        k = rot((M * O + M * K) * 2)
        dys = [torch.randn(M, O, device=DEV).to(bf16) for _ in range(k)]
        xs = [torch.randn(M, K, device=DEV).to(bf16) for _ in range(k)]
        g = torch.zeros(O, K, device=DEV)
        res = []
        for bn in (None, 64, 128, 192, 256):
            if bn is not None and bn > ((K + 63) // 64) * 64:
                continue
            bn_ = bn if bn is not None else raw.choose_block_n(O, K, mn_major=True)
            tiles = ((O + 127) // 128) * ((K + bn_ - 1) // bn_)
            kb = (M + 63) // 64
            for split in sorted({max(1, min(kb // 32, raw.num_sms() // max(tiles, 1))), max(1, min(kb // 8, -(-raw.num_sms() // max(tiles, 1)))), 1}):
                def f(dy, x, bn_=bn_, split=split):
                    raw.tapgemm(dy, x, g, M=O, N=K, K=M, a_mn=True, b_mn=True, split_k=split, out_dtype=raw.OUT_F32_ATOMIC, block_n=bn_,
                                lda=dy.stride(0), ldb=x.stride(0))
                t = graph_time([lambda dy=dy, x=x: f(dy, x) for dy, x in zip(dys, xs)], reps=5)
                res.append((t, bn_, split, bn is None))
        fl = 2.0 * M * O * K
        best = min(res)
        cur = [r for r in res if r[3]][0]
        print(f"  M={M:6d} O={O:5d} K={K:5d}: current bn={cur[1]} split={cur[2]} {cur[0]:7.1f} us ({fl / cur[0] / 1e6:5.0f} TF/s) | best bn={best[1]} split={best[2]} "
              f"{best[0]:7.1f} us ({fl / best[0] / 1e6:5.0f} TF/s) | all: " + " ".join(f"{b}/{s}:{t:.0f}" for t, b, s, _ in sorted(res, key=lambda r: (r[1], r[2]))), flush=True)


def sec_gemm():
    print("# forward / dgrad GEMMs and 3x3 convs: tile width (block_n) sweep, us per launch in a graph; * = raw.choose_block_n's pick")
    cases = [("conv 10x16 C1280", 14, 10, 16, 1280, 1280, 9), ("linear M2240 N1280 K1280", 14, 10, 16, 1280, 1280, 1), ("linear M2240 N1280 K5120", 14, 10, 16, 5120, 1280, 1),
             ("linear M2240 N5120 K1280", 14, 10, 16, 1280, 5120, 1), ("conv 20x32 C640", 14, 20, 32, 640, 640, 9), ("linear M8960 N640 K2560", 14, 20, 32, 2560, 640, 1),
             ("linear M8960 N640 K640", 14, 20, 32, 640, 640, 1), ("conv 40x64 C320", 14, 40, 64, 320, 320, 9), ("linear M35840 N320 K1280", 14, 40, 64, 1280, 320, 1),
             ("linear M35840 N320 K320", 14, 40, 64, 320, 320, 1), ("linear M35840 N960 K320", 14, 40, 64, 320, 960, 1),
             ("linear M35840 N320 K2560", 14, 40, 64, 2560, 320, 1), ("linear M35840 N320 K960", 14, 40, 64, 960, 320, 1),
             ("conv 40x64 640->320", 14, 40, 64, 640, 320, 9), ("linear M8960 N640 K5120", 14, 20, 32, 5120, 640, 1),
             ("linear M8960 N640 K1920", 14, 20, 32, 1920, 640, 1), ("linear M8960 N1920 K640", 14, 20, 32, 640, 1920, 1),
             ("linear M2240 N1280 K10240", 14, 10, 16, 10240, 1280, 1), ("linear M35840 N1280 K320", 14, 40, 64, 320, 1280, 1)]
    for name, T, H, W, K, N, taps in cases:
        M = T * H * W
        k = rot((M * K + M * N) * 2)
        xs = [torch.randn(M, K, device=DEV).to(bf16) for _ in range(k)]
        w = (torch.randn(N, taps * K, device=DEV) * (taps * K) ** -0.5).to(bf16)
        outs = [torch.empty(M, N, device=DEV, dtype=bf16) for _ in range(k)]
        bias = torch.randn(N, device=DEV)
        pick = raw.choose_block_n(M, N)
        res = []
        for bn in (320, 256, 160, 128, 96, 64):
            if N % bn or (bn < 128 and N > 640):
                continue

            def f(x, o, bn=bn):
                if taps == 9:
                    raw.tapgemm(x, w, o, M=M, N=N, K=K, mode=raw.A_CONV2D, taps=raw.CONV3x3_TAPS, conv_whn=(W, H, T), bias=bias, block_n=bn)
                else:
                    raw.tapgemm(x, w, o, M=M, N=N, K=K, bias=bias, block_n=bn)
            t = graph_time([lambda x=x, o=o: f(x, o) for x, o in zip(xs, outs)], reps=5)
            res.append((bn, t))
        fl = 2.0 * M * N * K * taps
        print(f"  {name:28s}: " + "  ".join(f"{'*' if bn == pick else ''}bn{bn}: {t:6.1f} us ({fl / t / 1e6:5.0f} TF/s)" for bn, t in res), flush=True)


def sec_gnb():
    print("# GroupNorm backward: pass 1 inside the dgrad epilogue (gnb) vs the separate partial kernel; us per launch in a graph")
    for name, T, H, W, C, taps in [("conv3x3 40x64 C320", 14, 40, 64, 320, 9), ("conv3x3 20x32 C640", 14, 20, 32, 640, 9),
                                   ("conv(3,1,1) 40x64 C320", 14, 40, 64, 320, 3), ("conv(3,1,1) 20x32 C640", 14, 20, 32, 640, 3),
                                   ("linear 40x64 C320", 14, 40, 64, 320, 1)]:
        M = T * H * W
        rows = H * W
        k = rot(M * C * 2 * 4)
        gs = [torch.randn(M, C, device=DEV).to(bf16) for _ in range(k)]          # upstream gradient (A operand of the dgrad)
        xs = [torch.randn(M, C, device=DEV).to(bf16) for _ in range(k)]          # GroupNorm input
        dys = [torch.empty(M, C, device=DEV, dtype=bf16) for _ in range(k)]
        dxs = [torch.empty(M, C, device=DEV, dtype=bf16) for _ in range(k)]
        w = (torch.randn(C, taps * C, device=DEV) * (taps * C) ** -0.5).to(bf16)
        gamma, beta = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
        st = [raw.groupnorm_stats(x, None, T, rows, 1e-5) for x in xs]
        abs_ = [torch.empty(T, 2, C, device=DEV) for _ in range(k)]
        for x, s_, ab, y in zip(xs, st, abs_, dys):
            raw.groupnorm_apply(x, None, T, rows, s_[0], s_[1], gamma, beta, True, y, ab=ab)
        sums = torch.zeros(T, 2, C, device=DEV)
        ws = torch.zeros(2 * T * 32, device=DEV)

        def dgrad(g, dy, gnb):
            kw = {} if gnb is None else {"gnb": gnb}
            if taps == 9:
                raw.tapgemm(g, w, dy, M=M, N=C, K=C, mode=raw.A_CONV2D, taps=raw.CONV3x3_TAPS, conv_whn=(W, H, T), **kw)
            elif taps == 3:
                raw.tapgemm(g, w, dy, M=M, N=C, K=C, taps=((-rows, 0, 0), (0, 0, 0), (rows, 0, 0)), rows_per_group=M, groups=1, **kw)
            else:
                raw.tapgemm(g, w, dy, M=M, N=C, K=C, **kw)
        t_plain = graph_time([lambda g=g, dy=dy: dgrad(g, dy, None) for g, dy in zip(gs, dys)], reps=5)
        t_gnb = graph_time([lambda g=g, dy=dy, x=x, ab=ab: dgrad(g, dy, dict(x=x, x2=None, ab=ab, rows=rows, silu=True, sum=sums))
                            for g, dy, x, ab in zip(gs, dys, xs, abs_)], reps=5)
        t_two = graph_time([lambda x=x, dy=dy, dx=dx, s_=s_: raw.groupnorm_bwd(x, None, dy, T, rows, s_[0], s_[1], gamma, beta, True, dx, None, ws=ws)
                            for x, dy, dx, s_ in zip(xs, dys, dxs, st)], reps=5)
        t_one = graph_time([lambda x=x, dy=dy, dx=dx, s_=s_: raw.groupnorm_bwd_fused(x, None, dy, T, rows, s_[0], s_[1], gamma, beta, True, sums, dx, None)
                            for x, dy, dx, s_ in zip(xs, dys, dxs, st)], reps=5)
        print(f"  {name:26s}: dgrad {t_plain:6.1f} us, with gnb sums {t_gnb:6.1f} us (+{t_gnb - t_plain:5.1f}) | GroupNorm bwd two kernels {t_two:6.1f} us, "
              f"fused one kernel {t_one:6.1f} us (-{t_two - t_one:5.1f}) | net {t_gnb - t_plain - (t_two - t_one):+6.1f} us", flush=True)


def sec_res():
    print("# GEMMs / convs with the residual epilogue (+res1, cold residual), us per launch in a graph, default tile width")
    for name, T, H, W, K, N, taps in [("linear M35840 N320 K320", 14, 40, 64, 320, 320, 1), ("linear M35840 N320 K1280", 14, 40, 64, 1280, 320, 1),
                                      ("conv 40x64 C320", 14, 40, 64, 320, 320, 9), ("linear M8960 N640 K640", 14, 20, 32, 640, 640, 1),
                                      ("linear M8960 N640 K2560", 14, 20, 32, 2560, 640, 1), ("conv 20x32 C640", 14, 20, 32, 640, 640, 9),
                                      ("linear M2240 N1280 K1280", 14, 10, 16, 1280, 1280, 1), ("linear M2240 N1280 K5120", 14, 10, 16, 5120, 1280, 1)]:
        M = T * H * W
        k = rot((M * K + 2 * M * N) * 2)
        xs = [torch.randn(M, K, device=DEV).to(bf16) for _ in range(k)]
        rs = [torch.randn(M, N, device=DEV).to(bf16) for _ in range(k)]
        w = (torch.randn(N, taps * K, device=DEV) * (taps * K) ** -0.5).to(bf16)
        outs = [torch.empty(M, N, device=DEV, dtype=bf16) for _ in range(k)]
        bias = torch.randn(N, device=DEV)

        def f(x, o, r):
            if taps == 9:
                raw.tapgemm(x, w, o, M=M, N=N, K=K, mode=raw.A_CONV2D, taps=raw.CONV3x3_TAPS, conv_whn=(W, H, T), bias=bias, res1=r)
            else:
                raw.tapgemm(x, w, o, M=M, N=N, K=K, bias=bias, res1=r)
        t0 = graph_time([lambda x=x, o=o: f(x, o, None) for x, o in zip(xs, outs)], reps=5)
        t1 = graph_time([lambda x=x, o=o, r=r: f(x, o, r) for x, o, r in zip(xs, outs, rs)], reps=5)
        print(f"  {name:28s}: plain {t0:6.1f} us | +res1 {t1:6.1f} us (+{t1 - t0:5.1f})", flush=True)


SECTIONS = {"gnb": sec_gnb, "res": sec_res, "gemm": sec_gemm, "gn": sec_gn, "ln": sec_ln, "attn": sec_attn, "wgrad": sec_wgrad}

if __name__ == "__main__":
    names = sys.argv[1:] or list(SECTIONS)
    t0 = time.time()
    for n in names:
        SECTIONS[n]()
    print(f"# done in {time.time() - t0:.1f} s")
