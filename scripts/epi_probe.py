"""Times individual tapgemm shapes with epilogue variants (GPU dev tool; not part of the product path)."""
import sys, torch
sys.path.insert(0, ".")
from svd_xtend_b200 import raw

dev = torch.device("cuda:0")
bf = torch.bfloat16


def timeit(fn, iters=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters * 1e3  # us


def run(M, N, K, geglu=False, bias=True, pre=False, res=False, block_n=None, f32=False, label="", res2=False):
    a = torch.randn(M, K, device=dev, dtype=bf)
    w = torch.randn(N, K, device=dev, dtype=bf) * 0.05
    n_out = N // 2 if geglu else N
    out = torch.empty(M, n_out, device=dev, dtype=torch.float32 if f32 else bf)
    b = torch.randn(N, device=dev) if bias else None
    p = torch.empty(M, N, device=dev, dtype=bf) if pre else None
    r = torch.randn(M, n_out, device=dev, dtype=bf) if res else None
    r2 = torch.randn(M, n_out, device=dev, dtype=bf) if res2 else None
    sc = torch.tensor([0.5, 1.0, 0.5, 0.0], device=dev) if res2 else None
    us = timeit(lambda: raw.tapgemm(a, w, out, M=M, N=N, K=K, geglu=geglu, bias=b, pre=p, res1=r, res2=r2, scales=sc, block_n=block_n))
    tf = 2.0 * M * N * K / us / 1e6
    print(f"{label:28s} M={M} N={N} K={K} geglu={int(geglu)} pre={int(pre)} res={int(res)} bn={block_n} : {us:8.1f} us  {tf:6.1f} TF", flush=True)


def runconv(W, H, nimg, C, N, block_n=None, auto=False, label=""):
    M = W * H * nimg
    a = torch.randn(M, C, device=dev, dtype=bf)
    w = torch.randn(N, 9 * C, device=dev, dtype=bf) * 0.02
    out = torch.empty(M, N, device=dev, dtype=bf)
    b = torch.randn(N, device=dev)
    fn = raw.tapgemm_auto if auto else raw.tapgemm
    us = timeit(lambda: fn(a, w, out, M=M, N=N, K=C, mode=raw.A_CONV2D, taps=raw.CONV3x3_TAPS, conv_whn=(W, H, nimg), bias=b,
                           block_n=block_n))
    tf = 2.0 * M * N * C * 9 / us / 1e6
    print(f"{label:22s} conv3x3 W={W} H={H} n={nimg} C={C} N={N} bn={block_n} auto={int(auto)} : {us:8.1f} us  {tf:6.1f} TF", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "conv":
        for bn in (None, 256, 160, 128):
            runconv(16, 10, 14, 1280, 1280, block_n=bn, label="L2 conv")
        runconv(16, 10, 14, 1280, 1280, auto=True, label="L2 conv auto")
        for bn in (None, 256, 160, 128):
            runconv(8, 5, 14, 1280, 1280, block_n=bn, label="L3 conv")
        runconv(8, 5, 14, 1280, 1280, auto=True, label="L3 conv auto")
        for bn in (None, 160, 128):
            runconv(32, 20, 14, 640, 640, block_n=bn, label="L1 conv")
        for bn in (None, 160):
            runconv(64, 40, 14, 320, 320, block_n=bn, label="L0 conv")
        run(2240, 1280, 11520, label="L2 as plain GEMM")
        run(2240, 1280, 11520, block_n=160, label="L2 as plain GEMM bn160")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "convone":
        runconv(16, 10, 14, 1280, 1280, label="L2 conv")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "res":
        for K in (320, 1280, 2560):
            run(35840, 320, K, label="L0 plain")
            run(35840, 320, K, res=True, label="L0 +res")
            run(35840, 320, K, res=True, res2=True, label="L0 +res+blend")
        run(8960, 640, 2560, label="L1 plain")
        run(8960, 640, 2560, res=True, label="L1 +res")
        run(2240, 1280, 5120, label="L2 plain")
        run(2240, 1280, 5120, res=True, label="L2 +res")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        run(35840, 2560, 320, bias=False, label="plain N=2560 nobias")
        run(35840, 2560, 320, geglu=True, pre=True, label="geglu+pre")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "short":
        run(35840, 2560, 320, label="plain N=2560")
        run(35840, 2560, 320, bias=False, label="plain N=2560 nobias")
        run(35840, 2560, 320, f32=True, label="plain N=2560 f32 out")
        run(35840, 320, 320, label="proj 320")
        run(35840, 320, 2560, label="ff2 L0")
        run(8960, 5120, 640, label="plain L1")
        run(8960, 640, 5120, label="ff2 L1")
        sys.exit(0)
    for bn in (None, 128, 256):
        run(35840, 2560, 320, geglu=True, pre=True, block_n=bn, label="geglu+pre")
        run(35840, 2560, 320, geglu=True, pre=False, block_n=bn, label="geglu")
        run(35840, 2560, 320, geglu=False, block_n=bn, label="plain N=2560")
    run(35840, 1280, 320, label="plain N=1280")
    run(8960, 5120, 640, geglu=True, pre=True, label="geglu+pre L1")
    run(8960, 5120, 640, geglu=True, pre=False, label="geglu L1")
    run(8960, 5120, 640, label="plain L1")
    run(35840, 320, 320, label="proj 320")
    run(35840, 320, 320, res=True, label="proj 320 + res")
    run(35840, 320, 320, bias=False, label="proj 320 nobias")
    run(8960, 640, 640, label="proj 640")
    run(2240, 1280, 1280, label="proj 1280")
    run(35840, 320, 2560, label="ff2 L0")
    run(35840, 960, 320, label="qkv L0")
