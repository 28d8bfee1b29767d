"""Times the attention kernels on the SVD shapes (GPU dev tool)."""
import sys, torch
sys.path.insert(0, ".")
from svd_xtend_b200 import raw

dev = torch.device("cuda:0")
bf = torch.bfloat16


def timeit(fn, iters=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters * 1e3


def spatial(S, heads, nseq, label):
    C = heads * 64
    tok = S * nseq
    qkv = torch.randn(tok, 3 * C, device=dev, dtype=bf)
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    o = torch.empty(tok, C, device=dev, dtype=bf)
    lse = torch.empty(tok, heads, device=dev, dtype=torch.float32)
    f = timeit(lambda: raw.attention_fwd(q, k, v, o, heads=heads, S=S, nseq=nseq, lse=lse))
    do = torch.randn(tok, C, device=dev, dtype=bf)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    b = timeit(lambda: raw.attention_bwd(q, k, v, o, do, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], lse, delta, heads=heads, S=S, nseq=nseq))
    fl = 4.0 * S * S * 64 * heads * nseq
    print(f"{label:10s} S={S} heads={heads} nseq={nseq}: fwd {f:7.1f} us ({fl / f / 1e6:6.1f} TF)  bwd(delta+dq+dkv) {b:7.1f} us ({2.5 * fl / b / 1e6:6.1f} TF)", flush=True)


if __name__ == "__main__":
    spatial(2560, 5, 14, "L0")
    spatial(640, 10, 14, "L1")
    spatial(160, 20, 14, "L2")
