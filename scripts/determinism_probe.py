"""Run-to-run reproducibility of the forward: where do two runs of the same kernels on the same inputs differ?
Sources by construction: fp32 atomics (GroupNorm statistics, split-K accumulation) whose order is not fixed. This probe runs
the tiny topology forward repeatedly with (a) everything as shipped, (b) GroupNorm statistics replaced by a deterministic torch
computation, (c) additionally split-K disabled, and reports the pairwise rel-L2 differences. If (c) is bit-identical the
noise of (a) is atomics ordering amplified by bf16 re-rounding through the network, not a race."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svd_xtend_b200 import raw  # noqa: E402
from svd_xtend_b200.unet import UNetSpatioTemporalConditionModel  # noqa: E402
from svd_xtend_b200.workload import synthetic_batch  # noqa: E402

TINY = dict(sample_size=None, in_channels=8, out_channels=4,
            down_block_types=("CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal"),
            up_block_types=("UpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal"),
            block_out_channels=(64, 128), addition_time_embed_dim=32, projection_class_embeddings_input_dim=96, layers_per_block=1,
            cross_attention_dim=64, transformer_layers_per_block=1, num_attention_heads=(1, 2), num_frames=4)


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def main():
    dev = "cuda:0"
    torch.manual_seed(0)
    m = UNetSpatioTemporalConditionModel(**TINY).to(dev).eval()
    b = synthetic_batch(2, 4, 16, 16, seed=3, device=dev, cross_dim=64)

    def fwd():
        with torch.no_grad():
            return m(b["sample"], b["timestep"], b["encoder_hidden_states"], b["added_time_ids"]).sample.clone()

    def report(tag):
        outs = [fwd() for _ in range(4)]
        torch.cuda.synchronize()
        print(f"{tag:55s} pairwise rel-l2 vs run 0:", " ".join(f"{rel(o, outs[0]):.3e}" for o in outs[1:]), flush=True)

    report("(a) as shipped")
    orig_stats = raw.groupnorm_stats

    def det_stats(x, x2, outer, rows, eps, groups=32):
        xx = x if x2 is None else torch.cat([x, x2], dim=-1)
        C = xx.shape[-1]
        v = xx.float().reshape(outer, rows, groups, C // groups).permute(0, 2, 1, 3).reshape(outer * groups, -1)
        mean = v.mean(dim=1)
        var = (v * v).mean(dim=1) - mean * mean
        return mean.contiguous(), torch.rsqrt(var.clamp_min(0) + eps).contiguous()

    raw.groupnorm_stats = det_stats
    report("(b) deterministic GroupNorm statistics")
    orig_sms = raw.num_sms
    raw.num_sms = lambda: 1          # tapgemm_auto: split = min(sms // tiles, ...) < 2 -> never splits
    report("(c) + split-K disabled")
    raw.groupnorm_stats, raw.num_sms = orig_stats, orig_sms


if __name__ == "__main__":
    main()
