"""Generates tests/golden/ref_wiring_tiny.pt by running the REFERENCE'S OWN top-level file.

    python tests/golden/make_ref_wiring_golden.py          (needs /root/reference; run in the build container)

/root/reference/src/unet_spatio_temporal_condition.py is imported unmodified from where it lies (oracle/ref_wiring.py
provides a stand-in for the absent `diffusers` package whose block classes are the oracle's), instantiated with the
small test topology and run in fp64 on a batch of TWO clips. The fixture freezes what the reference's file itself
decides: state-dict names and shapes, channel bookkeeping of the down / mid / up blocks, the order of embedding
repeats and skip connections in forward(), the attention-processor key set — so the oracle's top-level restatement
(and through it the CUDA path) is pinned to reference code, not to a recollection of it. The block arithmetic is the
oracle's on both sides and stays unpinned against diffusers (see the header of oracle/svd_unet_oracle.py).
"""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_wiring import load_reference_unet_class  # noqa: E402
from oracle.svd_unet_oracle import SVD_CONFIG, TINY_CONFIG, synthetic_batch  # noqa: E402

SEED = 20260923


def build(cls):
    torch.manual_seed(SEED)
    m = cls(**TINY_CONFIG).double()
    with torch.no_grad():      # de-symmetrise what default init leaves at constants, so that every affine / blend path matters
        g = torch.Generator().manual_seed(SEED + 1)
        for n, p in m.named_parameters():
            if "norm" in n or n.endswith("bias") or n.endswith("mix_factor"):
                p.add_(0.1 * torch.randn(p.shape, generator=g, dtype=p.dtype))
    return m.eval()


def batch():
    return synthetic_batch(2, 4, 16, 16, seed=4321, cross_dim=TINY_CONFIG["cross_attention_dim"], dtype=torch.float64)


def main():
    Ref = load_reference_unet_class("/root/reference")
    ref = build(Ref)
    b = batch()
    with torch.no_grad():
        out = ref(b["sample"], b["timestep"].double(), b["encoder_hidden_states"], b["added_time_ids"]).sample
        out_t = ref(b["sample"], b["timestep"].double(), b["encoder_hidden_states"], b["added_time_ids"], return_dict=False)[0]
    assert torch.equal(out, out_t)
    with torch.device("meta"):
        big = Ref(**SVD_CONFIG)
    big_keys = [(k, tuple(v.shape)) for k, v in big.state_dict().items()]
    fixture = {
        "seed": SEED,
        "reference_file_sha256": hashlib.sha256(open("/root/reference/src/unet_spatio_temporal_condition.py", "rb").read()).hexdigest(),
        "tiny_keys": [(k, tuple(v.shape)) for k, v in ref.state_dict().items()],
        "tiny_param_checksum": float(sum(p.double().abs().sum() for p in ref.parameters())),
        "tiny_out": out.clone(),
        "tiny_attn_processor_keys": sorted(ref.attn_processors.keys()),
        "svd_keys_sha256": hashlib.sha256(repr(big_keys).encode()).hexdigest(),
        "svd_num_keys": len(big_keys),
        "svd_total_params": sum(v.numel() for v in big.state_dict().values()),
        "svd_temporal_params": sum(p.numel() for n, p in big.named_parameters() if "temporal_transformer_block" in n),
        "svd_num_upsamplers": big.num_upsamplers,
        "svd_attn_processors": len(big.attn_processors),
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_wiring_tiny.pt")
    torch.save(fixture, path)
    print("wrote", path, "out std", float(out.std()), "svd params", fixture["svd_total_params"], fixture["svd_temporal_params"],
          "processors", fixture["svd_attn_processors"])


if __name__ == "__main__":
    main()
