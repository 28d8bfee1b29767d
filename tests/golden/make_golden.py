"""Generates tests/golden/tiny_oracle.pt — golden vectors of the ORACLE (oracle/svd_unet_oracle.py).

The reference (pixeli99/SVD_Xtend) ships no tests or golden tensors and its arithmetic lives in the
un-vendored diffusers dependency, which is not installable here (no network): parity against diffusers is
UNPINNED. These vectors freeze the oracle's own behaviour (fp32, CPU, seeded default init) so that any
later edit of the oracle — or a future A/B against a real diffusers install — is detected.

    python tests/golden/make_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.svd_unet_oracle import TINY_CONFIG, UNetSpatioTemporalConditionModel, edm_loss, synthetic_batch  # noqa: E402


def main():
    torch.manual_seed(20260922)
    torch.set_num_threads(4)
    model = UNetSpatioTemporalConditionModel(**TINY_CONFIG).double()
    batch = synthetic_batch(1, 4, 16, 16, seed=1234, cross_dim=TINY_CONFIG["cross_attention_dim"], dtype=torch.float64)
    model.requires_grad_(False)
    names = []
    for n, p in model.named_parameters():
        if "temporal_transformer_block" in n:  # train_svd.py:761-766
            p.requires_grad_(True)
            names.append(n)
    pred = model(batch["sample"], batch["timestep"].double(), batch["encoder_hidden_states"], batch["added_time_ids"]).sample
    loss = edm_loss(pred, batch["noisy"], batch["latents"], batch["sigmas"].double())
    loss.backward()
    grads = {n: p.grad for n, p in model.named_parameters() if p.requires_grad}
    keep = ["down_blocks.0.attentions.0.temporal_transformer_blocks.0.ff_in.net.0.proj.weight",
            "down_blocks.0.attentions.0.temporal_transformer_blocks.0.attn1.to_q.weight",
            "down_blocks.0.attentions.0.temporal_transformer_blocks.0.attn2.to_v.weight",
            "up_blocks.1.attentions.0.temporal_transformer_blocks.0.norm_in.weight"]
    out = {
        "seed": 20260922,
        "config": TINY_CONFIG,
        "pred": pred.float(),
        "loss": float(loss),
        "grad_norms": {n: float(g.norm()) for n, g in grads.items()},
        "grads": {n: grads[n].float() for n in keep},
        "param_checksum": float(sum(p.double().abs().sum() for p in model.parameters())),
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny_oracle.pt")
    torch.save(out, path)
    print("wrote", path, "loss", float(loss), "pred std", float(pred.std()))


if __name__ == "__main__":
    main()
