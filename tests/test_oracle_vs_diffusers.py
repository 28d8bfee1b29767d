"""A/B of the oracle against the REAL diffusers blocks — runs wherever `diffusers` is importable, skips loudly elsewhere.

The build container and the GPU boxes of this project have no diffusers (not installable offline), so here this test SKIPS
and the oracle's block arithmetic stays "parity unpinned" (oracle/svd_unet_oracle.py header). On any box
that has diffusers >= 0.29.1 (the window of train_svd_lora.py:63) it pins the oracle the moment it is run:

    python -m pytest tests/test_oracle_vs_diffusers.py -q            # config 1 of BASELINE.json on the small topology
    SVDX_WRITE_DIFFUSERS_GOLDEN=1 python -m pytest tests/test_oracle_vs_diffusers.py -q     # also writes tests/golden/diffusers_tiny.pt

Config 1 (single UNet forward, CPU fp32, tolerance 1e-5 rel-L2): same seeded state dict loaded into
diffusers' UNetSpatioTemporalConditionModel and into the oracle, same inputs (batch of TWO clips: the `time_context`
broadcast order of TransformerSpatioTemporalModel only shows at B > 1), outputs and the as-scripted parameter gradients
compared. Also checks the items SURVEY.md Appendix D lists as recalled from memory (GroupNorm eps per block type)."""
import os

import pytest
import torch

diffusers = pytest.importorskip("diffusers", reason="diffusers is not installed on this box: the oracle cannot be A/B-ed against it here "
                                                    "(parity of the block arithmetic stays UNPINNED; see DESIGN.md §4)")


def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _models():
    from diffusers import UNetSpatioTemporalConditionModel as Real
    from oracle.svd_unet_oracle import TINY_CONFIG, UNetSpatioTemporalConditionModel as Oracle
    torch.manual_seed(20260923)
    ora = Oracle(**TINY_CONFIG)
    with torch.no_grad():
        for n, p in ora.named_parameters():
            if "norm" in n or n.endswith("bias") or n.endswith("mix_factor"):
                p.add_(0.1 * torch.randn_like(p))
    real = Real(**TINY_CONFIG)
    missing, unexpected = real.load_state_dict(ora.state_dict(), strict=False)
    assert not missing and not unexpected, (missing[:5], unexpected[:5])
    return real, ora


def test_config1_forward_and_gradients_match_diffusers():
    from oracle.svd_unet_oracle import TINY_CONFIG, edm_loss, synthetic_batch
    real, ora = _models()
    b = synthetic_batch(2, 4, 16, 16, seed=4321, cross_dim=TINY_CONFIG["cross_attention_dim"])
    outs, grads = [], []
    for m in (real, ora):
        m.train()
        m.requires_grad_(False)
        for n, p in m.named_parameters():
            if "temporal_transformer_block" in n:      # train_svd.py:761-766
                p.requires_grad_(True)
        pred = m(b["sample"], b["timestep"], b["encoder_hidden_states"], added_time_ids=b["added_time_ids"]).sample
        edm_loss(pred, b["noisy"], b["latents"], b["sigmas"]).backward()
        outs.append(pred.detach())
        grads.append({n: p.grad.clone() for n, p in m.named_parameters() if p.requires_grad})
    e = _rel(outs[1], outs[0])
    print("oracle vs diffusers", diffusers.__version__, "output rel-l2", e)
    assert e < 1e-5, e
    for n, g in grads[0].items():
        if g.abs().max() == 0:
            assert grads[1][n].abs().max() == 0, n
        else:
            assert _rel(grads[1][n], g) < 1e-4, (n, _rel(grads[1][n], g))
    if os.environ.get("SVDX_WRITE_DIFFUSERS_GOLDEN"):
        path = os.path.join(os.path.dirname(__file__), "golden", "diffusers_tiny.pt")
        torch.save({"diffusers": diffusers.__version__, "pred": outs[0], "grad_norms": {n: float(g.norm()) for n, g in grads[0].items()}}, path)


def test_appendix_d_groupnorm_eps_per_block_type():
    real, ora = _models()
    eps_real = {n: m.eps for n, m in real.named_modules() if isinstance(m, torch.nn.GroupNorm)}
    eps_ora = {n: m.eps for n, m in ora.named_modules() if isinstance(m, torch.nn.GroupNorm)}
    assert eps_real == eps_ora
    ln_real = {n: m.eps for n, m in real.named_modules() if isinstance(m, torch.nn.LayerNorm)}
    ln_ora = {n: m.eps for n, m in ora.named_modules() if isinstance(m, torch.nn.LayerNorm)}
    assert ln_real == ln_ora
