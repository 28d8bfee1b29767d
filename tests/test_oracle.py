"""CPU tests pinning the oracle (oracle/svd_unet_oracle.py) — SURVEY.md §8c acceptance checks.

Parity against diffusers itself is UNPINNED (no diffusers here, no reference tests exist); these checks pin
what can be pinned without it: exact parameter census, state-dict key contract, the 1-token cross-attention
identity, precision self-consistency and the committed golden vectors (tests/golden/).
"""
import os
import re

import pytest
import torch

from oracle.svd_unet_oracle import (SVD_CONFIG, TINY_CONFIG, Attention, UNetSpatioTemporalConditionModel, edm_loss,
                                    get_timestep_embedding, synthetic_batch)

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "tiny_oracle.pt")


def test_parameter_census_svd_config():
    with torch.device("meta"):
        m = UNetSpatioTemporalConditionModel(**SVD_CONFIG)
    total = sum(p.numel() for p in m.parameters())
    temporal = sum(p.numel() for n, p in m.named_parameters() if "temporal_transformer_block" in n)
    assert total == 1_524_623_082            # SURVEY.md Appendix A census
    assert temporal == 397_620_480           # the train_svd.py:761-766 trainable set
    n_res = sum(1 for n, _ in m.named_modules() if n.endswith("spatial_res_block"))
    n_tr = sum(1 for n, _ in m.named_modules() if n.endswith("temporal_transformer_blocks"))
    assert (n_res, n_tr) == (22, 16)


def test_state_dict_key_contract():
    """SURVEY.md Appendix C patterns cover every key."""
    with torch.device("meta"):
        m = UNetSpatioTemporalConditionModel(**SVD_CONFIG)
    blk = r"(down_blocks\.\d|mid_block|up_blocks\.\d)"
    pats = [
        r"conv_in\.(weight|bias)", r"(time|add)_embedding\.linear_[12]\.(weight|bias)",
        blk + r"\.resnets\.\d\.spatial_res_block\.(norm1|conv1|time_emb_proj|norm2|conv2|conv_shortcut)\.(weight|bias)",
        blk + r"\.resnets\.\d\.temporal_res_block\.(norm1|conv1|time_emb_proj|norm2|conv2)\.(weight|bias)",
        blk + r"\.resnets\.\d\.time_mixer\.mix_factor",
        blk + r"\.attentions\.\d\.(norm|proj_in|proj_out)\.(weight|bias)",
        blk + r"\.attentions\.\d\.transformer_blocks\.0\.(norm1|norm2|norm3)\.(weight|bias)",
        blk + r"\.attentions\.\d\.transformer_blocks\.0\.(attn1|attn2)\.(to_q|to_k|to_v)\.weight",
        blk + r"\.attentions\.\d\.transformer_blocks\.0\.(attn1|attn2)\.to_out\.0\.(weight|bias)",
        blk + r"\.attentions\.\d\.transformer_blocks\.0\.ff\.net\.(0\.proj|2)\.(weight|bias)",
        blk + r"\.attentions\.\d\.temporal_transformer_blocks\.0\.(norm_in|norm1|norm2|norm3)\.(weight|bias)",
        blk + r"\.attentions\.\d\.temporal_transformer_blocks\.0\.(ff_in|ff)\.net\.(0\.proj|2)\.(weight|bias)",
        blk + r"\.attentions\.\d\.temporal_transformer_blocks\.0\.(attn1|attn2)\.(to_q|to_k|to_v)\.weight",
        blk + r"\.attentions\.\d\.temporal_transformer_blocks\.0\.(attn1|attn2)\.to_out\.0\.(weight|bias)",
        blk + r"\.attentions\.\d\.time_pos_embed\.linear_[12]\.(weight|bias)",
        blk + r"\.attentions\.\d\.time_mixer\.mix_factor",
        r"down_blocks\.\d\.downsamplers\.0\.conv\.(weight|bias)", r"up_blocks\.\d\.upsamplers\.0\.conv\.(weight|bias)",
        r"conv_norm_out\.(weight|bias)", r"conv_out\.(weight|bias)",
    ]
    rx = [re.compile("^" + p + "$") for p in pats]
    keys = list(m.state_dict().keys())
    assert len(keys) == len(set(keys))
    bad = [k for k in keys if not any(r.match(k) for r in rx)]
    assert not bad, bad[:5]
    sd = m.state_dict()
    assert tuple(sd["down_blocks.0.resnets.0.temporal_res_block.conv1.weight"].shape) == (320, 320, 3, 1, 1)
    assert tuple(sd["up_blocks.1.resnets.0.spatial_res_block.conv1.weight"].shape) == (1280, 2560, 3, 3)
    assert tuple(sd["down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k.weight"].shape) == (320, 1024)


def test_timestep_embedding_layout():
    t = torch.tensor([0.0, 1.5, 300.0])
    e = get_timestep_embedding(t, 8, flip_sin_to_cos=True, downscale_freq_shift=0)
    f = torch.exp(-torch.log(torch.tensor(10000.0)) * torch.arange(4) / 4)
    assert torch.allclose(e[:, :4], torch.cos(t[:, None] * f)) and torch.allclose(e[:, 4:], torch.sin(t[:, None] * f))


def test_single_token_cross_attention_identity():
    """with one key/value token, attention == to_out(to_v(e)) broadcast over queries."""
    torch.manual_seed(0)
    att = Attention(query_dim=128, cross_attention_dim=64, heads=2, dim_head=64)
    x = torch.randn(3, 50, 128)
    e = torch.randn(3, 1, 64)
    out = att(x, encoder_hidden_states=e)
    ref = att.to_out[0](att.to_v(e)).expand(3, 50, 128)
    assert torch.allclose(out, ref, atol=1e-5)
    out.sum().backward()
    assert att.to_q.weight.grad.abs().max() < 1e-6 and att.to_k.weight.grad.abs().max() < 1e-6


def _tiny(dtype):
    torch.manual_seed(20260922)
    return UNetSpatioTemporalConditionModel(**TINY_CONFIG).to(dtype)


def test_golden_vectors_tiny_config():
    gold = torch.load(GOLDEN, weights_only=False)
    torch.set_num_threads(4)
    model = _tiny(torch.float64)
    assert abs(float(sum(p.double().abs().sum() for p in model.parameters())) - gold["param_checksum"]) < 1e-6 * gold["param_checksum"]
    batch = synthetic_batch(1, 4, 16, 16, seed=1234, cross_dim=TINY_CONFIG["cross_attention_dim"], dtype=torch.float64)
    model.requires_grad_(False)
    for n, p in model.named_parameters():
        if "temporal_transformer_block" in n:
            p.requires_grad_(True)
    pred = model(batch["sample"], batch["timestep"].double(), batch["encoder_hidden_states"], batch["added_time_ids"]).sample
    loss = edm_loss(pred, batch["noisy"], batch["latents"], batch["sigmas"].double())
    loss.backward()
    assert torch.allclose(pred.float(), gold["pred"], atol=1e-5, rtol=1e-4)
    assert abs(float(loss.detach()) - gold["loss"]) < 1e-5
    grads = dict((n, p.grad) for n, p in model.named_parameters() if p.requires_grad)
    for n, g in gold["grads"].items():
        assert torch.allclose(grads[n].float(), g, atol=1e-6, rtol=1e-3), n
    for n, v in gold["grad_norms"].items():
        assert abs(float(grads[n].norm()) - v) <= 1e-4 * max(v, 1e-6) + 1e-9, n
    # quirk 3: these trainable parameters get exactly zero gradient
    zero = [n for n in grads if re.search(r"attn2\.to_(q|k)\.weight|temporal_transformer_blocks\.0\.norm2", n)]
    assert zero and all(float(grads[n].abs().max()) == 0.0 for n in zero)


def test_fp32_matches_fp64():
    m64 = _tiny(torch.float64)
    m32 = _tiny(torch.float32)
    b = synthetic_batch(1, 4, 16, 16, seed=5, cross_dim=TINY_CONFIG["cross_attention_dim"], dtype=torch.float64)
    with torch.no_grad():
        p64 = m64(b["sample"], b["timestep"].double(), b["encoder_hidden_states"], b["added_time_ids"]).sample
        p32 = m32(b["sample"].float(), b["timestep"], b["encoder_hidden_states"].float(), b["added_time_ids"].float()).sample
    rel = ((p32.double() - p64).norm() / p64.norm()).item()
    assert rel < 1e-5, rel


def test_gradient_checkpointing_equivalence():
    m = _tiny(torch.float32)
    m.train()
    b = synthetic_batch(1, 4, 16, 16, seed=9, cross_dim=TINY_CONFIG["cross_attention_dim"])

    def run():
        m.zero_grad(set_to_none=True)
        pred = m(b["sample"], b["timestep"], b["encoder_hidden_states"], b["added_time_ids"]).sample
        edm_loss(pred, b["noisy"], b["latents"], b["sigmas"]).backward()
        return pred.detach().clone(), m.conv_in.weight.grad.clone()

    p0, g0 = run()
    m.enable_gradient_checkpointing()
    p1, g1 = run()
    assert torch.allclose(p0, p1, atol=1e-6) and torch.allclose(g0, g1, atol=1e-6, rtol=1e-4)


def test_constructor_errors():
    with pytest.raises(ValueError):
        UNetSpatioTemporalConditionModel(**{**TINY_CONFIG, "up_block_types": ("UpBlockSpatioTemporal",)})
    with pytest.raises(ValueError):
        UNetSpatioTemporalConditionModel(**{**TINY_CONFIG, "block_out_channels": (64,)})
    with pytest.raises(ValueError):
        UNetSpatioTemporalConditionModel(**{**TINY_CONFIG, "num_attention_heads": (1, 2, 3)})
