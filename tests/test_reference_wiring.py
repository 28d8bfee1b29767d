"""The oracle's (and the product class's) top-level wiring against the REFERENCE'S OWN file.

tests/golden/ref_wiring_tiny.pt was produced by importing /root/reference/src/unet_spatio_temporal_condition.py unmodified
(tests/golden/make_ref_wiring_golden.py, through oracle/ref_wiring.py) and running it in fp64 on two clips. These CPU tests
need only the committed fixture; where /root/reference exists (the build container) the live A/B runs as well.
What is pinned: construction (:71-246), forward wiring (:357-490), plugin-API key set (:248-274), parameter census of the SVD
configuration. The arithmetic inside the blocks is the oracle's on both sides (unpinned against diffusers)."""
import hashlib
import importlib.util
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_wiring_tiny.pt")


def _gen():
    spec = importlib.util.spec_from_file_location("make_ref_wiring_golden", os.path.join(ROOT, "tests", "golden", "make_ref_wiring_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLDEN, weights_only=False)


def test_oracle_forward_equals_reference_file_golden(gold):
    from oracle.svd_unet_oracle import UNetSpatioTemporalConditionModel as Oracle
    gen = _gen()
    m = gen.build(Oracle)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == gold["tiny_keys"]
    chk = float(sum(p.detach().double().abs().sum() for p in m.parameters()))
    assert abs(chk - gold["tiny_param_checksum"]) < 1e-9 * gold["tiny_param_checksum"], "seeded init differs from the generator's"
    b = gen.batch()
    with torch.no_grad():
        out = m(b["sample"], b["timestep"].double(), b["encoder_hidden_states"], b["added_time_ids"]).sample
    assert out.shape == gold["tiny_out"].shape == (2, 4, 4, 16, 16)
    err = ((out - gold["tiny_out"]).norm() / gold["tiny_out"].norm()).item()
    assert err < 1e-12, err      # identical blocks, identical wiring: fp64 round-off only
    assert sorted(m_name + ".processor" for m_name, mod in m.named_modules() if hasattr(mod, "get_processor")) == gold["tiny_attn_processor_keys"]


def test_svd_configuration_census_equals_reference_file(gold):
    """names, shapes and counts the reference's constructor produces for the SVD configuration: oracle AND product class"""
    from oracle.svd_unet_oracle import SVD_CONFIG, UNetSpatioTemporalConditionModel as Oracle
    from svd_xtend_b200.unet import UNetSpatioTemporalConditionModel as Ours
    assert gold["svd_total_params"] == 1_524_623_082 and gold["svd_temporal_params"] == 397_620_480
    for cls in (Oracle, Ours):
        with torch.device("meta"):
            m = cls(**SVD_CONFIG)
        keys = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
        assert len(keys) == gold["svd_num_keys"]
        assert hashlib.sha256(repr(keys).encode()).hexdigest() == gold["svd_keys_sha256"], cls
        assert m.num_upsamplers == gold["svd_num_upsamplers"]
    assert len(m.attn_processors) == gold["svd_attn_processors"] == 64


def test_product_class_plugin_keys_equal_reference_file(gold):
    from oracle.svd_unet_oracle import TINY_CONFIG
    from svd_xtend_b200.unet import UNetSpatioTemporalConditionModel as Ours
    m = Ours(**TINY_CONFIG)
    assert sorted(m.attn_processors.keys()) == gold["tiny_attn_processor_keys"]
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == gold["tiny_keys"]


@pytest.mark.skipif(not os.path.exists("/root/reference/src/unet_spatio_temporal_condition.py"), reason="the reference checkout is not on this box")
def test_live_reference_file_vs_oracle(gold):
    """build container only: re-run the reference file itself and compare with the oracle and with the fixture"""
    from oracle.ref_wiring import load_reference_unet_class
    from oracle.svd_unet_oracle import UNetSpatioTemporalConditionModel as Oracle
    sha = hashlib.sha256(open("/root/reference/src/unet_spatio_temporal_condition.py", "rb").read()).hexdigest()
    assert sha == gold["reference_file_sha256"], "the fixture was generated from a different reference file"
    gen = _gen()
    Ref = load_reference_unet_class("/root/reference")
    ref, ora = gen.build(Ref), gen.build(Oracle)
    for (ka, va), (kb, vb) in zip(ref.state_dict().items(), ora.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    b = gen.batch()
    ref.train(), ora.train()
    for m in (ref, ora):
        m.requires_grad_(False)
        for n, p in m.named_parameters():
            if "temporal_transformer_block" in n:
                p.requires_grad_(True)
    outs = []
    for m in (ref, ora):
        out = m(b["sample"], b["timestep"].double(), b["encoder_hidden_states"], b["added_time_ids"]).sample
        out.square().mean().backward()
        outs.append(out.detach())
    assert torch.equal(outs[0], outs[1])
    assert ((outs[0] - gold["tiny_out"]).norm() / gold["tiny_out"].norm()).item() < 1e-12
    for (n, p), (_, q) in zip(ref.named_parameters(), ora.named_parameters()):
        if p.requires_grad:
            assert torch.equal(p.grad, q.grad), n
    # the plugin API of the reference file on the oracle's Attention modules
    ref.set_default_attn_processor()
    assert {type(p).__name__ for p in ref.attn_processors.values()} == {"AttnProcessor"}
    with pytest.raises(ValueError):
        ref.set_attn_processor({})
    ref.enable_gradient_checkpointing()
    assert sum(bool(getattr(m, "gradient_checkpointing", False)) for m in ref.modules()) == \
        sum(1 for m in ora.modules() if hasattr(m, "gradient_checkpointing"))
