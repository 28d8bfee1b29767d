"""The caller's side of the hot path, as train_svd.py runs it: model configuration, synthetic clip batches of the
BASELINE.json configs and the EDM loss.  Product-side helpers (bench.py, GraphedStep users, smoke()); the oracle has
its own, independent statements of the same formulas (oracle/svd_unet_oracle.py) which the tests compare with.

References: /root/reference/train_svd.py:951-1017 (batch assembly), :1025-1036 (loss),
/root/reference/src/unet_spatio_temporal_condition.py:71-97 (default configuration).
"""
from __future__ import annotations

import torch

# src/unet_spatio_temporal_condition.py:71-97 — the defaults ARE the SVD img2vid configuration
SVD_CONFIG = dict(
    sample_size=None, in_channels=8, out_channels=4,
    down_block_types=("CrossAttnDownBlockSpatioTemporal",) * 3 + ("DownBlockSpatioTemporal",),
    up_block_types=("UpBlockSpatioTemporal",) + ("CrossAttnUpBlockSpatioTemporal",) * 3,
    block_out_channels=(320, 640, 1280, 1280), addition_time_embed_dim=256,
    projection_class_embeddings_input_dim=768, layers_per_block=2, cross_attention_dim=1024,
    transformer_layers_per_block=1, num_attention_heads=(5, 10, 20, 20), num_frames=25,
)

# BASELINE.json configs: (frames, latent H, latent W, gradient checkpointing, LoRA rank or None)
BENCH_CONFIGS = {
    2: dict(frames=14, h=40, w=64, grad_ckpt=False, lora_rank=None,
            name="train_svd.py full-finetune step as scripted (trainable = *temporal_transformer_block* params, "
                 "train_svd.py:761-766), bs=1/GPU, 14 frames 320x512 (latents 14x8x40x64), bf16 compute, fp32 master weights, AdamW"),
    4: dict(frames=25, h=72, w=128, grad_ckpt=True, lora_rank=None,
            name="train_svd.py as scripted, bs=1/GPU, 25 frames 576x1024 (latents 25x8x72x128), bf16 compute, fp32 master weights, "
                 "gradient checkpointing (train_svd.py:731-732), AdamW"),
    5: dict(frames=14, h=40, w=64, grad_ckpt=False, lora_rank=64,
            name="train_svd_lora.py rank-64 LoRA on to_q/to_k/to_v/to_out.0 (train_svd_lora.py:659-671), bs=1/GPU, 14 frames 320x512, "
                 "bf16 base weights, fp32 LoRA parameters, AdamW"),
}


def edm_loss(model_pred, noisy_latents, latents, sigmas):
    """EDM preconditioning + sigma-weighted MSE in fp32, train_svd.py:1025-1036."""
    s2 = sigmas.float() ** 2
    c_out = -sigmas.float() / (s2 + 1).sqrt()
    c_skip = 1.0 / (s2 + 1)
    denoised = model_pred.float() * c_out + c_skip * noisy_latents.float()
    weighing = (1 + s2) / s2
    per_clip = (weighing * (denoised - latents.float()) ** 2).reshape(latents.shape[0], -1).mean(dim=1)
    return per_clip.mean()


def synthetic_batch(batch_size=1, num_frames=14, h=40, w=64, seed=1234, device="cpu", dtype=torch.float32, cross_dim=1024):
    """One synthetic clip batch shaped like the tensors train_svd.py feeds the UNet (seed = 1234 + rank):
    latents/noise ~ N(0,1), sigma ~ LogNormal(0.7, 1.6) (:964-967), timestep = 0.25 ln sigma (:969-970), model input =
    noisy / sqrt(sigma^2 + 1) concatenated with the per-clip conditioning latents repeated over the frames (:972, :1014-1017),
    added_time_ids = [fps-1 = 7, motion bucket 127, conditioning noise sigma ~ LogNormal(-3, 0.5)] (:954, :981-987),
    encoder_hidden_states = a unit-scale stand-in of the CLIP image embedding [B, 1, cross_dim] (:1000-1001)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    latents = torch.randn(batch_size, num_frames, 4, h, w, generator=g)
    noise = torch.randn(batch_size, num_frames, 4, h, w, generator=g)
    cond_latents = torch.randn(batch_size, 4, h, w, generator=g)
    sigmas = torch.exp(0.7 + 1.6 * torch.randn(batch_size, generator=g))[:, None, None, None, None]
    cond_sigma = torch.exp(-3.0 + 0.5 * torch.randn(batch_size, generator=g))
    enc = torch.randn(batch_size, 1, cross_dim, generator=g)
    noisy = latents + noise * sigmas
    timesteps = (0.25 * sigmas.log()).reshape(batch_size)
    inp = noisy / ((sigmas ** 2 + 1) ** 0.5)
    sample = torch.cat([inp, cond_latents.unsqueeze(1).repeat(1, num_frames, 1, 1, 1)], dim=2)
    added_time_ids = torch.stack([torch.full((batch_size,), 7.0), torch.full((batch_size,), 127.0), cond_sigma], dim=1)
    to = dict(device=device, dtype=dtype)
    return dict(sample=sample.to(**to), timestep=timesteps.to(device=device, dtype=torch.float32),
                encoder_hidden_states=enc.to(**to), added_time_ids=added_time_ids.to(**to),
                latents=latents.to(**to), noisy=noisy.to(**to), sigmas=sigmas.to(device=device, dtype=torch.float32))
