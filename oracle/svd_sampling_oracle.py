"""ORACLE — test infrastructure only. Plain-PyTorch restatement of the UNet-facing part of the inference / validation path
the reference reaches through diffusers' `StableVideoDiffusionPipeline.__call__` (train_svd.py:1106-1140, infer_svd.ipynb
cell 3): EulerDiscreteScheduler in the SVD configuration (v-prediction, continuous timesteps 0.25 ln sigma, Karras sigmas
rho = 7 between sigma_min 0.002 and sigma_max 700, `timestep_spacing = "leading"` => init_noise_sigma = sqrt(sigma_max^2 + 1)),
classifier-free guidance with a per-frame guidance scale (linspace(min, max, num_frames)) on a doubled batch, conditioning
latents concatenated on the channel axis.

[D] = diffusers pipelines/stable_video_diffusion/pipeline_stable_video_diffusion.py and schedulers/scheduling_euler_discrete.py
(absent from /root/reference; restated from their published algorithm — PARITY UNPINNED against diffusers, see
oracle/svd_unet_oracle.py). VAE encode / decode and the CLIP image encoder are outside this restatement: the functions take
the image latents and the image embedding as inputs and return the denoised latents.
"""
from __future__ import annotations

import math

import torch


def karras_sigmas(num_inference_steps: int, sigma_min: float = 0.002, sigma_max: float = 700.0, rho: float = 7.0) -> torch.Tensor:
    """[D] EulerDiscreteScheduler._convert_to_karras, then the appended final sigma 0 (set_timesteps)."""
    ramp = torch.linspace(0, 1, num_inference_steps, dtype=torch.float64)
    min_inv_rho, max_inv_rho = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    sig = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return torch.cat([sig, torch.zeros(1, dtype=torch.float64)]).float()


def sample_latents(unet, image_latents, image_embeddings, *, num_frames, fps=7, motion_bucket_id=127, noise_aug_strength=0.02,
                   num_inference_steps=25, min_guidance_scale=1.0, max_guidance_scale=3.0, noise=None, generator=None):
    """image_latents [B,4,h,w] (VAE mode of the noise-augmented conditioning image, NOT scaled), image_embeddings [B,1,D].
    Returns the denoised video latents [B,T,4,h,w]. Follows [D] StableVideoDiffusionPipeline.__call__ steps 3-8."""
    B, _, h, w = image_latents.shape
    dev, dt = image_latents.device, image_latents.dtype
    cfg = max_guidance_scale > 1.0
    # [D] _encode_image / _encode_vae_image: the unconditional branch uses zeros
    emb = torch.cat([torch.zeros_like(image_embeddings), image_embeddings]) if cfg else image_embeddings
    lat = torch.cat([torch.zeros_like(image_latents), image_latents]) if cfg else image_latents
    lat = lat.unsqueeze(1).repeat(1, num_frames, 1, 1, 1)
    # [D] _get_add_time_ids: (fps - 1, motion_bucket_id, noise_aug_strength) — __call__ passes fps - 1
    ids = torch.tensor([[float(fps - 1), float(motion_bucket_id), float(noise_aug_strength)]], device=dev, dtype=dt).repeat(B, 1)
    ids = torch.cat([ids, ids]) if cfg else ids
    sigmas = karras_sigmas(num_inference_steps).to(dev)
    timesteps = 0.25 * sigmas[:-1].log()
    if noise is None:
        noise = torch.randn(B, num_frames, 4, h, w, generator=generator, device=dev, dtype=dt)
    latents = noise * math.sqrt(float(sigmas[0]) ** 2 + 1.0)       # init_noise_sigma, timestep_spacing "leading"
    gs = torch.linspace(min_guidance_scale, max_guidance_scale, num_frames, device=dev, dtype=dt)[None, :, None, None, None]
    for i in range(num_inference_steps):
        sigma, sigma_next = sigmas[i], sigmas[i + 1]
        x = torch.cat([latents] * 2) if cfg else latents
        x = x / ((sigma ** 2 + 1) ** 0.5)                          # scheduler.scale_model_input
        x = torch.cat([x, lat], dim=2)
        t = timesteps[i].expand(x.shape[0])
        v = unet(x, t, emb, added_time_ids=ids).sample
        if cfg:
            v_u, v_c = v.chunk(2)
            v = v_u + gs * (v_c - v_u)
        # [D] EulerDiscreteScheduler.step, v_prediction, s_churn = 0
        pred_x0 = v * (-sigma / (sigma ** 2 + 1) ** 0.5) + latents / (sigma ** 2 + 1)
        derivative = (latents - pred_x0) / sigma
        latents = latents + derivative * (sigma_next - sigma)
    return latents
