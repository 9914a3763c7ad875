"""The denoising-step arithmetic TweedieMix puts into ``I2VGenXLPipeline.__call__`` (``video_gen/pipeline_i2vgen_xl.py:647-656,679-719``):
schedule helpers (``skip = 1000 // n``, un-shifted ``alpha(t)`` with ``final_alpha_cumprod`` below zero, the injection window = the first
``int(n * injection_timestep)`` timesteps), the fused CFG + v-prediction Tweedie + DDIM update (``tmx_vpred_cfg_ddim_fwd``) and the
denoising loop itself (``:655-656,677-719``) around any U-Net callable."""
from __future__ import annotations

from typing import Sequence

import torch

from .. import ops


class VideoStepper:
    def __init__(self, alphas_cumprod: torch.Tensor, final_alpha_cumprod: float, timesteps: Sequence[int], guidance_scale: float,
                 injection_timestep: float = 0.02, num_train_timesteps: int = 1000, ref_rounding: bool = False):
        self.timesteps = [int(t) for t in timesteps]
        n = len(self.timesteps)
        self.skip = num_train_timesteps // n                                         # :647
        self._alpha = [float(v) for v in alphas_cumprod.tolist()]
        self._final = float(final_alpha_cumprod)                                     # :648
        k = int(n * injection_timestep)                                              # :653-654
        self.injection_schedule = self.timesteps[:k] if k >= 0 else []
        self.guidance_scale = float(guidance_scale)
        self.ref_rounding = ref_rounding

    def alpha(self, t: int) -> float:
        """pipeline_i2vgen_xl.py:480-482 — NOT shifted (unlike the image sampler's table)."""
        return self._alpha[t] if t >= 0 else self._final

    @torch.no_grad()
    def step(self, latents: torch.Tensor, noise_pred: torch.Tensor, t: int, out=None) -> torch.Tensor:
        """latents [B, C, F, H, W]; noise_pred [2B, C, F, H, W] = (unconditional, text) halves of the U-Net's v-prediction (:682-713)."""
        B = latents.shape[0]
        v_u, v_c = noise_pred[:B], noise_pred[B:]
        return ops.vpred_cfg_ddim(latents.contiguous(), v_u.contiguous(), v_c.contiguous(), self.alpha(int(t)), self.alpha(int(t) - self.skip),
                                  self.guidance_scale, out=out, ref_rounding=self.ref_rounding)

    @torch.no_grad()
    def denoise_loop(self, model, latents: torch.Tensor, unet_forward, interp_ratio: float = 0.7, callback=None) -> torch.Tensor:
        """The reference's loop (``pipeline_i2vgen_xl.py:655-656,677-719``) around ``unet_forward(latent_model_input, t) -> v-prediction``:
        installs the frame-0 injection hooks for this run's window, then per timestep stamps ``t`` on the hooked blocks
        (``register_time``), feeds both CFG halves of the latents (``torch.cat([latents] * 2)``; DDIM's ``scale_model_input`` is the
        identity) and takes the fused guidance + Tweedie + DDIM step.  ``model`` is whatever carries ``.unet`` (the pipeline object in the
        reference); the U-Net itself — the I2VGen-XL model of diffusers, or a stand-in — is the caller's."""
        from .utils_attn import register_conv_control_efficient, register_time
        register_conv_control_efficient(model, self.injection_schedule, interp_ratio)                 # :655-656
        for i, t in enumerate(self.timesteps):
            register_time(model, t)                                                                    # :683
            noise_pred = unet_forward(torch.cat([latents] * 2), t)                                     # :681-696
            latents = self.step(latents, noise_pred, t)                                                # :698-713
            if callback is not None:
                callback(i, t, latents)
        return latents
