"""DDIM schedule as ``Tweediemix`` consumes it.

The reference asks diffusers' ``DDIMScheduler`` (SDXL scheduler config) for exactly four things —
``timesteps``, ``alphas_cumprod``, ``final_alpha_cumprod`` and ``init_noise_sigma``
(``fusion_generation/fusion_sampling.py:212-218,488``); ``scheduler.step`` is never called on the
image path.  This class provides those four with the same arithmetic ([D] diffusers 0.29.2
``scheduling_ddim.py``: ``scaled_linear`` betas in fp32, ``leading`` spacing, ``steps_offset=1``,
``set_alpha_to_one=False``), so the table indexes and values agree with a diffusers scheduler.
"""
from __future__ import annotations

import torch


class DDIMSchedule:
    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 steps_offset: int = 1):
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        root = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32)
        self.betas = root * root
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]            # set_alpha_to_one = False
        self.init_noise_sigma = 1.0
        # before set_timesteps a diffusers scheduler exposes all training steps, descending;
        # the reference measures N_ts = len(timesteps) at that point (fusion_sampling.py:213)
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1, dtype=torch.int64)

    def set_timesteps(self, num_inference_steps: int, device=None):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("num_inference_steps exceeds the training schedule")
        ratio = self.num_train_timesteps // num_inference_steps
        ts = torch.arange(num_inference_steps - 1, -1, -1, dtype=torch.int64) * ratio + self.steps_offset
        self.timesteps = ts if device is None else ts.to(device)
        self.num_inference_steps = num_inference_steps
