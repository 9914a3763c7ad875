"""DDIM schedule restatement (oracle; test infrastructure only).

Follows:
  * [D] diffusers 0.29.2 ``schedulers/scheduling_ddim.py`` with the SDXL-base
    scheduler config (scaled_linear, beta 0.00085..0.012, 1000 train steps,
    steps_offset 1, timestep_spacing "leading", set_alpha_to_one False);
  * ``fusion_generation/fusion_sampling.py:212-218`` (N_ts taken *before*
    set_timesteps, ``skip = N_ts // n``, the ``cat([1.0], alphas_cumprod)``
    one-slot shift) and ``:305-307`` (``alpha(t)``).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

NUM_TRAIN_TIMESTEPS = 1000
BETA_START = 0.00085
BETA_END = 0.012
STEPS_OFFSET = 1


@dataclass
class RefSchedule:
    timesteps: torch.Tensor          # int64 [n], descending (981, 961, ... 1 for n=50)
    alphas_cumprod: torch.Tensor     # fp32 [1001], the reference's SHIFTED table
    final_alpha_cumprod: torch.Tensor  # 0-dim fp32
    skip: int
    init_noise_sigma: float = 1.0

    def alpha(self, t) -> torch.Tensor:
        """fusion_sampling.py:305-307 — index the shifted table, or final alpha for t<0."""
        t = int(t)
        return self.alphas_cumprod[t] if t >= 0 else self.final_alpha_cumprod


def make_schedule(n_timesteps: int) -> RefSchedule:
    # [D] betas = linspace(sqrt(b0), sqrt(b1), T, fp32) ** 2 ; alphas_cumprod = cumprod(1 - betas)
    betas = torch.linspace(BETA_START ** 0.5, BETA_END ** 0.5, NUM_TRAIN_TIMESTEPS, dtype=torch.float32) ** 2
    alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
    # [D] set_alpha_to_one=False -> final_alpha_cumprod = alphas_cumprod[0]
    final_alpha = alphas_cumprod[0].clone()
    # fusion_sampling.py:213 — len(scheduler.timesteps) before set_timesteps is the train length
    n_ts_before = NUM_TRAIN_TIMESTEPS
    # [D] "leading": arange(n) * (T // n), reversed, + steps_offset
    step_ratio = NUM_TRAIN_TIMESTEPS // n_timesteps
    ts = (np.arange(0, n_timesteps) * step_ratio).round()[::-1].copy().astype(np.int64) + STEPS_OFFSET
    skip = n_ts_before // n_timesteps                      # fusion_sampling.py:216
    shifted = torch.cat([torch.tensor([1.0]), alphas_cumprod])  # fusion_sampling.py:218
    return RefSchedule(torch.from_numpy(ts), shifted, final_alpha, skip)
