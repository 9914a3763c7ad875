"""``Tweediemix`` sampling-loop restatement (oracle; test infrastructure only).

Follows ``fusion_generation/fusion_sampling.py``: ``denoise_step`` ``:309-474``, ``init_fusion``
``:476-483``, ``run_fusion`` ``:485-489``, ``sample_loop`` ``:491-494``; LoRA differences from
``fusion_sampling_lora.py:324,378,476-490`` (``t_stop`` window; hook window one step shorter
than the sampler window — quirk ⑦).  Text encoders, VAE and the segmentation subprocess are
outside the hot path: text embeddings and region masks are inputs here (north_star: masks are
precomputed), so the jump loop ``:431-455`` — which never alters the returned latent — is run
only when ``run_jump=True`` (to reproduce the forward-call count).

Pinned against the reference's own ``init_fusion`` / ``alpha`` / ``denoise_step`` run unmodified on
CPU (``tests/golden/make_golden_sampler.py`` -> ``tests/golden/sampler_*.pt``, ``step_math_ref.pt``;
checked by ``tests/test_oracle_vs_reference_sampler.py`` at 2e-5 relative per step).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import step_math as sm
from .hooks_ref import register_time_ref
from .schedule import make_schedule


@dataclass
class RefConfig:
    guidance_scale: float = 0.8          # sample_catdog.sh:34
    n_timesteps: int = 50
    t_cond: float = 0.2
    t_stop: float | None = None          # LoRA variant only (sample_catdog.sh / _lora.py:547)
    resampling_steps: int = 10
    jumping_steps: int = 5
    resolution_h: int = 1024
    resolution_w: int = 1024
    seed: int = 3821


class TweediemixRef:
    def __init__(self, unet, text_embeds, text_embeds_single, masks, config: RefConfig,
                 concept_num: int, lora: bool = False, run_jump: bool = False):
        self.unet = unet
        self.text_embeds = text_embeds                  # ([K+2,77,D], [K+2,P]) = [uncond, multi, c_1..c_K]
        self.text_embeds_single = text_embeds_single    # ([K,77,D], [K,P])     = [uncond, single_1..single_{K-1}]
        self.masks = masks                              # [K,1,h,w]
        self.config = config
        self.concept_num = concept_num
        self.lora = lora
        self.run_jump = run_jump
        self.sched = make_schedule(config.n_timesteps)
        self.skip = self.sched.skip
        # fusion_sampling.py:70-78,222
        self.add_time_ids = torch.tensor([[config.resolution_h, config.resolution_w, 0, 0,
                                           config.resolution_h, config.resolution_w]])
        self.n_forward_rows = 0
        ts = self.sched.timesteps
        ic = int(config.n_timesteps * config.t_cond)                      # :486
        self.t_cond_prev, self.t_cond_cur, self.start_t = int(ts[ic - 1]), int(ts[ic]), int(ts[0])   # :478-480
        if lora:
            istop = int(config.n_timesteps * config.t_stop)               # _lora.py:489
            self.hook_window = set(int(v) for v in ts[ic:istop])          # _lora.py:477
            self.t_stop_cur = int(ts[istop])                              # _lora.py:478
        else:
            self.hook_window = set(int(v) for v in ts[ic:])               # :477
            self.t_stop_cur = None

    # -- helpers ----------------------------------------------------------------------------
    def alpha(self, t):
        return self.sched.alpha(t)

    def _unet(self, latents, t, ehs, pooled):
        self.n_forward_rows += latents.shape[0]
        cond = {"time_ids": self.add_time_ids.repeat(ehs.shape[0], 1).to(latents.device), "text_embeds": pooled}
        return self.unet(latents, t, encoder_hidden_states=ehs, added_cond_kwargs=cond)["sample"]

    def in_fused_phase(self, t) -> bool:
        ok = t <= self.t_cond_cur                                         # :324
        if self.lora:
            ok = ok and t >= self.t_stop_cur                              # _lora.py:324
        return ok

    def hook_gate_window(self):
        return self.hook_window

    # -- the step ---------------------------------------------------------------------------
    @torch.no_grad()
    def denoise_step(self, x, t):
        cfg, K, g = self.config, self.concept_num, self.config.guidance_scale
        t = int(t)
        E, P = self.text_embeds
        next_t = t - self.skip
        at, at_next = self.alpha(t), self.alpha(next_t)
        register_time_ref(self.unet, t, lora=self.lora)                   # :322

        if self.in_fused_phase(t):
            ehs = torch.cat([E[0:1], E[2:]])                              # :325-336
            pool = torch.cat([P[0:1], P[2:]])
            eps = self._unet(torch.cat([x] * (K + 1)), t, ehs, pool)      # :331,340
            x_next, _ = sm.fused_step(x, eps, self.masks.to(x.device), at, at_next, g, is_last=(t == 1))
            return x_next

        if t == self.start_t:                                             # :347-359
            Es, Ps = self.text_embeds_single
            ehs = torch.cat([E[0:1], E[1:2], Es[1:]])
            pool = torch.cat([P[0:1], P[1:2], Ps[1:]])
            eps = self._unet(torch.cat([x] * (K + 1)), t, ehs, pool)
            if cfg.resampling_steps <= 0:
                # fusion_sampling.py:417 deletes names only bound inside the loop (quirk ⑩)
                raise UnboundLocalError("reference cannot run with resampling_steps == 0")
            for _ in range(cfg.resampling_steps):                         # :390-415
                x0 = sm.resample_x0(x, eps, at, g, K)
                x_low = sm.ddim_update(x0, eps[:1], at_next)
                eps_next = self._unet(torch.cat([x_low, x_low]), next_t, ehs[:2], pool[:2])
                x = sm.renoise(x_low, eps_next, at, at_next, g)
                eps = self._unet(torch.cat([x] * (K + 1)), t, ehs, pool)
            eps2 = eps[:2]                                                # :421-423
        else:
            ehs, pool = E[:2], P[:2]                                      # :362-366
            eps2 = self._unet(torch.cat([x, x]), t, ehs, pool)
        x_next, x0 = sm.cfg_step(x, eps2, at, at_next, g, is_last=(t == 1))

        if t == self.t_cond_prev and self.run_jump and cfg.jumping_steps > 0:   # :431-448 (output-neutral)
            lat, t_tmp = x_next, next_t
            for _ in range(cfg.jumping_steps):
                a_tmp = self.alpha(t_tmp)
                e = self._unet(torch.cat([lat, lat]), t_tmp, E[:2], P[:2])
                t_tmp -= 150                                              # :444 (quirk ⑧)
                lat, self.jumped_x0 = sm.cfg_step(lat, e, a_tmp, self.alpha(t_tmp), g)
        return x_next

    @torch.no_grad()
    def sample_loop(self, x, callback=None):
        for i, t in enumerate(self.sched.timesteps):                      # :493-494
            x = self.denoise_step(x, int(t))
            if callback is not None:
                callback(i, int(t), x)
        return x

    def initial_latent(self):
        """:488,587 — seed everything, draw on the CPU generator, scale by init_noise_sigma (=1)."""
        torch.manual_seed(self.config.seed)
        h, w = self.config.resolution_h // 8, self.config.resolution_w // 8
        return torch.randn(1, 4, h, w) * self.sched.init_noise_sigma
