"""Per-step arithmetic of TweedieMix ``denoise_step`` (oracle; test infrastructure only).

Pure functions of ``(x, eps, masks, at, at_next, g)``.  ``at`` / ``at_next`` are
0-dim fp32 tensors exactly as the reference holds them: a 0-dim
tensor does not take part in dtype promotion, so with fp16 ``eps`` the products
``(1-at).sqrt() * eps`` stay fp16 and only the subtraction from the fp32 latent
promotes to fp32 — evaluating these functions with fp16/bf16 ``eps`` therefore
reproduces the reference's mixed-precision rounding, and with fp32 ``eps`` gives
the fp32 ground truth.

Reference: ``fusion_generation/fusion_sampling.py`` (line numbers per function).
Pinned: the reference's unmodified ``denoise_step`` is run on CPU by
``tests/golden/make_golden_sampler.py`` (third-party imports stubbed) and
``tests/test_oracle_vs_reference_sampler.py`` compares this module to its latents in
all four phases; ``tests/test_oracle_step_math.py`` adds hand-derived known answers.
"""
from __future__ import annotations

import torch


def _s(v) -> torch.Tensor:
    """0-dim fp32 CPU tensor, like ``scheduler.alphas_cumprod[t]`` (fusion_sampling.py:305-307)."""
    return v if isinstance(v, torch.Tensor) else torch.tensor(float(v), dtype=torch.float32)


def cfg_combine(eps_uncond: torch.Tensor, eps_cond: torch.Tensor, g: float) -> torch.Tensor:
    """fusion_sampling.py:383,394,399,409,423,426 — evaluated in eps' own dtype."""
    return eps_uncond + g * (eps_cond - eps_uncond)


def tweedie_x0(x: torch.Tensor, eps: torch.Tensor, at) -> torch.Tensor:
    """fusion_sampling.py:385,395,400,411,428,446 — Tweedie's formula for the posterior mean."""
    at = _s(at)
    return (x - (1 - at).sqrt() * eps) / at.sqrt()


def ddim_update(x0: torch.Tensor, eps_uncond: torch.Tensor, at_next) -> torch.Tensor:
    """fusion_sampling.py:403,412,430,447 — noise term uses the UNCONDITIONAL eps (quirk ③)."""
    at_next = _s(at_next)
    return at_next.sqrt() * x0 + (1 - at_next).sqrt() * eps_uncond


def fused_x0(x, eps, masks, at, g: float) -> torch.Tensor:
    """fusion_sampling.py:376-385 — per-concept CFG, Tweedie x0, mask-weighted sum.

    x [1,4,h,w] fp32; eps [K+1,4,h,w] (row 0 uncond, row 1+c concept c); masks [K,1,h,w].
    """
    eps_u = eps[:1]
    x0 = 0
    for c in range(masks.shape[0]):
        eps_c = cfg_combine(eps_u, eps[1 + c:2 + c], g)
        x0 = x0 + masks[c].unsqueeze(0) * tweedie_x0(x, eps_c, at)
    return x0


def fused_step(x, eps, masks, at, at_next, g: float, is_last: bool = False):
    """Steady-state fused step: fusion_sampling.py:376-385, 430, 471-472. Returns (x_next, x0)."""
    x0 = fused_x0(x, eps, masks, at, g)
    x_next = ddim_update(x0, eps[:1], at_next)
    if is_last:                      # t == 1  (fusion_sampling.py:471-472)
        x_next = x0
    return x_next, x0


def cfg_step(x, eps, at, at_next, g: float, is_last: bool = False):
    """Plain two-row CFG step: fusion_sampling.py:425-430. eps [2,4,h,w]. Returns (x_next, x0)."""
    eps_u = eps[:1]
    x0 = tweedie_x0(x, cfg_combine(eps_u, eps[1:2], g), at)
    x_next = ddim_update(x0, eps_u, at_next)
    if is_last:
        x_next = x0
    return x_next, x0


def resample_x0(x, eps, at, g: float, concept_num: int) -> torch.Tensor:
    """Concept-aware x0 of the start-step resampling: fusion_sampling.py:392-401.

    eps rows: [uncond, multi, single_1 .. single_{K-1}];
    x0 = (K-1) * x0(multi) - sum_c x0(single_c).
    """
    eps_u = eps[:1]
    x0_mult = tweedie_x0(x, cfg_combine(eps_u, eps[1:2], g), at)
    x0 = (concept_num - 1) * x0_mult
    for c in range(concept_num - 1):
        x0 = x0 - tweedie_x0(x, cfg_combine(eps_u, eps[2 + c:3 + c], g), at)
    return x0


def renoise(x_low, eps2_next, at, at_next, g: float):
    """Second half of a resampling iteration: fusion_sampling.py:407-412.

    ``x_low`` is the latent already stepped to ``t - skip``; ``eps2_next`` the
    two-row prediction there.  Returns the latent re-noised back to ``t``.
    """
    eps_u = eps2_next[:1]
    x0_next = tweedie_x0(x_low, cfg_combine(eps_u, eps2_next[1:2], g), at_next)
    at = _s(at)
    return at.sqrt() * x0_next + (1 - at).sqrt() * eps_u


# --------------------------------------------------------------------------------------
# Linear (sharded) form used by the multi-GPU path, SURVEY §8e:
#   x0 = [ M x - s (1-g) M eps_u - s g sum_c m_c eps_c ] / sqrt(at),  M = sum_c m_c, s = sqrt(1-at)
# It is algebraically identical to fused_x0 in exact arithmetic and is what
# tmx_blend_partial_fwd / tmx_blend_finish_fwd evaluate in fp32.
# --------------------------------------------------------------------------------------

def blend_partial(eps_rows, mask_rows) -> torch.Tensor:
    """A_r = sum_{c in rank} m_c * eps_c, fp32.  eps_rows [R,4,h,w], mask_rows [R,1,h,w]."""
    return (mask_rows.float() * eps_rows.float()).sum(dim=0, keepdim=True)


def blend_finish(x, acc_masked_eps, eps_u, mask_sum, at, at_next, g: float, is_last: bool = False):
    """Finish from the all-reduced partials.  All fp32.  Returns (x_next, x0)."""
    at = _s(at)
    at_next = _s(at_next)
    s = (1 - at).sqrt()
    x0 = (mask_sum * x - s * (1 - g) * mask_sum * eps_u.float() - s * g * acc_masked_eps) / at.sqrt()
    x_next = at_next.sqrt() * x0 + (1 - at_next).sqrt() * eps_u.float()
    if is_last:
        x_next = x0
    return x_next, x0
