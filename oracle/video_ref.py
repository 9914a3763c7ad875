"""Video-stage restatement (oracle; test infrastructure only).

Follows ``video_gen/utils_attn.py``: ``register_time`` ``:14-23``, the patched ``ResnetBlock2D.forward`` with the frame-0 feature
injection ``:389-474`` (diffusers body ``:391-431``, replacement ``:433-443``, interpolation ``:445-456``, targets ``:459-473``), and
``video_gen/pipeline_i2vgen_xl.py``: ``alpha`` ``:480-482``, guidance + v-prediction Tweedie + DDIM update ``:682-713``.

PINNED: ``tests/golden/make_golden_video.py`` runs the reference's own unmodified ``utils_attn.py`` on the stand-in below and executes
the reference's own source lines of the step (extracted verbatim from ``pipeline_i2vgen_xl.py`` at generation time) -> ``tests/golden/
video_*.pt``; ``tests/test_oracle_video.py`` replays them.  The I2VGen-XL U-Net itself is a diffusers dependency and is not restated.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .unet_ref import ResnetBlock2D

GROUPS, FRAMES = 2, 16          # the reference's literals `b=2, t=16`


class ResnetBlock2DVideo(ResnetBlock2D):
    """The diffusers ``ResnetBlock2D`` attribute surface the reference's patched forward touches (``:391-431``)."""

    def __init__(self, cin, cout, temb_dim, groups=8, eps=1e-5):
        super().__init__(cin, cout, temb_dim, groups, eps)
        self.nonlinearity = F.silu
        self.upsample = self.downsample = None
        self.time_embedding_norm = "default"


class VideoUNetStub(nn.Module):
    """Only the module paths the hooks address: ``mid_block.resnets[0,1]`` and ``up_blocks[1].resnets[0,1]``."""

    def __init__(self, c=32, temb=64, seed=0):
        super().__init__()
        torch.manual_seed(seed)
        mk = lambda ci, co: ResnetBlock2DVideo(ci, co, temb)
        self.mid_block = nn.Module()
        self.mid_block.resnets = nn.ModuleList([mk(c, c), mk(c, c)])
        self.up_blocks = nn.ModuleList([nn.Module(), nn.Module()])
        self.up_blocks[0].resnets = nn.ModuleList([mk(c, c)])
        self.up_blocks[1].resnets = nn.ModuleList([mk(2 * c, c), mk(c, c)])
        for p in self.parameters():
            p.requires_grad_(False)


def register_time_ref(model, t):
    for m in (model.unet.up_blocks[1].resnets[0], model.unet.up_blocks[1].resnets[1], model.unet.mid_block.resnets[0], model.unet.mid_block.resnets[1]):
        m.t = t


def _hit(t, schedule) -> bool:
    if schedule is None:
        return False
    vals = schedule.tolist() if torch.is_tensor(schedule) else list(schedule)
    return t in vals or t == 1000


def inject_ref(out: torch.Tensor, interp: float | None) -> torch.Tensor:
    """``:433-456`` on a [(b t), c, h, w] tensor; ``interp is None`` = replacement."""
    bt, c, h, w = out.shape
    o = out.reshape(GROUPS, FRAMES, c, h, w).clone()
    first = o[:, :1]
    if interp is None:
        o[:, 1:] = first.repeat(1, FRAMES - 1, 1, 1, 1)
    else:
        o[:, 1:] = interp * first.repeat(1, FRAMES - 1, 1, 1, 1) + (1 - interp) * o[:, 1:]
    return o.reshape(bt, c, h, w)


def conv_forward_ref(block, input_tensor, temb):
    """``:390-458`` for the branch the I2VGen-XL blocks take (no up/down-sampling, ``time_embedding_norm == "default"``)."""
    h = block.conv1(F.silu(block.norm1(input_tensor)))
    if temb is not None:
        h = h + block.time_emb_proj(F.silu(temb))[:, :, None, None]
    h = block.conv2(block.dropout(F.silu(block.norm2(h))))
    x = block.conv_shortcut(input_tensor) if block.conv_shortcut is not None else input_tensor
    out = (x + h) / block.output_scale_factor
    if _hit(block.t, block.injection_schedule):
        out = inject_ref(out, None)
    if _hit(block.t, block.injection_schedule2):
        out = inject_ref(out, block.interp)
    return out


def register_conv_ref(model, injection_schedule, interp):
    def patch(m, s1, s2, ip):
        m.injection_schedule, m.injection_schedule2 = s1, s2
        if ip is not None:
            m.interp = ip
        m.forward = lambda input_tensor, temb, _m=m: conv_forward_ref(_m, input_tensor, temb)
    u = model.unet
    patch(u.mid_block.resnets[0], injection_schedule, None, None)
    patch(u.mid_block.resnets[1], injection_schedule, None, None)
    patch(u.up_blocks[1].resnets[0], None, injection_schedule, interp)


def vpred_step_ref(latents, noise_pred, at, at_next, guidance_scale):
    """``pipeline_i2vgen_xl.py:682-713``.  latents [B,C,F,H,W]; noise_pred [2B,C,F,H,W]; at / at_next 0-dim tensors (``alphas_cumprod[t]``).
    Evaluated with torch ops in the tensors' own dtype, so the roundings are PyTorch's."""
    u, c = noise_pred.chunk(2)
    v = u + guidance_scale * (c - u)
    b, ch, f, hh, ww = latents.shape
    x = latents.permute(0, 2, 1, 3, 4).reshape(b * f, ch, hh, ww)
    v = v.permute(0, 2, 1, 3, 4).reshape(b * f, ch, hh, ww)
    eps = at.sqrt() * v + (1 - at).sqrt() * x
    x0 = at.sqrt() * x - (1 - at).sqrt() * v
    x = at_next.sqrt() * x0 + (1 - at_next).sqrt() * eps
    back = lambda z: z[None, :].reshape(b, f, ch, hh, ww).permute(0, 2, 1, 3, 4)
    return back(x), back(x0)


def video_loop_ref(model, latents, unet_forward, alphas_cumprod, final_alpha, timesteps, guidance_scale, injection_timestep, interp):
    """``pipeline_i2vgen_xl.py:647-656,677-719``: the denoising loop with the reference-style hooks of this module installed
    (``register_conv_ref``), ``register_time_ref`` per step, both CFG halves through ``unet_forward`` and the step lines."""
    n = len(timesteps)
    skip = 1000 // n
    k = int(n * injection_timestep)
    schedule = list(timesteps[:k]) if k >= 0 else []
    register_conv_ref(model, schedule, interp)
    alpha = lambda t: alphas_cumprod[t] if t >= 0 else final_alpha
    for t in timesteps:
        register_time_ref(model, int(t))
        noise_pred = unet_forward(torch.cat([latents] * 2), int(t))
        latents = vpred_step_ref(latents, noise_pred, torch.as_tensor(alpha(int(t))), torch.as_tensor(alpha(int(t) - skip)), guidance_scale)[0]
    return latents
