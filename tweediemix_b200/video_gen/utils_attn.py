"""Hook API of the video stage — drop-in for the live part of ``video_gen/utils_attn.py``:

  ``seed_everything(seed)``                                            utils_attn.py:8-12
  ``register_time(model, t)``                                          utils_attn.py:14-23
  ``register_conv_control_efficient(model, injection_schedule, interp)``   utils_attn.py:389-474

The reference replaces the ``forward`` of three ``ResnetBlock2D`` modules of the I2VGen-XL U-Net with a copy of diffusers' body that
ends in the frame-0 feature injection: on ``mid_block.resnets[0]`` and ``[1]`` frames 1..15 of both CFG halves are overwritten with
frame 0 while ``t`` is in ``injection_schedule`` (or ``t == 1000``); on ``up_blocks[1].resnets[0]`` they are blended,
``interp * frame0 + (1 - interp) * frame_t`` (``injection_schedule2``).  The batch is ``(b t) c h w`` with ``b = 2, t = 16`` hard-coded.

Here the module keeps whatever ``forward`` it has (diffusers' ``ResnetBlock2D``, the oracle stand-in, or this package's own block on the
tmx kernels) and the hook appends the injection as ONE in-place kernel (``tmx_frame_inject_fwd``) on its output, so nothing of the
ResNet body is restated.  Same attributes are left on the modules (``t``, ``injection_schedule``, ``injection_schedule2``, ``interp``).
``groups`` / ``frames`` default to the reference's literals.
"""
from __future__ import annotations

import torch

from .. import ops
from ..utils_custom import seed_everything  # noqa: F401  (same helper)


def _targets(model):
    u = model.unet
    return u.mid_block.resnets[0], u.mid_block.resnets[1], u.up_blocks[1].resnets[0], u.up_blocks[1].resnets[1]


def register_time(model, t):
    """utils_attn.py:14-23 — stamps ``t`` on up_blocks[1].resnets[0,1] and mid_block.resnets[0,1]."""
    t = int(t)
    for m in _targets(model):
        setattr(m, "t", t)


def _in_schedule(t, schedule) -> bool:
    if schedule is None:
        return False
    if torch.is_tensor(schedule):
        schedule = schedule.tolist()
    return int(t) in set(int(v) for v in schedule) or int(t) == 1000          # utils_attn.py:433,445


def register_conv_control_efficient(model, injection_schedule, interp, groups: int = 2, frames: int = 16, ref_rounding: bool = False):
    def patch(module, schedule, schedule2, interp_value):
        inner = module.__dict__.get("_tmx_inner_forward") or module.forward        # re-registration does not stack hooks
        module.__dict__["_tmx_inner_forward"] = inner

        def forward(*args, **kwargs):
            out = inner(*args, **kwargs)
            t = getattr(module, "t", None)
            if t is None:
                return out
            if _in_schedule(t, module.injection_schedule):                          # replace: frames 1.. := frame 0
                out = ops.frame_inject(_dense(out), groups, frames, 1.0, ref_rounding=ref_rounding)
            if _in_schedule(t, module.injection_schedule2):                         # blend with self.interp
                out = ops.frame_inject(_dense(out), groups, frames, float(module.interp), ref_rounding=ref_rounding)
            return out

        module.forward = forward
        setattr(module, "injection_schedule", schedule)
        setattr(module, "injection_schedule2", schedule2)
        if interp_value is not None:
            setattr(module, "interp", interp_value)

    mid0, mid1, up10, _ = _targets(model)
    patch(mid0, injection_schedule, None, None)                                     # utils_attn.py:459-466
    patch(mid1, injection_schedule, None, None)
    patch(up10, None, injection_schedule, interp)                                   # utils_attn.py:468-473


def _dense(x: torch.Tensor) -> torch.Tensor:
    if x.is_contiguous() or (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)):
        return x
    return x.contiguous()
