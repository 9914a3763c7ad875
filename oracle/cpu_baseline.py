"""CPU baseline leg of bench.py (oracle; test/measurement infrastructure only — never shipped).

Times the reference's PyTorch path for the hot loop on the host cores: the diffusers-shaped SDXL
U-Net stand-in (``oracle/unet_ref.py``) in fp32 eager with the restated hook forward
(``oracle/hooks_ref.py``, pinned against the reference's own ``utils_custom.py`` outputs) — i.e.
``kind = "port"``: the reference sampler itself cannot be imported (diffusers / xformers absent,
SURVEY §8c) and ``/root/reference`` does not exist on the GPU box.

A full 1024² image is 242 sample-forwards x 6.76 TFLOP = 1.64 PFLOP (~1 h on 8-16 cores), so the
timed sample is ONE sample-forward (one batch row of one U-Net call) at the full 1024² size, and
images/s is extrapolated by the schedule's forward count.  Weight values do not matter for timing;
they are tiled from one seeded vector because drawing 2.6 G normals serially takes longer than the
measurement itself.
"""
from __future__ import annotations

import os
import time

import torch

from .hooks_ref import register_custom_ref, register_time_ref
from .unet_ref import UNet2DConditionModelRef, UNetConfig, transformer_blocks_in_hook_order

FORWARDS_PER_IMAGE = 242          # 64 (step 0 incl. 10 resampling iters) + 18 (steps 1-9) + 160 (steps 10-49); SURVEY §3.3
FORWARDS_PER_FUSED_STEP = 4       # K + 1 rows, K = 3


def _fast_fill_(model: torch.nn.Module, seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(1 << 20, generator=g)
    for name, p in model.named_parameters():
        n = p.numel()
        reps = -(-n // base.numel())
        std = 0.02 if p.ndim == 1 else p[0].numel() ** -0.5 * (0.3 if name.endswith(("to_out.0.weight", "ff.net.2.weight", "conv2.weight", "proj_out.weight")) else 1.0)
        vals = base.repeat(reps)[:n].reshape(p.shape) * std
        if p.ndim == 1 and not name.endswith("bias"):
            vals = vals + 1.0
        p.data = vals.contiguous()
    return model


class CpuReference:
    """Built once, then ``forward_seconds()`` per timed step."""

    def __init__(self, res: int = 1024, threads: int | None = None, cfg: UNetConfig | None = None):
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.cfg = cfg or UNetConfig.sdxl()
        with torch.device("meta"):
            unet = UNet2DConditionModelRef(self.cfg)
        unet = unet.to_empty(device="cpu")
        self.unet = _fast_fill_(unet).eval().requires_grad_(False)
        # the reference's hooked attn2 forward (naive einsum/softmax/einsum); batch 1 never meets the
        # B == 4 gate, so the base weights are used — same arithmetic volume as a routed row
        register_custom_ref(self.unet, [], torch.tensor([781]), 0)
        register_time_ref(self.unet, 781)
        h = res // 8
        g = torch.Generator().manual_seed(1)
        self.x = torch.randn(1, 4, h, h, generator=g)
        self.E = torch.randn(1, 77, self.cfg.cross_attention_dim, generator=g)
        self.cond = {"text_embeds": torch.randn(1, self.cfg.pooled_embed_dim, generator=g),
                     "time_ids": torch.tensor([[res, res, 0, 0, res, res]])}
        self.res = res

    @torch.no_grad()
    def forward_seconds(self) -> float:
        t0 = time.perf_counter()
        out = self.unet(self.x, 781, self.E, self.cond)["sample"]
        dt = time.perf_counter() - t0
        assert torch.isfinite(out).all()
        return dt

    @staticmethod
    def images_per_second(seconds_per_forward: float) -> float:
        return 1.0 / (FORWARDS_PER_IMAGE * seconds_per_forward)

    def sample_description(self) -> str:
        return (f"1 U-Net sample-forward (1 batch row of the K=3 fused step) at {self.res}x{self.res}, fp32 eager, "
                f"{self.threads} threads; images/s extrapolated x{FORWARDS_PER_IMAGE} forwards per 50-step image")
