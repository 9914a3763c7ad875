"""Region-mask ingest restatement (oracle; test infrastructure only).

Follows ``fusion_generation/fusion_sampling.py:81-89`` (``preprocess_mask``:
JPEG -> L -> /255 -> threshold 0.5 -> nearest resize to the latent grid) and
``:466-469`` (background = clamp(1 - sum(foreground), 0), appended LAST).
The on-disk format is what ``text_segment/run_expand.py:84-87`` writes.
"""
from __future__ import annotations

import numpy as np
import torch
from PIL import Image


def preprocess_mask_ref(mask_path: str, h: int, w: int) -> torch.Tensor:
    grey = np.array(Image.open(mask_path).convert("L")).astype(np.float32) / 255.0
    binary = np.where(grey < 0.5, 0.0, 1.0).astype(np.float32)[None, None]
    return torch.nn.functional.interpolate(torch.from_numpy(binary), size=(h, w), mode="nearest")


def build_masks_ref(fg_paths, h: int, w: int) -> torch.Tensor:
    """[K,1,h,w] fp32 = foreground masks in order + background last."""
    fg = torch.cat([preprocess_mask_ref(p, h, w) for p in fg_paths])
    bg = 1 - torch.sum(fg, dim=0, keepdim=True)
    bg[bg < 0] = 0
    return torch.cat([fg, bg])
