"""Region-mask ingest: the on-disk hand-off from the segmentation stage to the sampler.

``text_segment/run_expand.py:84-87`` writes one boolean image per foreground concept as
``<seg_concept>.jpg`` (1024² L-mode JPEG, so the 0/255 levels carry JPEG ringing);
``fusion_generation/fusion_sampling.py:81-89`` reads each back as /255 -> threshold 0.5 ->
nearest-neighbour resize to the latent grid, and ``:466-469`` appends the background
``clamp(1 - sum(fg), 0)`` as the LAST concept.  ``load_region_masks`` is that ingest.
"""
from __future__ import annotations

import os
from typing import Sequence

import numpy as np
import torch
from PIL import Image


def preprocess_mask(mask_path: str, h: int, w: int, device="cpu") -> torch.Tensor:
    """Same name / arguments / result as the reference helper: [1,1,h,w] fp32 in {0,1}."""
    levels = np.asarray(Image.open(mask_path).convert("L"), dtype=np.float32) / 255.0
    hard = torch.from_numpy((levels >= 0.5).astype(np.float32))[None, None].to(device)
    return torch.nn.functional.interpolate(hard, size=(h, w), mode="nearest")


def assemble_masks(fg_masks: torch.Tensor) -> torch.Tensor:
    """[K-1,1,h,w] foreground -> [K,1,h,w] with the clamped complement appended (:467-469)."""
    background = (1.0 - fg_masks.sum(dim=0, keepdim=True)).clamp_(min=0.0)
    return torch.cat([fg_masks, background])


def load_region_masks(directory: str, seg_concepts: Sequence[str] | str, h: int, w: int, device="cpu") -> torch.Tensor:
    if isinstance(seg_concepts, str):
        seg_concepts = seg_concepts.split("+")
    paths = [os.path.join(directory, name + ".jpg") for name in seg_concepts]
    missing = [p for p in paths if not os.path.exists(p)]
    if missing:
        raise FileNotFoundError(f"region masks not found: {missing}")
    return assemble_masks(torch.cat([preprocess_mask(p, h, w, device) for p in paths]))


def stripe_masks(k: int, h: int, w: int, device="cpu") -> torch.Tensor:
    """Synthetic vertical-stripe partition (benchmark stand-in when K has no fixture masks)."""
    col = torch.arange(w, device=device)
    owner = torch.clamp((col * k) // w, max=k - 1)
    m = (owner[None, :] == torch.arange(k, device=device)[:, None]).float()       # [k, w]
    return m[:, None, None, :].expand(k, 1, h, w).contiguous()
