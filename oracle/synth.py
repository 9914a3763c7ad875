"""Seeded synthetic inputs for the oracle (test infrastructure only).

No SDXL checkpoint, concept ``delta-*.bin`` or text encoder is reachable offline, so weights,
per-concept deltas and text embeddings are seeded draws.  Shared by the golden
generator, the tests and the CPU-baseline leg of ``bench.py`` so they all see the same tensors.
"""
from __future__ import annotations

import copy
import os

import torch

from .masks_ref import build_masks_ref
from .unet_ref import UNet2DConditionModelRef, UNetConfig, seeded_init_, transformer_blocks_in_hook_order

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def make_base_unet(cfg: UNetConfig, seed: int) -> UNet2DConditionModelRef:
    return seeded_init_(UNet2DConditionModelRef(cfg), seed).eval().requires_grad_(False)


def make_concept_unet(base, seed: int, rel: float = 0.5):
    """A 'fine-tuned' copy: only ``attn2.to_k/to_v`` differ from the base, which is exactly what a
    Custom-Diffusion ``delta.bin`` carries (``fusion_sampling.py:206-209``)."""
    u = copy.deepcopy(base)
    g = torch.Generator().manual_seed(seed)
    for _, blk in transformer_blocks_in_hook_order(u):
        for lin in (blk.attn2.to_k, blk.attn2.to_v):
            w = lin.weight
            with torch.no_grad():
                w.add_(torch.randn(w.shape, generator=g) * (rel * w.std()))
    return u


def make_text(cfg: UNetConfig, concept_num: int, seed: int, tokens: int = 77):
    """(text_embeds, text_embeds_single) in the reference's row order
    (``fusion_sampling.py:194-196``): [uncond, multi, c_1..c_K] and [uncond, single_1..single_{K-1}]."""
    g = torch.Generator().manual_seed(seed)
    E = torch.randn(concept_num + 2, tokens, cfg.cross_attention_dim, generator=g)
    P = torch.randn(concept_num + 2, cfg.pooled_embed_dim, generator=g)
    Es = torch.cat([E[0:1], torch.randn(concept_num - 1, tokens, cfg.cross_attention_dim, generator=g)])
    Ps = torch.cat([P[0:1], torch.randn(concept_num - 1, cfg.pooled_embed_dim, generator=g)])
    return (E, P), (Es, Ps)


def fixture_masks(h: int, w: int, which: str = "test_out") -> torch.Tensor:
    """Real region masks shipped by the reference under ``example_results/<which>/`` (copied as
    data fixtures to ``tests/golden/masks``), through the restated ingest: [3,1,h,w]."""
    d = os.path.join(GOLDEN_DIR, "masks", which)
    names = sorted(f for f in os.listdir(d) if f.endswith(".jpg"))
    return build_masks_ref([os.path.join(d, n) for n in names], h, w)


def stripe_masks(k: int, h: int, w: int) -> torch.Tensor:
    """Synthetic vertical-stripe partition for K != 3."""
    m = torch.zeros(k, 1, h, w)
    edges = [round(i * w / k) for i in range(k + 1)]
    for c in range(k):
        m[c, :, :, edges[c]:edges[c + 1]] = 1
    return m
