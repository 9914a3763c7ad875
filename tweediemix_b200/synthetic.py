"""Seeded stand-ins for everything the hot path loads from disk or from other models.

Neither the build image nor the GPU box has a network, SDXL weights, concept ``delta-*.bin`` files or
text encoders, so benchmarks and the end-to-end tests run the real architecture (SDXL-base U-Net
shapes, 70 transformer blocks, per-concept K/V matrices or rank-4 LoRA layers) on seeded random
weights, seeded text embeddings and either the reference's shipped example masks or a stripe
partition.  The arithmetic and memory traffic are those of the real run; only the values differ.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from .fusion_sampling import FusionComponents
from .masks import stripe_masks
from .model_lora import LoRAAttnProcessor_base
from .schedule import DDIMSchedule
from .unet import TmxUNet2DConditionModel, UNetConfig, init_synthetic_


class SparseUNet(nn.Module):
    """A module tree that only has the leaves a concept checkpoint carries, addressable by the same
    dotted paths as the full U-Net (``get_submodule('up_blocks.0.attentions.1.transformer_blocks.3.attn2')``)."""

    def add_leaf(self, path: str, module: nn.Module):
        node = self
        parts = path.split(".")
        for part in parts[:-1]:
            nxt = node._modules.get(part)
            if nxt is None:
                nxt = nn.Module()
                node.add_module(part, nxt)
            node = nxt
        node.add_module(parts[-1], module)
        return module


class _KV(nn.Module):
    def __init__(self, to_k: nn.Linear, to_v: nn.Linear):
        super().__init__()
        self.to_k, self.to_v = to_k, to_v


def make_custom_concept(unet: TmxUNet2DConditionModel, seed: int, rel: float = 0.5) -> SparseUNet:
    """What a Custom-Diffusion ``delta.bin`` holds: every ``attn2.to_k/to_v`` (``fusion_sampling.py:206-209``),
    here = base + N(0, (rel * std(W))^2)."""
    donor = SparseUNet()
    gen = {}
    for name, blk in unet.transformer_blocks():
        pair = []
        for lin in (blk.attn2.to_k, blk.attn2.to_v):
            w = lin.weight.detach()
            g = gen.setdefault(w.device, torch.Generator(device=w.device).manual_seed(seed))
            new = nn.Linear(lin.in_features, lin.out_features, bias=False, device=w.device, dtype=w.dtype)
            noise = torch.randn(w.shape, generator=g, device=w.device, dtype=torch.float32)
            new.weight.data = (w.float() + noise * (rel * w.float().std())).to(w.dtype)
            pair.append(new)
        donor.add_leaf(name + ".attn2", _KV(*pair))
    return donor.requires_grad_(False)


class _Proc(nn.Module):
    def __init__(self, processor):
        super().__init__()
        self.processor = processor


def make_lora_concept(unet: TmxUNet2DConditionModel, seed: int, up_std: float = 1e-2, rank: int = 4) -> SparseUNet:
    """A LoRA concept: ``<attention>.processor.to_{q,k,v,out}_lora`` on all 140 attentions; ``up`` is drawn
    non-zero so that the routing is observable (a freshly initialised LoRA is the identity)."""
    donor = SparseUNet()
    g = torch.Generator().manual_seed(seed)
    for name, attn in unet.attention_modules():
        hidden = attn.to_q.out_features
        proc = LoRAAttnProcessor_base(hidden, attn.to_k.in_features if attn.is_cross else None, rank)
        for layer in (proc.to_q_lora, proc.to_k_lora, proc.to_v_lora, proc.to_out_lora):
            layer.down.weight.data = torch.randn(layer.down.weight.shape, generator=g) / rank
            layer.up.weight.data = torch.randn(layer.up.weight.shape, generator=g) * up_std
        donor.add_leaf(name, _Proc(proc.to(attn.to_q.weight.device, attn.to_q.weight.dtype)))
    return donor.requires_grad_(False)


def make_text(cfg: UNetConfig, concept_num: int, seed: int, tokens: int = 77, device="cpu", dtype=torch.float32):
    """([uncond, multi, c_1..c_K], pooled) and ([uncond, single_1..single_{K-1}], pooled), the row
    orders of ``fusion_sampling.py:194-196``."""
    g = torch.Generator().manual_seed(seed)
    E = torch.randn(concept_num + 2, tokens, cfg.cross_attention_dim, generator=g)
    P = torch.randn(concept_num + 2, cfg.pooled_dim, generator=g)
    Es = torch.cat([E[0:1], torch.randn(concept_num - 1, tokens, cfg.cross_attention_dim, generator=g)])
    Ps = torch.cat([P[0:1], torch.randn(concept_num - 1, cfg.pooled_dim, generator=g)])
    mv = lambda t: t.to(device=device, dtype=dtype)
    return (mv(E), mv(P)), (mv(Es), mv(Ps))


def make_components(concept_num: int = 3, variant: str = "custom", seed: int = 0, device="cuda",
                    dtype=torch.bfloat16, latent_hw: Tuple[int, int] = (128, 128),
                    unet_cfg: Optional[UNetConfig] = None, masks: Optional[torch.Tensor] = None,
                    unet: Optional[TmxUNet2DConditionModel] = None) -> FusionComponents:
    cfg = unet_cfg or UNetConfig.sdxl_base()
    if unet is None:
        with torch.device(device):
            unet = TmxUNet2DConditionModel(cfg).to(dtype)
        init_synthetic_(unet, seed + 2)
        unet.requires_grad_(False).eval().finalize()
    make = make_lora_concept if variant == "lora" else make_custom_concept
    donors = [make(unet, seed + 100 + i) for i in range(concept_num)]
    text, text_single = make_text(cfg, concept_num, seed + 1, device=device, dtype=dtype)
    if masks is None:
        masks = stripe_masks(concept_num, latent_hw[0], latent_hw[1], device=device)
    return FusionComponents(unet=unet, concept_unets=donors, text_embeds=text, text_embeds_single=text_single,
                            scheduler=DDIMSchedule(), masks=masks)
