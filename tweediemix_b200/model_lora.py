"""LoRA layer definitions used at sampling time — the live part of ``fusion_generation/model_lora.py``
(``LoRALinearLayer`` ``:28-48``, ``LoRAAttnProcessor_base`` ``:104-116``, ``create_lora_diffusion_base``
``:169-189``).  The rest of that reference file is training-pipeline code that the sampler never
executes and is out of scope.

State-dict keys follow the reference's checkpoints:
``<attention path>.processor.to_{q,k,v,out}_lora.{down,up}.weight``.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class LoRALinearLayer(nn.Module):
    """``up(down(x))`` with rank 4 and no alpha/scale; ``down ~ N(0, 1/rank)``, ``up = 0`` at init."""

    def __init__(self, in_features: int, out_features: int, rank: int = 4):
        super().__init__()
        if rank > min(in_features, out_features):
            raise ValueError(f"LoRA rank {rank} must be less or equal than {min(in_features, out_features)}")
        self.down = nn.Linear(in_features, rank, bias=False)
        self.up = nn.Linear(rank, out_features, bias=False)
        nn.init.normal_(self.down.weight, std=1 / rank)
        nn.init.zeros_(self.up.weight)

    def forward(self, hidden_states):
        w_dtype = self.down.weight.dtype
        return self.up(self.down(hidden_states.to(w_dtype))).to(hidden_states.dtype)

    def pair(self, like: torch.Tensor):
        """(down [r, in], up [out, r]) on ``like``'s device / dtype."""
        return (self.down.weight.detach().to(device=like.device, dtype=like.dtype),
                self.up.weight.detach().to(device=like.device, dtype=like.dtype))


class LoRAAttnProcessor_base(nn.Module):
    """Holder of the four LoRA layers of one attention module (the reference stores them on the
    attention's ``processor`` and the hook reads them from there, utils_lora.py:139-144)."""

    def __init__(self, hidden_size: int, cross_attention_dim=None, rank: int = 4):
        super().__init__()
        self.hidden_size, self.cross_attention_dim, self.rank = hidden_size, cross_attention_dim, rank
        kv_in = cross_attention_dim or hidden_size
        self.to_q_lora = LoRALinearLayer(hidden_size, hidden_size, rank)
        self.to_k_lora = LoRALinearLayer(kv_in, hidden_size, rank)
        self.to_v_lora = LoRALinearLayer(kv_in, hidden_size, rank)
        self.to_out_lora = LoRALinearLayer(hidden_size, hidden_size, rank)


def create_lora_diffusion_base(unet, rank: int = 4):
    """Attach a ``LoRAAttnProcessor_base`` as ``.processor`` to every attention of ``unet``."""
    for _, attn in unet.attention_modules():
        hidden = attn.to_q.out_features
        cross = attn.to_k.in_features if attn.is_cross else None
        attn.processor = LoRAAttnProcessor_base(hidden, cross, rank).to(attn.to_q.weight.device)
    return unet
