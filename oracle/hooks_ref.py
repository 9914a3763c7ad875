"""Attention-hook restatement (oracle; test infrastructure only).

Restates, in one parameterised forward, what the reference's two hook files install on every
patched diffusers ``Attention`` module:

  * Custom-Diffusion variant — ``fusion_generation/utils_custom.py:45-158``: only ``attn2``
    (70 modules) is patched; while ``t`` is in the fusion window AND the call is cross-attention
    AND the batch is exactly 4, batch row 0 uses the base ``to_k``/``to_v`` and row ``i+1`` uses
    concept ``i``'s ``to_k_i``/``to_v_i`` (``:61-82``); otherwise shared weights (``:84-89``).
    Attention itself is einsum·scale → softmax → einsum → ``to_out[0]`` (``:91-106``); the hook
    never applies ``to_out[1]`` (dropout, identity).
  * LoRA variant — ``fusion_generation/utils_lora.py:47-218``: ``attn1`` AND ``attn2`` (140
    modules); gate has no ``is_cross`` term (``:63``); rows ``i+1`` get rank-4 deltas on q, k, v
    (``:65-79``) and on the output, computed from the PRE-``to_out[0]`` tensor and added after
    ``to_out[0]`` (bias included) and before ``to_out[1]`` (``:113-121``).
  * ``register_time`` — ``utils_custom.py:16-42`` / ``utils_lora.py:16-44``: stamps ``t``.
  * LoRA layer — ``fusion_generation/model_lora.py:28-48``: ``up(down(x))``, rank 4, no scale.

PINNED: ``tests/golden/make_golden.py`` runs the reference's own two files on seeded inputs in
the build container and stores their outputs; ``tests/test_oracle_hooks.py`` replays them here.

``gate`` defaults to the reference's literal 4 (quirk ⑥); callers that generalise to
``num_concepts + 1`` say so explicitly.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .unet_ref import transformer_blocks_in_hook_order

REFERENCE_GATE = 4


class LoRALinearRef(nn.Module):
    """model_lora.py:28-48."""

    def __init__(self, in_features: int, out_features: int, rank: int = 4):
        super().__init__()
        self.down = nn.Linear(in_features, rank, bias=False)
        self.up = nn.Linear(rank, out_features, bias=False)
        nn.init.normal_(self.down.weight, std=1 / rank)
        nn.init.zeros_(self.up.weight)

    def forward(self, x):
        return self.up(self.down(x.to(self.down.weight.dtype))).to(x.dtype)


def _in_window(t, window) -> bool:
    if isinstance(window, torch.Tensor):
        return bool((window == int(t)).any())
    return int(t) in window


def _sdpa_naive(mod, q, k, v):
    """utils_custom.py:91-105 — [B*h, N, d] naive attention, then back to [B, N, h*d]."""
    q, k, v = mod.head_to_batch_dim(q), mod.head_to_batch_dim(k), mod.head_to_batch_dim(v)
    sim = torch.einsum("b i d, b j d -> b i j", q, k) * mod.scale
    attn = sim.softmax(dim=-1)
    out = torch.einsum("b i j, b j d -> b i d", attn.to(v.dtype), v)
    return mod.batch_to_head_dim(out)


def custom_forward_ref(mod, x, encoder_hidden_states=None, gate: int = REFERENCE_GATE):
    """utils_custom.py:53-108."""
    is_cross = encoder_hidden_states is not None
    ctx = encoder_hidden_states if is_cross else x
    q = mod.to_q(x)
    if is_cross and _in_window(mod.t, mod.t_cond) and ctx.shape[0] == gate:
        k_rows = [mod.to_k(ctx[0:1])]
        v_rows = [mod.to_v(ctx[0:1])]
        for i in range(mod.num_concepts):
            k_rows.append(getattr(mod, f"to_k_{i}")(ctx[i + 1:i + 2]))
            v_rows.append(getattr(mod, f"to_v_{i}")(ctx[i + 1:i + 2]))
        k, v = torch.cat(k_rows), torch.cat(v_rows)
    else:
        k, v = mod.to_k(ctx), mod.to_v(ctx)
    return mod.to_out[0](_sdpa_naive(mod, q, k, v))


def lora_forward_ref(mod, x, encoder_hidden_states=None, gate: int = REFERENCE_GATE):
    """utils_lora.py:55-123."""
    ctx = encoder_hidden_states if encoder_hidden_states is not None else x
    routed = _in_window(mod.t, mod.t_cond) and ctx.shape[0] == gate
    q, k, v = mod.to_q(x), mod.to_k(ctx), mod.to_v(ctx)
    if routed:
        q, k, v = q.clone(), k.clone(), v.clone()
        for i in range(mod.num_concepts):
            r = slice(i + 1, i + 2)
            q[r] = q[r] + getattr(mod, f"to_q_{i}_lora")(x[r])
            k[r] = k[r] + getattr(mod, f"to_k_{i}_lora")(ctx[r])
            v[r] = v[r] + getattr(mod, f"to_v_{i}_lora")(ctx[r])
    pre = _sdpa_naive(mod, q, k, v)
    out = mod.to_out[0](pre)
    if routed:
        out = out.clone()
        for i in range(mod.num_concepts):
            r = slice(i + 1, i + 2)
            out[r] = out[r] + getattr(mod, f"to_out_{i}_lora")(pre[r])
    return mod.to_out[1](out)


def _patch(mod, fn, gate):
    def forward(x, encoder_hidden_states=None, attention_mask=None):
        assert attention_mask is None, "reference path never passes a mask (utils_custom.py:95-99 is dead)"
        return fn(mod, x, encoder_hidden_states, gate)
    mod.forward = forward


def register_custom_ref(unet, concept_unets, t_cond, num_concepts: int, gate: int = REFERENCE_GATE):
    """utils_custom.py:113-157.  ``concept_unets[i]`` plays ``model.unet_{i}``."""
    for name, blk in transformer_blocks_in_hook_order(unet):
        m = blk.attn2
        for i in range(num_concepts):
            donor = concept_unets[i].get_submodule(name + ".attn2")     # a full U-Net or any tree with these leaves
            setattr(m, f"to_k_{i}", donor.to_k)
            setattr(m, f"to_v_{i}", donor.to_v)
        m.t_cond, m.num_concepts = t_cond, num_concepts
        _patch(m, custom_forward_ref, gate)


def register_lora_ref(unet, concept_loras, t_cond, num_concepts: int, gate: int = REFERENCE_GATE):
    """utils_lora.py:134-217.  ``concept_loras[i][f"{block_name}.attn{1|2}"]`` is a dict with
    ``to_q_lora / to_k_lora / to_v_lora / to_out_lora`` LoRA layers (the processor of ``unet_i``)."""
    for name, blk in transformer_blocks_in_hook_order(unet):
        for which in ("attn2", "attn1"):
            m = getattr(blk, which)
            for i in range(num_concepts):
                proc = concept_loras[i][f"{name}.{which}"]
                for p in ("q", "k", "v", "out"):
                    setattr(m, f"to_{p}_{i}_lora", proc[f"to_{p}_lora"])
            m.t_cond, m.num_concepts = t_cond, num_concepts
            _patch(m, lora_forward_ref, gate)


def register_time_ref(unet, t, lora: bool = False):
    """utils_custom.py:16-42 (attn2 only) / utils_lora.py:16-44 (attn1 and attn2)."""
    for _, blk in transformer_blocks_in_hook_order(unet):
        blk.attn2.t = t
        if lora:
            blk.attn1.t = t


def make_lora_set(unet, seed: int, up_std: float = 1e-2, rank: int = 4):
    """One concept's LoRA processors for every attention, seeded; ``up`` is non-zero so that the
    routing is observable (SURVEY §8d synthetic inputs)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, blk in transformer_blocks_in_hook_order(unet):
        for which in ("attn1", "attn2"):
            m = getattr(blk, which)
            d = m.to_q.in_features
            dk = m.to_k.in_features
            inner = m.to_q.out_features
            layers = {"to_q_lora": LoRALinearRef(d, inner, rank), "to_k_lora": LoRALinearRef(dk, inner, rank),
                      "to_v_lora": LoRALinearRef(dk, inner, rank), "to_out_lora": LoRALinearRef(inner, d, rank)}
            for l in layers.values():
                with torch.no_grad():
                    l.down.weight.copy_(torch.randn(l.down.weight.shape, generator=g) / rank)
                    l.up.weight.copy_(torch.randn(l.up.weight.shape, generator=g) * up_std)
            out[f"{name}.{which}"] = layers
    return out
