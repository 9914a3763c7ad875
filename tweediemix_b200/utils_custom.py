"""Hook API of the Custom-Diffusion variant — drop-in for ``fusion_generation/utils_custom.py``.

Same three entry points, same argument meaning, same attributes left on the modules:

  ``seed_everything(seed)``                                      utils_custom.py:10-14
  ``register_time(model, t)``                                    utils_custom.py:16-42
  ``register_attention_control_efficient(model, t_cond, num_concepts)``   utils_custom.py:45-158

``model`` is anything with ``.unet`` and ``.unet_{i}`` attributes (the reference's ``Tweediemix``).
After registration every ``attn2`` of the 70 transformer blocks carries ``t_cond``,
``num_concepts``, ``to_k_{i}`` / ``to_v_{i}`` and an instance-level
``forward(x, encoder_hidden_states=None, attention_mask=None)``, exactly like the reference; the
body of that forward runs on the tmx kernels (cached routed K/V projection + tcgen05 attention).

Differences, all deliberate:
  * the routing gate is ``batch == num_concepts + 1`` instead of the literal ``4``
    (utils_custom.py:61-62), which is the same thing for the only configuration the reference can
    run (3 concepts) and lets K != 3 work; ``gate=4`` restores the literal;
  * ``t in t_cond`` is evaluated against a host ``frozenset`` built once, not against a CUDA tensor
    (one device sync per module per step in the reference);
  * ``unet_{i}`` may be a full U-Net or any module tree that has the ``attn2.to_k/to_v`` leaves;
  * ``model.unet`` may be the product's ``TmxUNet2DConditionModel`` OR a diffusers-shaped U-Net (the tree the reference
    patches, ``utils_custom.py:113-158``): its ``Attention`` modules keep their parameters and get the same instance-level
    ``forward`` — running the tmx kernels through a ``TmxAttentionView`` — so the hook layer drops into a diffusers pipeline
    without replacing the U-Net (INTEGRATION.md path B; ``tests/test_adapter.py``).
"""
from __future__ import annotations

import random

import numpy as np
import torch

from .routing import CustomRouting
from .unet import TmxAttentionView, iter_transformer_blocks


def seed_everything(seed):
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    random.seed(seed)
    np.random.seed(seed)


def _cross_attentions(unet):
    for name, blk in iter_transformer_blocks(unet):
        yield name + ".attn2", blk.attn2


def as_window(t_cond) -> frozenset:
    if torch.is_tensor(t_cond):
        t_cond = t_cond.tolist()
    return frozenset(int(v) for v in t_cond)


def register_time(model, t):
    t = int(t)
    for _, attn in _cross_attentions(model.unet):
        attn.t = t


def _packed_kv_of(donor_attn, like: torch.Tensor) -> torch.Tensor:
    w = torch.cat([donor_attn.to_k.weight, donor_attn.to_v.weight]).detach()
    return w.to(device=like.device, dtype=like.dtype).contiguous()


def register_attention_control_efficient(model, t_cond, num_concepts, gate=None):
    gate = num_concepts + 1 if gate is None else int(gate)
    window = as_window(t_cond)
    donors = [getattr(model, f"unet_{i}") for i in range(num_concepts)]

    def install(attn, name: str):
        # `attn` is the product's TmxAttention or a foreign diffusers-shaped Attention; `core` runs the arithmetic
        core = TmxAttentionView.of(attn, is_cross=True)
        rows = [None]
        for i, donor in enumerate(donors):
            d = donor.get_submodule(name)
            setattr(attn, f"to_k_{i}", d.to_k)
            setattr(attn, f"to_v_{i}", d.to_v)
            rows.append(_packed_kv_of(d, attn.to_k.weight))
        routing = CustomRouting(rows)
        attn.routing = routing
        attn.t_cond = t_cond
        attn.num_concepts = num_concepts
        attn.fusion_window = window

        def forward(x, encoder_hidden_states=None, attention_mask=None, residual=None):
            if attention_mask is not None:
                raise RuntimeError("attention_mask is not supported (dead branch in the reference, utils_custom.py:95-99)")
            local = getattr(attn, "local_rows", None)        # concept-parallel: this rank's rows of the gate-sized batch
            routed = (encoder_hidden_states is not None and attn.t in attn.fusion_window
                      and encoder_hidden_states.shape[0] == (gate if local is None else len(local)))
            if not routed:
                return core.run(x, encoder_hidden_states, None, residual)
            return core.run(x, encoder_hidden_states, routing if local is None else routing.subset(local), residual)

        attn.forward = forward

    for name, attn in _cross_attentions(model.unet):
        install(attn, name)
