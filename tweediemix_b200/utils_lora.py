"""Hook API of the LoRA variant — drop-in for ``fusion_generation/utils_lora.py``.

  ``register_time(model, t)``                                    utils_lora.py:16-44   (attn1 AND attn2)
  ``register_attention_control_efficient(model, t_cond, num_concepts)``   utils_lora.py:47-218

Every ``attn1`` and ``attn2`` of the 70 transformer blocks (140 modules) gets
``to_{q,k,v,out}_{i}_lora`` taken from ``model.unet_{i}``'s ``<attention>.processor``, ``t_cond``,
``num_concepts`` and an instance-level forward.  While ``t`` is in the window and the batch equals
the gate, row ``i+1`` adds concept ``i``'s rank-4 deltas to q, k, v (utils_lora.py:65-79) and to
the output — computed from the pre-``to_out[0]`` tensor, added after its bias
(utils_lora.py:113-121).  The gate has no ``is_cross`` term (utils_lora.py:63): self-attention is
routed too.  The naive N x N attention of the reference's self-attention path (1.3 GB of scores
per call at 1024²) is replaced by the tcgen05 flash kernel.

Gate / window handling as in ``utils_custom`` (``gate=None`` -> ``num_concepts + 1``).
"""
from __future__ import annotations

from .routing import LoRARouting, LoRARows
from .unet import TmxAttentionView, iter_transformer_blocks
from .utils_custom import as_window, seed_everything  # noqa: F401  (same helper in both reference files)


def _pair(layer, like):
    """(down [r, in], up [out, r]) of a LoRA layer — the product's ``LoRALinearLayer`` or anything with ``.down`` /
    ``.up`` Linear children (the reference's ``model_lora.LoRALinearLayer``) — on ``like``'s device / dtype."""
    return (layer.down.weight.detach().to(device=like.device, dtype=like.dtype),
            layer.up.weight.detach().to(device=like.device, dtype=like.dtype))


def _all_attentions(unet):
    for name, blk in iter_transformer_blocks(unet):
        yield name + ".attn2", blk.attn2
        yield name + ".attn1", blk.attn1


def register_time(model, t):
    t = int(t)
    for _, attn in _all_attentions(model.unet):
        attn.t = t


def register_attention_control_efficient(model, t_cond, num_concepts, gate=None):
    gate = num_concepts + 1 if gate is None else int(gate)
    window = as_window(t_cond)
    donors = [getattr(model, f"unet_{i}") for i in range(num_concepts)]

    def install(attn, name: str):
        # `attn` is the product's TmxAttention or a foreign diffusers-shaped Attention; `core` runs the arithmetic
        core = TmxAttentionView.of(attn, is_cross=name.endswith("attn2"))
        like = attn.to_q.weight
        rows = [None]
        for i, donor in enumerate(donors):
            proc = donor.get_submodule(name).processor
            for p in ("q", "k", "v", "out"):
                setattr(attn, f"to_{p}_{i}_lora", getattr(proc, f"to_{p}_lora"))
            rows.append(LoRARows(_pair(proc.to_q_lora, like), _pair(proc.to_k_lora, like),
                                 _pair(proc.to_v_lora, like), _pair(proc.to_out_lora, like)))
        routing = LoRARouting(rows)
        attn.routing = routing
        attn.t_cond = t_cond
        attn.num_concepts = num_concepts
        attn.fusion_window = window

        def forward(x, encoder_hidden_states=None, attention_mask=None, residual=None):
            if attention_mask is not None:
                raise RuntimeError("attention_mask is not supported (dead branch in the reference, utils_lora.py:103-107)")
            batch = (x if encoder_hidden_states is None else encoder_hidden_states).shape[0]
            local = getattr(attn, "local_rows", None)        # concept-parallel: this rank's rows of the gate-sized batch
            routed = attn.t in attn.fusion_window and batch == (gate if local is None else len(local))
            if not routed:
                return core.run(x, encoder_hidden_states, None, residual)
            return core.run(x, encoder_hidden_states, routing if local is None else routing.subset(local), residual)

        attn.forward = forward

    for name, attn in _all_attentions(model.unet):
        install(attn, name)
