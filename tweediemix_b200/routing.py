"""Per-row weight routing installed on ``TmxAttention`` by the hook layer.

In the fused phase the reference sends batch row 0 (uncond) through the base attention weights and
row ``i+1`` through concept ``i``'s weights — Custom-Diffusion K/V matrices for cross-attention
(``fusion_generation/utils_custom.py:64-82``) or rank-4 LoRA deltas on q, k, v and the output
projection of every attention (``utils_lora.py:65-79,113-121``; layers per ``model_lora.py:28-48``).
A routing object carries those per-row weights in the packed layout ``TmxAttention.run`` consumes.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn.functional as F

_next_tag = [0]


def _tag() -> int:
    _next_tag[0] += 1
    return _next_tag[0]


class CustomRouting:
    """Row r uses ``kv_rows[r]`` = packed ``[to_k_i ; to_v_i]`` ([2*inner, cross_dim]); ``None`` = base."""
    kind = "custom"

    def __init__(self, kv_weights: List[Optional[torch.Tensor]]):
        self.kv_weights = kv_weights
        self.cache_tag = _tag()
        self._subsets = {}

    def subset(self, row_ids):
        """Routing for a rank that holds only ``row_ids`` of the full batch (concept-parallel path)."""
        key = tuple(row_ids)
        if key not in self._subsets:
            self._subsets[key] = CustomRouting([self.kv_weights[r] for r in key])
        return self._subsets[key]

    def kv_rows(self, attn, ehs: torch.Tensor) -> torch.Tensor:
        if ehs.shape[0] != len(self.kv_weights):
            raise RuntimeError(f"routing built for {len(self.kv_weights)} rows, got batch {ehs.shape[0]}")
        base = attn.packed_kv()
        out = torch.empty(ehs.shape[0], ehs.shape[1], base.shape[0], dtype=ehs.dtype, device=ehs.device)
        for r, w in enumerate(self.kv_weights):
            torch.matmul(ehs[r], (base if w is None else w).t(), out=out[r])
        return out


class LoRARows:
    """Packed rank-r deltas of ONE batch row of one attention module."""

    def __init__(self, q, k, v, out):
        # each argument: (down [r, in], up [out, r])
        (dq, uq), (dk, uk), (dv, uv), (do, uo) = q, k, v, out
        self.q_down, self.q_up_t = dq.contiguous(), uq.t().contiguous()
        self.out_down, self.out_up_t = do.contiguous(), uo.t().contiguous()
        self.kv_down = torch.cat([dk, dv]).contiguous()                       # [2r, in_kv]
        self.kv_up_t = torch.block_diag(uk, uv).t().contiguous()              # [2r, 2*inner]
        if dq.shape[1] == dk.shape[1]:                                        # self-attention: q, k, v share the input
            self.qkv_down = torch.cat([dq, dk, dv]).contiguous()              # [3r, d]
            self.qkv_up_t = torch.block_diag(uq, uk, uv).t().contiguous()     # [3r, 3*inner]
        else:
            self.qkv_down = self.qkv_up_t = None


class LoRARouting:
    """``rows[r]`` is a ``LoRARows`` or ``None`` (row 0, the unconditional row, is never routed)."""
    kind = "lora"

    def __init__(self, rows: List[Optional[LoRARows]]):
        self.rows = rows
        self.cache_tag = _tag()
        self._subsets = {}

    def subset(self, row_ids):
        key = tuple(row_ids)
        if key not in self._subsets:
            self._subsets[key] = LoRARouting([self.rows[r] for r in key])
        return self._subsets[key]

    def _check(self, batch: int):
        if batch != len(self.rows):
            raise RuntimeError(f"routing built for {len(self.rows)} rows, got batch {batch}")

    def add_qkv_self(self, attn, x, qkv):
        self._check(x.shape[0])
        for r, lr in enumerate(self.rows):
            if lr is not None:
                qkv[r].addmm_(F.linear(x[r], lr.qkv_down), lr.qkv_up_t)

    def add_q(self, attn, x, q):
        self._check(x.shape[0])
        for r, lr in enumerate(self.rows):
            if lr is not None:
                q[r].addmm_(F.linear(x[r], lr.q_down), lr.q_up_t)

    def kv_rows(self, attn, ehs):
        self._check(ehs.shape[0])
        kv = F.linear(ehs, attn.packed_kv())
        for r, lr in enumerate(self.rows):
            if lr is not None:
                kv[r].addmm_(F.linear(ehs[r], lr.kv_down), lr.kv_up_t)
        return kv

    def add_out(self, attn, a, o):
        """delta from the PRE-``to_out[0]`` tensor, added after its bias (utils_lora.py:113-119)."""
        self._check(a.shape[0])
        for r, lr in enumerate(self.rows):
            if lr is not None:
                o[r].addmm_(F.linear(a[r], lr.out_down), lr.out_up_t)
