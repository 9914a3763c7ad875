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

import os

from . import ops

# 'k3' (default): one tmx_routed_linear_fwd launch per projection for all routed rows; 'cublas': two skinny cuBLAS GEMMs
# per routed row (kept for A/B measurements: profiles/README.md).
LORA_IMPL = os.environ.get("TMX_LORA_IMPL", "k3")

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
        # k3: ONE grouped tcgen05 GEMM over the batch rows, each with its own [2*inner, cross_dim] weight
        return ops.routed_linear(ehs.contiguous(), [base if w is None else w for w in self.kv_weights])


class LoRARows:
    """Packed rank-r deltas of ONE batch row of one attention module."""

    def __init__(self, q, k, v, out):
        # each argument: (down [r, in], up [out, r])
        (dq, uq), (dk, uk), (dv, uv), (do, uo) = q, k, v, out
        # k3 layout: down [nseg*r, in] (segments stacked), up [Nout, r] (segments stacked along the output columns)
        self.q_down, self.q_up = dq.contiguous(), uq.contiguous()
        self.out_down, self.out_up = do.contiguous(), uo.contiguous()
        self.kv_down = torch.cat([dk, dv]).contiguous()                       # [2r, in_kv]
        self.kv_up = torch.cat([uk, uv]).contiguous()                         # [2*inner, r]
        if dq.shape[1] == dk.shape[1]:                                        # self-attention: q, k, v share the input
            self.qkv_down = torch.cat([dq, dk, dv]).contiguous()              # [3r, d]
            self.qkv_up = torch.cat([uq, uk, uv]).contiguous()                # [3*inner, r]
        else:
            self.qkv_down = self.qkv_up = None


class LoRARouting:
    """``rows[r]`` is a ``LoRARows`` or ``None`` (row 0, the unconditional row, is never routed)."""
    kind = "lora"

    def __init__(self, rows: List[Optional[LoRARows]]):
        self.rows = rows
        self.cache_tag = _tag()
        self._subsets = {}
        self._lists = {}

    def subset(self, row_ids):
        key = tuple(row_ids)
        if key not in self._subsets:
            self._subsets[key] = LoRARouting([self.rows[r] for r in key])
        return self._subsets[key]

    def _check(self, batch: int):
        if batch != len(self.rows):
            raise RuntimeError(f"routing built for {len(self.rows)} rows, got batch {batch}")

    def _factors(self, which: str):
        """(downs, ups) per batch row for projection ``which`` ('qkv', 'q', 'kv', 'out'); ``None`` entries = not routed."""
        got = self._lists.get(which)
        if got is None:
            got = ([None if lr is None else getattr(lr, which + "_down") for lr in self.rows],
                   [None if lr is None else getattr(lr, which + "_up") for lr in self.rows])
            self._lists[which] = got
        return got

    def _ups_ext(self, which: str, nseg: int):
        """Per batch row, the up factors of projection ``which`` as the B operand of the GEMM's K = 16 tail step: [Nout, 64]
        with segment s's [seg, r] block in columns [s*r, (s+1)*r) of its own rows, zeros elsewhere."""
        key = which + "_ext"
        got = self._lists.get(key)
        if got is None:
            _, ups = self._factors(which)
            got = []
            for u in ups:
                if u is None:
                    got.append(None)
                    continue
                nout, r = u.shape
                seg = nout // nseg
                e = torch.zeros(nout, 64, dtype=u.dtype, device=u.device)
                for s_ in range(nseg):
                    e[s_ * seg:(s_ + 1) * seg, s_ * r:(s_ + 1) * r] = u[s_ * seg:(s_ + 1) * seg]
                got.append(e)
            self._lists[key] = got
        return got

    def tail(self, which: str, nseg: int, x):
        """LoRA deltas of projection ``which`` as a fused tail of the projection GEMM (``ops.linear(lora_tail=...)``):
        launches the skinny t = x @ down^T kernel and returns (t, ups, rows_per_batch) — or None when the shape is off the
        fused path (token count per batch row not a multiple of the 128-row tile, rank layout beyond 16 slots), in which
        case the caller adds the deltas with the stand-alone k3 kernel instead."""
        self._check(x.shape[0])
        downs, ups = self._factors(which)
        routed = [d for d in downs if d is not None]
        if not routed or LORA_IMPL != "k3":
            return None
        sr = routed[0].shape[0]
        if x.dim() != 3 or x.shape[1] % 128 != 0 or sr not in (4, 8, 12, 16) or not x.is_contiguous() or x.shape[0] > 16:
            return None
        t = ops.lora_t(x, downs, sr)
        return t, self._ups_ext(which, nseg), x.shape[1]

    def _add(self, which: str, nseg: int, x, y):
        """y[r] += segment-wise (x[r] @ down_r^T) @ up_r^T for every routed row r: ONE k3 launch for the whole batch."""
        self._check(x.shape[0])
        downs, ups = self._factors(which)
        if all(d is None for d in downs):
            return y
        if not (x.is_contiguous() and y.is_contiguous()):
            raise RuntimeError("LoRA routing needs contiguous activations")
        if LORA_IMPL == "cublas":
            seg = y.shape[-1] // nseg
            for r, (d, u) in enumerate(zip(downs, ups)):
                if d is None:
                    continue
                t = F.linear(x[r], d)                                         # [M, nseg*rank]
                rank = d.shape[0] // nseg
                for s_ in range(nseg):
                    y[r, :, s_ * seg:(s_ + 1) * seg].addmm_(t[:, s_ * rank:(s_ + 1) * rank], u[s_ * seg:(s_ + 1) * seg].t())
            return y
        return ops.routed_linear(x, None, downs, ups, nseg=nseg, out=y)

    def add_qkv_self(self, attn, x, qkv):
        self._add("qkv", 3, x, qkv)

    def add_q(self, attn, x, q):
        self._add("q", 1, x, q)

    def kv_rows(self, attn, ehs):
        ehs = ehs.contiguous()
        kv = F.linear(ehs, attn.packed_kv())
        return self._add("kv", 2, ehs, kv)

    def add_out(self, attn, a, o):
        """delta from the PRE-``to_out[0]`` tensor, added after its bias (utils_lora.py:113-119)."""
        self._add("out", 1, a, o)
