#!/usr/bin/env python
"""Numerics of the batch-size-dependent kernel paths at the per-rank batch of configs[3] on 8 GPUs (5 units per rank) and at odd
batches: stream-K attention, k2t, slab GroupNorm (N x G = 160 CTAs > #SM)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tweediemix_b200 import build, ops  # noqa: E402


def main():
    build.build()
    g = torch.Generator(device="cuda").manual_seed(0)
    bf = torch.bfloat16
    for B in (5, 3, 7):
        for H, Nq, Nk in [(20, 1024, 1024), (10, 4096, 4096), (20, 1024, 77), (10, 4096, 77)]:
            q = torch.randn(B, Nq, H * 64, generator=g, device="cuda").to(bf)
            k = torch.randn(B, Nk, H * 64, generator=g, device="cuda").to(bf)
            v = torch.randn(B, Nk, H * 64, generator=g, device="cuda").to(bf)
            o = ops.attention(q, k, v, H)
            sp = lambda t, n: t.view(B, n, H, 64).transpose(1, 2).float()
            ref = F.scaled_dot_product_attention(sp(q, Nq), sp(k, Nk), sp(v, Nk)).transpose(1, 2).reshape(B, Nq, H * 64)
            err = (o.float() - ref).abs().max().item()
            print(f"attention B{B} H{H} Nq{Nq} Nk{Nk}: max|err| {err:.3e}")
            assert err <= 3e-2
        for C, hw in [(1280, 32), (640, 64), (320, 128), (2560, 32), (1280, 64)]:
            x = (torch.randn(B, C, hw, hw, generator=g, device="cuda") * 1.5 + 0.3).to(bf).contiguous(memory_format=torch.channels_last)
            gm, bt = 1 + 0.2 * torch.randn(C, generator=g, device="cuda"), 0.1 * torch.randn(C, generator=g, device="cuda")
            y = ops.group_norm(x, gm, bt, 32, 1e-5, silu=True)
            ref = F.silu(F.group_norm(x.float(), 32, gm, bt, 1e-5))
            err = (y.float() - ref).abs().max().item()
            print(f"groupnorm B{B} C{C} {hw}x{hw}: max|err| {err:.3e}")
            assert err <= 6e-2
    print("check_b5 ok")


if __name__ == "__main__":
    main()
