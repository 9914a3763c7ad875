#!/usr/bin/env python
"""Numerics check of the attention kernel under a given tmx_attn_set_variant() value (used for opt-in build variants):
max-abs error vs an fp32 softmax reference at the SDXL self-attention shapes and a cross-attention shape."""
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from tweediemix_b200 import _lib, build, ops  # noqa: E402


def main():
    build.build()
    v = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    assert _lib.load().tmx_attn_set_variant(v) == 0
    g = torch.Generator().manual_seed(0)
    worst = 0.0
    # the B = 1 and ragged shapes cut units along K/V into 2-4 pieces (stream-K merge path), 300 rows = an odd tile count (lone last tile)
    for B, H, Nq, Nk in [(2, 10, 4096, 4096), (4, 20, 1024, 1024), (2, 20, 1024, 77), (1, 5, 200, 333), (1, 20, 1024, 1024), (1, 10, 4096, 4096),
                         (1, 3, 300, 700), (2, 2, 128, 2000), (2, 20, 1024, 1024), (4, 20, 1024, 77), (4, 10, 4096, 77), (1, 3, 200, 77), (2, 5, 130, 65), (1, 2, 64, 80)]:
        for dt in (torch.bfloat16, torch.float16):
            q, k, v_ = (torch.randn(B, n, H * 64, generator=g).to(dt).cuda() for n in (Nq, Nk, Nk))
            o = ops.attention(q, k, v_, H)
            qh, kh, vh = (t.float().reshape(B, -1, H, 64).permute(0, 2, 1, 3) for t in (q, k, v_))
            ref = torch.softmax(qh @ kh.transpose(-1, -2) * 0.125, dim=-1) @ vh
            err = (o.float().reshape(B, Nq, H, 64).permute(0, 2, 1, 3) - ref).abs().max().item()
            tol = 3e-2 if dt == torch.bfloat16 else 4e-3
            print(f"variant {v} B{B} H{H} Nq{Nq} Nk{Nk} {dt}: max|err| = {err:.3e} (tol {tol})")
            assert err <= tol
            worst = max(worst, err)
    print("attn_check ok, worst", worst)


if __name__ == "__main__":
    main()
