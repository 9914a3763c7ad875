#!/usr/bin/env python
"""Small-shape sweep of every tmx kernel for `compute-sanitizer --tool memcheck` (GPU box):
    compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py
Ragged shapes on purpose (partial tiles, odd row counts); results are also checked loosely so a wrong answer is not silent."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tweediemix_b200 import build, ops  # noqa: E402


def main():
    build.build()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *s: torch.randn(*s, generator=g, device=dev)
    bf = torch.bfloat16
    # attention: partial q tile, partial kv tile, cross length 77, several heads / batches
    for (B, H, Nq, Nk) in [(1, 1, 1, 1), (1, 2, 200, 77), (2, 3, 130, 129), (1, 1, 384, 256)]:
        q, k, v = rnd(B, Nq, H * 64).to(bf), rnd(B, Nk, H * 64).to(bf), rnd(B, Nk, H * 64).to(bf)
        o = ops.attention(q, k, v, H)
        ref = torch.nn.functional.scaled_dot_product_attention(
            q.view(B, Nq, H, 64).transpose(1, 2).float(), k.view(B, Nk, H, 64).transpose(1, 2).float(), v.view(B, Nk, H, 64).transpose(1, 2).float())
        assert (o.float() - ref.transpose(1, 2).reshape(B, Nq, H * 64)).abs().max() < 3e-2
    # stream-K schedule forced on: units of 6 K/V tiles cut into 3 pieces each (dump -> flag -> merge path)
    from tweediemix_b200 import _lib as _l
    _l.load().tmx_attn_set_variant(22)
    for (B, H, Nq, Nk) in [(1, 2, 200, 700), (1, 1, 300, 333)]:
        q, k, v = rnd(B, Nq, H * 64).to(bf), rnd(B, Nk, H * 64).to(bf), rnd(B, Nk, H * 64).to(bf)
        o = ops.attention(q, k, v, H)
        ref = torch.nn.functional.scaled_dot_product_attention(
            q.view(B, Nq, H, 64).transpose(1, 2).float(), k.view(B, Nk, H, 64).transpose(1, 2).float(), v.view(B, Nk, H, 64).transpose(1, 2).float())
        assert (o.float() - ref.transpose(1, 2).reshape(B, Nq, H * 64)).abs().max() < 3e-2
    _l.load().tmx_attn_set_variant(0)
    # GroupNorm: fused (NHWC 16-bit), two-pass (fp32), NCHW
    for shape, dt, cl in [((2, 64, 5, 7), bf, True), ((3, 320, 9, 9), bf, True), ((1, 32, 1, 8), torch.float32, True), ((2, 64, 4, 6), bf, False)]:
        x = rnd(*shape).to(dt)
        if cl:
            x = x.contiguous(memory_format=torch.channels_last)
        y = ops.group_norm(x, torch.ones(shape[1], device=dev), torch.zeros(shape[1], device=dev), 32, 1e-5, silu=True)
        ref = torch.nn.functional.silu(torch.nn.functional.group_norm(x.float(), 32))
        assert (y.float() - ref).abs().max() < 3e-2
    # two-source GroupNorm (cat-free up-block ResNets): fused and two-pass paths, group straddling the source boundary
    for shp in [(2, 64, 128, 5, 7), (1, 8, 56, 3, 3)]:
        xa = rnd(shp[0], shp[1], shp[3], shp[4]).to(bf).contiguous(memory_format=torch.channels_last)
        xb = rnd(shp[0], shp[2], shp[3], shp[4]).to(bf).contiguous(memory_format=torch.channels_last)
        Cc = shp[1] + shp[2]
        y = ops.group_norm(xa, torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev), 32, 1e-5, silu=True, x2=xb)
        ref = torch.nn.functional.silu(torch.nn.functional.group_norm(torch.cat([xa, xb], 1).float(), 32))
        assert (y.float() - ref).abs().max() < 3e-2
    # LayerNorm, fused residual + LayerNorm, residual add, GEGLU
    x = rnd(3, 5, 64).to(bf)
    ops.layer_norm(x, torch.ones(64, device=dev), torch.zeros(64, device=dev), 1e-5)
    ops.residual_add_layer_norm(x.clone(), x, torch.ones(64, device=dev), torch.zeros(64, device=dev), 1e-5)
    ops.residual_add(x, x)
    ops.geglu(rnd(7, 3, 48).to(bf))
    # k13 / k14: channel concat and nearest 2x upsample (NHWC)
    ca = rnd(2, 8, 3, 5).to(bf).contiguous(memory_format=torch.channels_last)
    cb = rnd(2, 24, 3, 5).to(bf).contiguous(memory_format=torch.channels_last)
    assert torch.equal(ops.cat_channels(ca, cb), torch.cat([ca, cb], 1))
    assert torch.equal(ops.upsample_nearest2x(cb), torch.nn.functional.interpolate(cb, scale_factor=2.0, mode="nearest"))
    # k7 and its sharded halves
    xl = rnd(2, 4, 8, 8)
    eps = rnd(2, 4, 4, 8, 8).to(bf)
    m = (torch.rand(3, 1, 8, 8, generator=g, device=dev) > 0.5).float()
    ops.tweedie_blend_ddim(xl, eps, m, 0.3, 0.4, 0.8)
    acc = torch.empty(2, 2, 4, 8, 8, device=dev)
    ops.blend_partial(eps[:, 1:3].contiguous(), m, [1, 2], acc, 2)
    ops.blend_finish(xl, acc, m, 0.3, 0.4, 0.8)
    # k3: grouped GEMM with ragged M / Nout, LoRA deltas with an unrouted row
    x = rnd(3, 130, 128).to(bf)
    ws = [(rnd(136, 128) / 11).to(bf) for _ in range(3)]
    y = ops.routed_linear(x, ws)
    assert (y.float() - torch.stack([x[b].float() @ ws[b].float().t() for b in range(3)])).abs().max() < 5e-2
    downs = [None] + [(rnd(4, 128) / 4).to(bf) for _ in range(2)]
    ups = [None] + [(rnd(136, 4) * 0.05).to(bf) for _ in range(2)]
    ops.routed_linear(x, None, downs, ups, nseg=1, out=y)
    x2, y2 = rnd(2, 37, 64).to(bf), rnd(2, 37, 48).to(bf)
    ops.routed_linear(x2, None, [None, (rnd(12, 64) / 4).to(bf)], [None, (rnd(48, 4) * 0.05).to(bf)], nseg=3, out=y2)
    # k10: ragged M / N tiles, both tile widths, bias + residual, GEGLU, LoRA tail (+ the t = x . down^T kernel)
    from tweediemix_b200 import _lib
    from tweediemix_b200.routing import LoRARouting
    for bn in (128, 192, 256, 320):
        assert _lib.load().tmx_linear_set_variant(bn) == 0
        xa, wa = rnd(300, 192).to(bf), (rnd(328, 192) / 14).to(bf)
        ba, ra = rnd(328), rnd(300, 328).to(bf)
        ya = ops.linear(xa, wa, ba, residual=ra)
        assert (ya.float() - (xa.float() @ wa.float().t() + ba + ra.float())).abs().max() < 6e-2
        wg = (rnd(192, 64) / 8).to(bf)
        idx = ops.geglu_interleave_index(96, dev)
        yg = ops.linear(rnd(70, 64).to(bf), wg[idx].contiguous(), rnd(192)[idx].contiguous(), geglu=True)
        assert yg.shape == (70, 96) and torch.isfinite(yg).all()
        xl_ = rnd(3, 128, 64).to(bf)
        dn = [None] + [(rnd(8, 64) / 4).to(bf) for _ in range(2)]
        up = [None] + [(rnd(144, 4) * 0.05).to(bf) for _ in range(2)]
        rt = LoRARouting.__new__(LoRARouting)
        rt.rows, rt._lists, rt._subsets, rt.cache_tag = [None] * 3, {"w": (dn, up)}, {}, 0
        wl = (rnd(144, 64) / 8).to(bf)
        yl = ops.linear(xl_, wl, lora_tail=rt.tail("w", 2, xl_))
        want = xl_.float() @ wl.float().t()
        for b in (1, 2):
            t = xl_[b].float() @ dn[b].float().t()
            for s_ in range(2):
                want[b, :, s_ * 72:(s_ + 1) * 72] += t[:, s_ * 4:(s_ + 1) * 4] @ up[b][s_ * 72:(s_ + 1) * 72].float().t()
        assert (yl.float() - want).abs().max() < 6e-2
    # k11 / k12: video step and frame-0 injection
    lat, vv = rnd(1, 4, 16, 3, 4).to(bf), rnd(2, 4, 16, 3, 4).to(bf)
    ops.vpred_cfg_ddim(lat, vv[:1].contiguous(), vv[1:].contiguous(), 0.3, 0.4, 9.0, x0_out=torch.empty_like(lat), ref_rounding=True)
    ops.frame_inject(rnd(32, 8, 3, 3).to(bf), 2, 16, 0.7)
    ops.frame_inject(rnd(32, 8, 3, 3).to(bf), 2, 16, 1.0, ref_rounding=True)
    xs_, ws_ = rnd(256, 2560).to(bf), (rnd(256, 2560) / 50).to(bf)          # 2 tiles, K = 2560: split-K tail (cooperative launch)
    _lib.load().tmx_linear_set_variant(0)
    ys_ = ops.linear(xs_, ws_, rnd(256), residual=rnd(256, 256).to(bf))
    assert torch.isfinite(ys_).all()
    torch.cuda.synchronize()
    print("sanitize_small: all kernels ran")


if __name__ == "__main__":
    main()
