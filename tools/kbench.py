#!/usr/bin/env python
"""Micro-benchmark of every tmx kernel at the shapes one K=3 fused step of SDXL 1024² launches them with.

    python tools/kbench.py [--only attention,groupnorm,...] [--reps 20]

Each shape is timed with CUDA events over `reps` back-to-back launches on rotating buffers whose total
footprint exceeds the 126 MB L2 (so bandwidth kernels see HBM, not L2), after 3 warm-up launches.
Prints per-shape microseconds and achieved GB/s or TFLOP/s against MEASURED_PEAKS.json.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tweediemix_b200 import build, ops  # noqa: E402

B = 4
DT = torch.bfloat16


def peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return d["hbm_gbs"], d["bf16_tflops"]
    except Exception:
        return 6650.0, 1590.0


def timeit(fn, nbuf, reps):
    """Device time per launch: the `reps` launches are captured in ONE CUDA graph (as the sampler replays
    them), so the Python / ctypes launch cost (~10 us per call in eager mode) is not what is measured."""
    for i in range(3):
        fn(i % nbuf)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(reps):
                fn(i % nbuf)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps          # us


def nbuf_for(nbytes):
    return max(2, min(64, int(300e6 // max(nbytes, 1)) + 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--attn-variant", type=int, default=0, help="tmx_attn_set_variant() value (11 / 12 = one / two softmax threads per row)")
    ap.add_argument("--gn-variant", type=int, default=0, help="tmx_groupnorm_set_variant(): 0 default (per-group slab kernel, else fused cooperative, else two launches), 1 two launches, 2 fused plain launch, 3 per-group slab kernel off")
    ap.add_argument("--batch", type=int, default=4, help="U-Net batch rows (4 = one K=3 fused step on one GPU; 2 / 1 = the per-rank batch of a 2- / 4-rank concept-parallel group)")
    ap.add_argument("--shapes", default="", help="comma-separated substrings; only shapes whose tag contains one are run (profiling aid)")
    ap.add_argument("--compare", action="store_true",
                    help="also time the same-box library kernels for the same shapes (measurement only, never a product dependency): "
                         "F.scaled_dot_product_attention, flash-attn, F.group_norm+F.silu, F.layer_norm, F.linear (cuBLAS)")
    args = ap.parse_args()
    global B
    B = args.batch
    if args.gn_variant:
        from tweediemix_b200 import _lib
        build.build()
        assert _lib.load().tmx_groupnorm_set_variant(args.gn_variant) == 0
    if args.attn_variant:
        from tweediemix_b200 import _lib
        build.build()
        assert _lib.load().tmx_attn_set_variant(args.attn_variant) == 0
    only = set(args.only.split(",")) if args.only else None
    build.build()
    hbm, tf = peaks()
    dev = "cuda"
    out = []

    def want(name):
        return only is None or name in only

    shapes = [t for t in args.shapes.split(",") if t]

    def keep(tag):
        return not shapes or any(t in tag for t in shapes)

    if want("groupnorm"):
        # (C, HW, count per forward): ResNet norm1/norm2 + Transformer2D norm + conv_norm_out sites
        sites = [(320, 16384, 5), (640, 16384, 2), (960, 16384, 1), (320, 4096, 1), (640, 4096, 9), (1280, 4096, 1), (1920, 4096, 1), (960, 4096, 1),
                 (640, 1024, 1), (1280, 1024, 20), (2560, 1024, 2), (1920, 1024, 1)]
        for C, HW, cnt in sites:
            if not keep(f"C{C}_HW{HW}"):
                continue
            h = int(HW ** 0.5)
            nb = nbuf_for(B * C * HW * 2 * 2)
            xs = [torch.randn(B, C, h, h, device=dev, dtype=DT).contiguous(memory_format=torch.channels_last) for _ in range(nb)]
            ys = [torch.empty_like(x) for x in xs]
            g, b_ = torch.ones(C, device=dev), torch.zeros(C, device=dev)
            us = timeit(lambda i: ops.group_norm(xs[i], g, b_, 32, 1e-5, silu=True, out=ys[i]), nb, args.reps)
            nbytes = 2 * B * C * HW * 2
            out.append(("groupnorm", f"C{C}_HW{HW} x{cnt}", us, nbytes / us / 1e3, "GB/s", hbm))
    if want("layernorm"):
        for N, D, cnt in [(4096, 640, 30), (1024, 1280, 180)]:
            if not keep(f"N{N}_D{D}"):
                continue
            nb = nbuf_for(B * N * D * 4)
            xs = [torch.randn(B, N, D, device=dev, dtype=DT) for _ in range(nb)]
            ys = [torch.empty_like(x) for x in xs]
            g, b_ = torch.ones(D, device=dev), torch.zeros(D, device=dev)
            us = timeit(lambda i: ops.layer_norm(xs[i], g, b_, 1e-5, out=ys[i]), nb, args.reps)
            out.append(("layernorm", f"N{N}_D{D} x{cnt}", us, 2 * B * N * D * 2 / us / 1e3, "GB/s", hbm))
    if want("resadd"):
        for n, cnt in [(B * 4096 * 640, 40), (B * 1024 * 1280, 190), (B * 16384 * 320, 8)]:
            nb = nbuf_for(n * 6)
            a = [torch.randn(n, device=dev, dtype=DT) for _ in range(nb)]
            b_ = [torch.randn(n, device=dev, dtype=DT) for _ in range(nb)]
            us = timeit(lambda i: ops.residual_add(a[i], b_[i], out=a[i]), nb, args.reps)
            out.append(("resadd", f"n{n} x{cnt}", us, 3 * n * 2 / us / 1e3, "GB/s", hbm))
    if want("resadd_ln"):
        for N, D, cnt in [(4096, 640, 30), (1024, 1280, 180)]:
            if not keep(f"N{N}_D{D}"):
                continue
            nb = nbuf_for(B * N * D * 8)
            a = [torch.randn(B, N, D, device=dev, dtype=DT) for _ in range(nb)]
            b_ = [torch.randn(B, N, D, device=dev, dtype=DT) for _ in range(nb)]
            ns = [torch.empty(B, N, D, device=dev, dtype=DT) for _ in range(nb)]
            g, bb = torch.ones(D, device=dev), torch.zeros(D, device=dev)
            us = timeit(lambda i: ops.residual_add_layer_norm(a[i], b_[i], g, bb, 1e-5, h_out=a[i], n_out=ns[i]), nb, args.reps)
            out.append(("resadd_ln", f"N{N}_D{D} x{cnt}", us, 4 * B * N * D * 2 / us / 1e3, "GB/s", hbm))
    if want("geglu"):
        for N, D, cnt in [(4096, 640, 10), (1024, 1280, 60)]:
            nb = nbuf_for(B * N * 8 * D * 2 * 1.5)
            xs = [torch.randn(B, N, 8 * D, device=dev, dtype=DT) for _ in range(nb)]
            ys = [torch.empty(B, N, 4 * D, device=dev, dtype=DT) for _ in range(nb)]
            us = timeit(lambda i: ops.geglu(xs[i], out=ys[i]), nb, args.reps)
            out.append(("geglu", f"N{N}_F{4 * D} x{cnt}", us, 3 * B * N * 4 * D * 2 / us / 1e3, "GB/s", hbm))
    if want("attention"):
        for N, Nk, H, cnt in [(4096, 4096, 10, 10), (1024, 1024, 20, 60), (4096, 77, 10, 10), (1024, 77, 20, 60)]:
            if not keep(f"Nq{N}_Nk{Nk}_H{H}"):
                continue
            nb = 4
            if Nk == N:
                qkv = [torch.randn(B, N, 3 * H * 64, device=dev, dtype=DT) for _ in range(nb)]
                q = [t[..., :H * 64] for t in qkv]; k = [t[..., H * 64:2 * H * 64] for t in qkv]; v = [t[..., 2 * H * 64:] for t in qkv]
            else:
                q = [torch.randn(B, N, H * 64, device=dev, dtype=DT) for _ in range(nb)]
                kv = [torch.randn(B, Nk, 2 * H * 64, device=dev, dtype=DT) for _ in range(nb)]
                k = [t[..., :H * 64] for t in kv]; v = [t[..., H * 64:] for t in kv]
            o = [torch.empty(B, N, H * 64, device=dev, dtype=DT) for _ in range(nb)]
            us = timeit(lambda i: ops.attention(q[i], k[i], v[i], H, out=o[i]), nb, args.reps)
            fl = 4.0 * B * H * N * Nk * 64
            out.append(("attention", f"Nq{N}_Nk{Nk}_H{H} x{cnt}", us, fl / us / 1e6, "TFLOP/s", tf))
    if want("layout"):
        # k13 / k14 at the up-path sites of one forward, with the ATen kernels they replace under --compare
        import torch.nn.functional as F
        for Ca, Cb, hw, cnt in [(1280, 1280, 32, 2), (1280, 640, 32, 1), (1280, 640, 64, 1), (640, 640, 64, 1), (640, 320, 64, 1),
                                (640, 320, 128, 1), (320, 320, 128, 2)]:
            tag = f"cat_C{Ca}+{Cb}_HW{hw * hw}"
            if not keep(tag):
                continue
            nb = nbuf_for(2 * B * (Ca + Cb) * hw * hw * 2)
            mk = lambda c: torch.randn(B, c, hw, hw, device=dev, dtype=DT).contiguous(memory_format=torch.channels_last)
            as_, bs = [mk(Ca) for _ in range(nb)], [mk(Cb) for _ in range(nb)]
            ys = [torch.empty(B, Ca + Cb, hw, hw, device=dev, dtype=DT).contiguous(memory_format=torch.channels_last) for _ in range(nb)]
            nbytes = 2 * B * (Ca + Cb) * hw * hw * 2
            us = timeit(lambda i: ops.cat_channels(as_[i], bs[i], out=ys[i]), nb, args.reps)
            out.append(("layout", f"{tag} x{cnt}", us, nbytes / us / 1e3, "GB/s", hbm))
            if args.compare:
                us = timeit(lambda i: torch.cat([as_[i], bs[i]], dim=1, out=ys[i]), nb, args.reps)
                out.append(("cmp:aten", f"torch.cat {tag} x{cnt}", us, nbytes / us / 1e3, "GB/s", hbm))
        for Cc, hw, cnt in [(1280, 32, 1), (640, 64, 1)]:
            tag = f"upsample_C{Cc}_HW{hw * hw}"
            if not keep(tag):
                continue
            nb = nbuf_for(5 * B * Cc * hw * hw * 2)
            xs = [torch.randn(B, Cc, hw, hw, device=dev, dtype=DT).contiguous(memory_format=torch.channels_last) for _ in range(nb)]
            ys = [torch.empty(B, Cc, 2 * hw, 2 * hw, device=dev, dtype=DT).contiguous(memory_format=torch.channels_last) for _ in range(nb)]
            nbytes = 5 * B * Cc * hw * hw * 2
            us = timeit(lambda i: ops.upsample_nearest2x(xs[i], out=ys[i]), nb, args.reps)
            out.append(("layout", f"{tag} x{cnt}", us, nbytes / us / 1e3, "GB/s", hbm))
            if args.compare:
                us = timeit(lambda i: F.interpolate(xs[i], scale_factor=2.0, mode="nearest"), nb, args.reps)
                out.append(("cmp:aten", f"F.interpolate nearest {tag} x{cnt}", us, nbytes / us / 1e3, "GB/s", hbm))
    if want("blend"):
        for imgs in (1, 2048):
            if not keep(f"imgs{imgs}"):
                continue
            x = torch.randn(imgs, 4, 128, 128, device=dev)
            e = torch.randn(imgs, 4, 4, 128, 128, device=dev, dtype=DT)
            m = (torch.rand(3, 1, 128, 128, device=dev) > 0.5).float()
            o = torch.empty_like(x)
            us = timeit(lambda i: ops.tweedie_blend_ddim(x, e, m, 0.0438, 0.0518, 0.8, out=o), 1, args.reps)
            nbytes = imgs * (4 * 16384 * 8 + 16 * 16384 * 2) + 3 * 16384 * 4
            out.append(("blend", f"imgs{imgs}", us, nbytes / us / 1e3, "GB/s", hbm))
    if want("video"):
        # k11 / k12 of the I2VGen-XL stage: one 1280x720 16-frame video (4 x 16 x 90 x 160 latents) and a stack of 64 of them
        for vids in (1, 64):
            n = vids * 4 * 16 * 90 * 160
            x, vu, vc, o = (torch.randn(n, device=dev, dtype=DT) for _ in range(4))
            us = timeit(lambda i: ops.vpred_cfg_ddim(x, vu, vc, 0.5, 0.6, 9.0, out=o), 1, args.reps)
            out.append(("video", f"vpred_videos{vids}", us, 4 * n * 2 / us / 1e3, "GB/s", hbm))
        for C, hw in [(1280, 23 * 40), (1280, 45 * 80)]:                          # mid block / up block 1 feature maps of a 720p run
            y = torch.randn(32, C, hw, device=dev, dtype=DT)
            us = timeit(lambda i: ops.frame_inject(y, 2, 16, 0.7), 1, args.reps)
            out.append(("video", f"inject_C{C}_HW{hw}", us, 2 * 30 * C * hw * 2 / us / 1e3, "GB/s", hbm))
    if want("routed"):
        # k3: grouped tcgen05 GEMM (custom cross K/V, text-only: once per prompt set) and the rank-4 LoRA delta kernel
        for d, H in [(1280, 20), (640, 10)]:
            if not keep(f"kv_d{d}"):
                continue
            ehs = torch.randn(B, 77, 2048, device=dev, dtype=DT)
            ws = [torch.randn(2 * d, 2048, device=dev, dtype=DT) * 0.02 for _ in range(B)]
            y = torch.empty(B, 77, 2 * d, device=dev, dtype=DT)
            us = timeit(lambda i: ops.routed_linear(ehs, ws, out=y), 1, args.reps)
            out.append(("routed", f"gemm_kv_d{d}_M77_K2048_N{2 * d}", us, 2.0 * B * 77 * 2048 * 2 * d / us / 1e6, "TFLOP/s", tf))
        for N, d, cnt in [(1024, 1280, 60), (4096, 640, 10)]:
            for which, nseg, nout, per in [("qkv", 3, 3 * d, 1), ("q|out", 1, d, 3)]:
                tag = f"lora_{which}_N{N}_d{d}"
                if not keep(tag):
                    continue
                nb = nbuf_for(B * N * (d + 2 * nout) * 2)
                xs = [torch.randn(B, N, d, device=dev, dtype=DT) for _ in range(nb)]
                ys = [torch.randn(B, N, nout, device=dev, dtype=DT) for _ in range(nb)]
                downs = [None] + [torch.randn(nseg * 4, d, device=dev, dtype=DT) * 0.25 for _ in range(B - 1)]
                ups = [None] + [torch.randn(nout, 4, device=dev, dtype=DT) * 0.01 for _ in range(B - 1)]
                us = timeit(lambda i: ops.routed_linear(xs[i], None, downs, ups, nseg=nseg, out=ys[i]), nb, args.reps)
                nbytes = (B - 1) * N * (d + 2 * nout) * 2           # routed rows: read x, read + write y
                out.append(("routed", f"{tag} x{cnt * per}", us, nbytes / us / 1e3, "GB/s", hbm))
    if want("linear"):
        # k10 at the GEMM sites of one fused K=3 step (M = 4 rows x tokens): tile width auto / forced, and — with --compare —
        # the library path it replaces (cuBLAS GEMM + the stand-alone epilogue kernel)
        import torch.nn.functional as F
        from tweediemix_b200 import _lib
        lib = _lib.load()
        m32, m64 = B * 1024, B * 4096          # tokens of the 32x32 / 64x64 feature maps over the B rows this GPU holds
        sites = [("qkv", m32, 3840, 1280, "", 60), ("out|q", m32, 1280, 1280, "res", 180), ("ff1", m32, 10240, 1280, "geglu", 60),
                 ("ff2", m32, 1280, 5120, "res", 60), ("qkv", m64, 1920, 640, "", 10), ("out|q", m64, 640, 640, "res", 30),
                 ("ff1", m64, 5120, 640, "geglu", 10), ("ff2", m64, 640, 2560, "res", 10)]
        for name, M, N, K, epi, cnt in sites:
            tag = f"{name}_M{M}_N{N}_K{K}"
            if not keep(tag):
                continue
            n_out = N // 2 if epi == "geglu" else N
            nb = nbuf_for((M * K + M * n_out * 2) * 2)
            xs = [torch.randn(M, K, device=dev, dtype=DT) for _ in range(nb)]
            w = torch.randn(N, K, device=dev, dtype=DT) * K ** -0.5
            bias = torch.randn(N, device=dev)
            bias16 = bias.to(DT)
            rs = [torch.randn(M, N, device=dev, dtype=DT) for _ in range(nb)] if epi == "res" else None
            ys = [torch.empty(M, n_out, device=dev, dtype=DT) for _ in range(nb)]
            fl = 2.0 * M * N * K
            for bn in (0, 128, 192, 256, 320, 1000):         # + 1000 = split-K tail disabled (+ 2000 = split-K tail at any K)
                assert lib.tmx_linear_set_variant(bn) == 0
                us = timeit(lambda i: ops.linear(xs[i], w, bias, residual=rs[i] if rs else None, geglu=epi == "geglu", out=ys[i]), nb, args.reps)
                out.append(("linear" if bn == 0 else f"linear:v{bn}", f"{tag}_{epi or 'bias'} x{cnt if bn == 0 else 0}", us, fl / us / 1e6, "TFLOP/s", tf))
            lib.tmx_linear_set_variant(0)
            if args.compare:
                if epi == "geglu":
                    fn = lambda i: ops.geglu(F.linear(xs[i], w, bias16), out=ys[i])
                elif epi == "res":
                    fn = lambda i: ops.residual_add(F.linear(xs[i], w, bias16), rs[i], out=ys[i])
                else:
                    fn = lambda i: F.linear(xs[i], w, bias16)
                us = timeit(fn, nb, args.reps)
                out.append(("cmp:cublas", f"cuBLAS + tmx epilogue kernel {tag}_{epi or 'bias'} x{cnt}", us, fl / us / 1e6, "TFLOP/s", tf))
                us = timeit(lambda i: F.linear(xs[i], w, bias16), nb, args.reps)
                out.append(("cmp:cublas-gemm", f"cuBLAS GEMM only {tag} x{cnt}", us, fl / us / 1e6, "TFLOP/s", tf))
    if args.compare:
        import torch.nn.functional as F
        for N, H, cnt in [(4096, 10, 10), (1024, 20, 60)]:
            qkv = [torch.randn(B, N, 3, H, 64, device=dev, dtype=DT) for _ in range(4)]
            fl = 4.0 * B * H * N * N * 64
            hv = [[t[:, :, j].permute(0, 2, 1, 3) for j in range(3)] for t in qkv]          # [B, H, N, 64] views, like the product reads them
            try:
                us = timeit(lambda i: F.scaled_dot_product_attention(hv[i][0], hv[i][1], hv[i][2]), 4, args.reps)
                out.append(("cmp:sdpa", f"torch F.sdpa Nq{N}_H{H} x{cnt}", us, fl / us / 1e6, "TFLOP/s", tf))
            except Exception as e:
                print("cmp:sdpa failed:", str(e)[:200])
            for backend in ("CUDNN_ATTENTION", "FLASH_ATTENTION"):
                try:
                    from torch.nn.attention import SDPBackend, sdpa_kernel
                    with sdpa_kernel(getattr(SDPBackend, backend)):
                        us = timeit(lambda i: F.scaled_dot_product_attention(hv[i][0], hv[i][1], hv[i][2]), 4, args.reps)
                    out.append((f"cmp:sdpa", f"torch {backend} Nq{N}_H{H} x{cnt}", us, fl / us / 1e6, "TFLOP/s", tf))
                except Exception as e:
                    print(f"cmp:sdpa {backend} failed:", str(e)[:200])
            try:
                from flash_attn import flash_attn_func
                us = timeit(lambda i: flash_attn_func(qkv[i][:, :, 0], qkv[i][:, :, 1], qkv[i][:, :, 2]), 4, args.reps)
                out.append(("cmp:flash", f"flash-attn Nq{N}_H{H} x{cnt}", us, fl / us / 1e6, "TFLOP/s", tf))
            except Exception as e:
                print("cmp:flash failed:", str(e)[:200])
        for C, HW, cnt in [(320, 16384, 5), (640, 4096, 9), (1280, 1024, 20)]:
            h = int(HW ** 0.5)
            nb = nbuf_for(B * C * HW * 4)
            xs = [torch.randn(B, C, h, h, device=dev, dtype=DT).contiguous(memory_format=torch.channels_last) for _ in range(nb)]
            g, b_ = torch.ones(C, device=dev, dtype=DT), torch.zeros(C, device=dev, dtype=DT)
            us = timeit(lambda i: F.silu(F.group_norm(xs[i], 32, g, b_, 1e-5)), nb, args.reps)
            out.append(("cmp:gn", f"torch group_norm+silu C{C}_HW{HW} x{cnt}", us, 2 * B * C * HW * 2 / us / 1e3, "GB/s", hbm))
        for N, D, cnt in [(4096, 640, 30), (1024, 1280, 180)]:
            nb = nbuf_for(B * N * D * 4)
            xs = [torch.randn(B, N, D, device=dev, dtype=DT) for _ in range(nb)]
            g, b_ = torch.ones(D, device=dev, dtype=DT), torch.zeros(D, device=dev, dtype=DT)
            us = timeit(lambda i: F.layer_norm(xs[i], (D,), g, b_, 1e-5), nb, args.reps)
            out.append(("cmp:ln", f"torch layer_norm N{N}_D{D} x{cnt}", us, 2 * B * N * D * 2 / us / 1e3, "GB/s", hbm))
    tot = {}
    for fam, tag, us, rate, unit, peak in out:
        print(f"{fam:10s} {tag:28s} {us:9.2f} us  {rate:9.1f} {unit:8s} {100 * rate / peak:5.1f}% of measured peak")
        cnt = int(tag.split(" x")[1]) if " x" in tag else 0
        tot[fam] = tot.get(fam, 0.0) + us * cnt
    print("per fused K=3 step (us, kernels back to back):", {k: round(v, 1) for k, v in tot.items() if v})


if __name__ == "__main__":
    main()
