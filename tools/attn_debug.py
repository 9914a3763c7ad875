#!/usr/bin/env python
"""GPU-side diagnosis of tmx_attn_fwd with structured inputs; writes gpurun_out/attn_debug.txt.
Each case isolates one piece of the kernel so a wrong result says where to look:
  ones_v    : V = 1          -> O must be 1 everywhere (softmax normalisation, PV accumulate flag)
  zero_q    : Q = 0          -> O = mean_j V[j]       (V tile layout / MN-major descriptor / P in TMEM)
  onehot_k  : q.k picks j*   -> O = V[j*]             (Q/K descriptors, K-step advance, swizzle)
  random    : randn          -> vs fp32 softmax reference
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tweediemix_b200 import _lib, ops  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(OUT, exist_ok=True)
log = open(os.path.join(OUT, "attn_debug.txt"), "w")


def say(*a):
    msg = " ".join(str(x) for x in a)
    print(msg)
    log.write(msg + "\n")
    log.flush()


def ref(q, k, v, H, scale):
    B, Nq, HD = q.shape
    D = HD // H
    qf = q.float().view(B, Nq, H, D).permute(0, 2, 1, 3)
    kf = k.float().view(B, -1, H, D).permute(0, 2, 1, 3)
    vf = v.float().view(B, -1, H, D).permute(0, 2, 1, 3)
    p = torch.softmax(qf @ kf.transpose(-1, -2) * scale, dim=-1)
    return (p @ vf).permute(0, 2, 1, 3).reshape(B, Nq, HD)


def report(name, got, want):
    d = (got.float() - want).abs()
    bad = d > 0.02
    say(f"  {name}: max|diff|={d.max().item():.4e} mean={d.mean().item():.3e} bad={int(bad.sum())}/{d.numel()} "
        f"finite={bool(torch.isfinite(got.float()).all())}")
    if bad.any():
        idx = bad.nonzero()
        rows = sorted(set(idx[:, 1].tolist()))
        cols = sorted(set(idx[:, 2].tolist()))
        say(f"    bad rows (first 16): {rows[:16]} ... count {len(rows)}; bad cols (first 16): {cols[:16]} ... count {len(cols)}")
        i = idx[0].tolist()
        say(f"    first bad at {i}: got {got[tuple(i)].item():.5f} want {want[tuple(i)].item():.5f}")
        say(f"    got[0,0,:8]  = {[round(x, 4) for x in got[0, 0, :8].float().tolist()]}")
        say(f"    want[0,0,:8] = {[round(x, 4) for x in want[0, 0, :8].tolist()]}")
    return float(d.max())


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda"
    lib = _lib.load()
    worst = 0.0
    for dtype in (torch.bfloat16, torch.float16):
        for nq in (1, 2):
            lib.tmx_attn_set_variant(nq)
            for (B, H, Nq, Nk) in [(1, 1, 128, 128), (1, 1, 128, 256), (1, 2, 256, 384), (2, 3, 200, 77), (1, 1, 1, 1), (2, 2, 1024, 1024)]:
                say(f"== dtype={dtype} NQ={nq} B={B} H={H} Nq={Nq} Nk={Nk}")
                g = torch.Generator(device="cpu").manual_seed(Nq * 7 + Nk)
                D = 64
                q = torch.randn(B, Nq, H * D, generator=g).to(dtype).to(dev)
                k = torch.randn(B, Nk, H * D, generator=g).to(dtype).to(dev)
                v = torch.randn(B, Nk, H * D, generator=g).to(dtype).to(dev)
                sc = D ** -0.5
                try:
                    o = ops.attention(q, k, torch.ones_like(v), H)
                    torch.cuda.synchronize()
                    worst = max(worst, report("ones_v ", o, torch.ones_like(o, dtype=torch.float32)))
                    o = ops.attention(torch.zeros_like(q), k, v, H)
                    torch.cuda.synchronize()
                    worst = max(worst, report("zero_q ", o, ref(torch.zeros_like(q), k, v, H, sc)))
                    # one-hot: q_i = 8 * e_{i % 64}, k_j = 8 * e_{j % 64} -> row i attends to all j with j%64 == i%64
                    qi = torch.zeros(B, Nq, H, D, device=dev)
                    qi[:, torch.arange(Nq), :, torch.arange(Nq) % D] = 8.0
                    ki = torch.zeros(B, Nk, H, D, device=dev)
                    ki[:, torch.arange(Nk), :, torch.arange(Nk) % D] = 8.0
                    qi, ki = qi.reshape(B, Nq, H * D).to(dtype), ki.reshape(B, Nk, H * D).to(dtype)
                    o = ops.attention(qi, ki, v, H)
                    torch.cuda.synchronize()
                    worst = max(worst, report("onehot ", o, ref(qi, ki, v, H, sc)))
                    o = ops.attention(q, k, v, H)
                    torch.cuda.synchronize()
                    worst = max(worst, report("random ", o, ref(q, k, v, H, sc)))
                except Exception as e:  # noqa: BLE001
                    say("  EXCEPTION:", repr(e))
                    say("RESULT: FAIL (exception)")
                    return 1
    lib.tmx_attn_set_variant(0)
    say(f"RESULT: worst max|diff| = {worst:.4e} ->", "PASS" if worst < 0.02 else "FAIL")
    return 0 if worst < 0.02 else 1


if __name__ == "__main__":
    sys.exit(main())
