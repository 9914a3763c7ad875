"""Torch-tensor front-ends of the libtmx C ABI (device pointers + sizes go straight through ctypes).

PyTorch is plumbing here — it owns the device memory and the stream; all arithmetic of these ops
is in the hand-written sm_100a kernels of ``csrc/``.  Nothing in this module falls back to a
PyTorch implementation: non-CUDA tensors, unsupported shapes or a missing library raise.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib

_DT = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise RuntimeError(f"tmx: unsupported dtype {t.dtype}") from None


def _dev(*ts: torch.Tensor) -> int:
    d = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("tmx ops run on CUDA tensors only (no CPU fallback)")
        if d is None:
            d = t.device.index
        elif t.device.index != d:
            raise RuntimeError("tmx: tensors on different devices")
    _lib.init(d)
    return d


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


# ---- bookkeeping: how many tmx kernels were launched, and (optionally) how long each took ---------
LAUNCHES = {}          # kernel family -> launches issued through this module (graph replays are added by the sampler)
_profile = None        # a KernelProfile while bench.py's instrumented step runs


def launch_count() -> int:
    return sum(LAUNCHES.values())


def add_launches(by_family: dict, times: int = 1) -> None:
    for k, v in by_family.items():
        LAUNCHES[k] = LAUNCHES.get(k, 0) + v * times


class KernelProfile:
    """CUDA-event timing of every tmx launch on the current stream (eager mode only).  ``work`` is the
    algorithmic FLOPs (attention) or bytes (bandwidth kernels) of the launch, as defined in tmx.h."""

    def __init__(self):
        self.records = []          # (family, tag, work, start_event, end_event)

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for fam, tag, work, e0, e1 in self.records:
            d = out.setdefault((fam, tag), {"launches": 0, "ms": 0.0, "work": 0.0})
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["work"] += work
        return out


def set_profile(p) -> None:
    global _profile
    _profile = p


class _Launch:
    """with _Launch(family, n_kernels, tag, work): <ctypes call>"""
    __slots__ = ("fam", "n", "tag", "work", "e0")

    def __init__(self, fam, n=1, tag="", work=0.0):
        self.fam, self.n, self.tag, self.work = fam, n, tag, work

    def __enter__(self):
        LAUNCHES[self.fam] = LAUNCHES.get(self.fam, 0) + self.n
        if _profile is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _profile is not None and exc[0] is None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _profile.records.append((self.fam, self.tag, self.work, self.e0, e1))
        return False


# ------------------------------------------------------------------------------------------ k7

def tweedie_blend_ddim(x, eps, masks, a_t: float, a_next: float, g: float, *, is_last=False,
                       weights=None, out=None, x0_out=None, ref_rounding=False):
    """Fused CFG + Tweedie x0 + masked blend + DDIM (fusion_sampling.py:376-386,430,471-472).

    x [imgs,C,H,W] fp32; eps [imgs*(K+1),C,H,W] or [imgs,K+1,C,H,W]; masks [K,1,H,W] fp32 or None.
    Returns x_next (fp32, same shape as x)."""
    _dev(x, eps, masks, out, x0_out)
    assert x.dtype == torch.float32 and x.is_contiguous() and eps.is_contiguous()
    imgs, Cc = x.shape[0], x.shape[1]
    HW = x.shape[2] * x.shape[3]
    rows = eps.numel() // (imgs * Cc * HW)
    K = rows - 1
    if masks is not None:
        assert masks.dtype == torch.float32 and masks.is_contiguous() and masks.numel() == K * HW, "masks must be [K,1,H,W] fp32"
    if out is None:
        out = torch.empty_like(x)
    w = None
    if weights is not None:
        assert len(weights) == K
        w = (C.c_float * K)(*[float(v) for v in weights])
    nbytes = imgs * (Cc * HW * (8 + (4 if x0_out is not None else 0)) + (K + 1) * Cc * HW * eps.element_size()
                     + (K * HW * 4 if masks is not None else 0))
    with _Launch("blend", 1, f"imgs{imgs}_K{K}", nbytes):
        rc = _lib.load().tmx_tweedie_blend_ddim_fwd(
            _p(x), _p(eps), _p(masks), w, _p(out), _p(x0_out), imgs, K, Cc, HW,
            float(a_t), float(a_next), float(g), int(bool(is_last)), _dt(eps),
            _lib.ROUND_REF if ref_rounding else _lib.ROUND_FP32, _stream())
    _lib.check(rc, "tmx_tweedie_blend_ddim_fwd")
    return out


def _wptr(weights, K):
    if weights is None:
        return None
    assert len(weights) == K
    return (C.c_float * K)(*[float(v) for v in weights])


def blend_partial(eps_rows, masks, row_ids, acc, imgs: int = 1, *, K: int | None = None, weights=None):
    """acc[img,0] = sum_c w_c m_c eps_c over the concept rows this rank owns; acc[img,1] = eps_u if the
    uncond row (id 0) is owned else 0.  ``eps_rows`` [imgs, R, C, H, W] (or None when R == 0);
    ``masks`` [K,1,H,W] fp32 or None (= ones); ``weights`` K host floats or None (= ones)."""
    _dev(eps_rows, masks, acc)
    R = len(row_ids)
    assert acc.dtype == torch.float32 and acc.is_contiguous() and acc.dim() == 5 and acc.shape[0] == imgs and acc.shape[1] == 2
    Cc, HW = acc.shape[2], acc.shape[3] * acc.shape[4]
    if K is None:
        assert masks is not None, "K is required when masks is None"
        K = masks.shape[0]
    if masks is not None:
        assert masks.dtype == torch.float32 and masks.is_contiguous() and masks.numel() == K * HW
    if R:
        assert eps_rows.is_contiguous() and eps_rows.numel() == imgs * R * Cc * HW
    ids = (C.c_int * max(R, 1))(*[int(r) for r in row_ids])
    with _Launch("blend_partial"):
        rc = _lib.load().tmx_blend_partial_fwd(_p(eps_rows) if R else None, _p(masks), _wptr(weights, K), ids, _p(acc),
                                               imgs, R, K, Cc, HW, _dt(eps_rows) if R else _lib.F32, _stream())
    _lib.check(rc, "tmx_blend_partial_fwd")
    return acc


def blend_finish(x, acc, masks, a_t, a_next, g, *, is_last=False, out=None, x0_out=None, K: int | None = None, weights=None):
    """Every rank: finish from the all-reduced ``acc`` (bit-identical on all ranks)."""
    _dev(x, acc, masks, out, x0_out)
    imgs, Cc, HW = x.shape[0], x.shape[1], x.shape[2] * x.shape[3]
    if K is None:
        assert masks is not None, "K is required when masks is None"
        K = masks.shape[0]
    if out is None:
        out = torch.empty_like(x)
    with _Launch("blend_finish"):
        rc = _lib.load().tmx_blend_finish_fwd(_p(x), _p(acc), _p(masks), _wptr(weights, K), _p(out), _p(x0_out), imgs, K, Cc, HW,
                                              float(a_t), float(a_next), float(g), int(bool(is_last)), _stream())
    _lib.check(rc, "tmx_blend_finish_fwd")
    return out


# ------------------------------------------------------------------------------------------ k4/k5

_gn_ws = {}
_graph_pinned = False      # a CUDA graph has captured raw pointers into tmx-owned buffers
_retired = []              # buffers superseded / evicted after that: kept alive, replays still touch them


def pin_graph_resources() -> None:
    """Called by the sampler after every CUDA-graph capture: from now on the GroupNorm workspace and the cached
    cross-attention K/V tensors a graph may point at are never freed (a superseded workspace or an evicted cache entry is
    retired into a keep-alive list instead), so a replay can never read or write recycled memory."""
    global _graph_pinned
    _graph_pinned = True


def retire(obj) -> None:
    if _graph_pinned:
        _retired.append(obj)


def _workspace(dev: int, nbytes: int) -> torch.Tensor:
    ws = _gn_ws.get(dev)
    if ws is None or ws.numel() < nbytes:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("tmx: GroupNorm workspace would have to grow during CUDA-graph capture; run the op once eagerly first")
        if ws is not None:
            retire(ws)
        ws = torch.zeros(max(nbytes, 32 << 20), dtype=torch.uint8, device=f"cuda:{dev}")   # tickets must start at zero
        _gn_ws[dev] = ws
    return ws


def group_norm(x, gamma, beta, groups: int, eps: float, *, silu=False, add=None, out=None, x2=None):
    """GroupNorm(+add[n,c] before the norm)(+SiLU).  ``x`` is a 4-D NCHW-shaped tensor that is either
    contiguous (NCHW memory) or channels_last (NHWC memory); the output has the same memory format.
    ``x2`` (channels_last 16-bit, same N/H/W): normalise the channel concatenation ``cat([x, x2], 1)`` without materialising it."""
    d = _dev(x, gamma, beta, add, out, x2)
    N, Cc, H, W = x.shape
    if x.is_contiguous() and x2 is None:
        layout = _lib.NCHW
    elif x.is_contiguous(memory_format=torch.channels_last):
        layout = _lib.NHWC
    else:
        raise RuntimeError("tmx.group_norm: x must be contiguous or channels_last")
    C1 = Cc
    if x2 is not None:
        if not (x2.is_contiguous(memory_format=torch.channels_last) and x2.dtype == x.dtype and x2.shape[0] == N and tuple(x2.shape[2:]) == (H, W)):
            raise RuntimeError("tmx.group_norm: x2 must be channels_last with x's dtype, batch and spatial size")
        if x.dtype not in (torch.float16, torch.bfloat16) or Cc % 8 != 0:
            raise RuntimeError("tmx.group_norm: the two-source form needs 16-bit inputs and C1 % 8 == 0")
        Cc = C1 + x2.shape[1]
    if out is None:
        out = torch.empty_like(x) if x2 is None else torch.empty((N, Cc, H, W), dtype=x.dtype, device=x.device).contiguous(memory_format=torch.channels_last)
    assert out.dtype == x.dtype and (out.stride() == x.stride() if x2 is None else (out.shape[1] == Cc and out.is_contiguous(memory_format=torch.channels_last)))
    assert gamma.dtype == torch.float32 and beta.dtype == torch.float32 and gamma.numel() == Cc, "gamma/beta must be fp32 [C]"
    if add is not None:
        assert add.dtype == torch.float32 and add.is_contiguous() and add.shape == (N, Cc)
    lib = _lib.load()
    ws = _workspace(d, lib.tmx_groupnorm_workspace_bytes(N, Cc, H * W, groups, layout))
    nbytes = 2.0 * N * Cc * H * W * x.element_size()
    with _Launch("groupnorm", lib.tmx_groupnorm_launches(N, Cc, H * W, layout, _dt(x)), f"C{Cc}_HW{H * W}" + ("_cat" if x2 is not None else ""), nbytes):
        if x2 is None:
            rc = lib.tmx_groupnorm_fwd(_p(x), _p(gamma), _p(beta), _p(add), _p(out), _p(ws), N, Cc, H * W, groups,
                                       float(eps), _lib.ACT_SILU if silu else _lib.ACT_NONE, layout, _dt(x), _stream())
        else:
            rc = lib.tmx_groupnorm_cat_fwd(_p(x), _p(x2), C1, _p(gamma), _p(beta), _p(add), _p(out), _p(ws), N, Cc, H * W, groups,
                                           float(eps), _lib.ACT_SILU if silu else _lib.ACT_NONE, _dt(x), _stream())
    _lib.check(rc, "tmx_groupnorm_fwd")
    return out


def cat_free_supported(a, b) -> bool:
    """Can ``group_norm(a, x2=b)`` stand in for a GroupNorm over ``cat([a, b], 1)``?  (16-bit channels_last CUDA tensors, C % 8 == 0)"""
    return (a.is_cuda and b.is_cuda and a.dtype == b.dtype and a.dtype in (torch.float16, torch.bfloat16)
            and a.shape[1] % 8 == 0 and b.shape[1] % 8 == 0
            and a.is_contiguous(memory_format=torch.channels_last) and b.is_contiguous(memory_format=torch.channels_last))


# ------------------------------------------------------------------------------------------ k6

def residual_add(a, b, inv_scale: float = 1.0, out=None):
    """(a + b) * inv_scale; a and b must share dtype and memory layout (dense)."""
    _dev(a, b, out)
    assert a.dtype == b.dtype and a.shape == b.shape and a.stride() == b.stride()
    if out is None:
        out = torch.empty_like(a)
    assert out.stride() == a.stride()
    with _Launch("resadd", 1, "", 3.0 * a.numel() * a.element_size()):
        rc = _lib.load().tmx_resadd_fwd(_p(a), _p(b), _p(out), a.numel(), float(inv_scale), _dt(a), _stream())
    _lib.check(rc, "tmx_resadd_fwd")
    return out


def layout_supported(*ts) -> bool:
    """4-D channels_last CUDA fp16 / bf16 tensors of one dtype whose channel counts are multiples of 8 (k13 / k14)."""
    return all(t.is_cuda and t.dim() == 4 and t.dtype in (torch.float16, torch.bfloat16) and t.dtype == ts[0].dtype and t.shape[1] % 8 == 0
               and t.is_contiguous(memory_format=torch.channels_last) for t in ts)


def cat_channels(a, b, out=None):
    """torch.cat([a, b], dim=1) of two channels_last 16-bit feature maps ([D] up-block `torch.cat([hidden_states, res_hidden_states], 1)`)."""
    _dev(a, b, out)
    if not layout_supported(a, b) or a.shape[0] != b.shape[0] or a.shape[2:] != b.shape[2:]:
        raise RuntimeError("tmx.cat_channels: inputs must be channels_last fp16 / bf16 [N, C, H, W] with C % 8 == 0 and equal N, H, W")
    N, Ca, H, W = a.shape
    Cb = b.shape[1]
    if out is None:
        out = torch.empty((N, Ca + Cb, H, W), dtype=a.dtype, device=a.device, memory_format=torch.channels_last)
    assert out.shape == (N, Ca + Cb, H, W) and out.dtype == a.dtype and out.is_contiguous(memory_format=torch.channels_last)
    with _Launch("cat", 1, f"C{Ca}+{Cb}_HW{H * W}", 2.0 * N * H * W * (Ca + Cb) * a.element_size()):
        rc = _lib.load().tmx_cat_channels_fwd(_p(a), _p(b), _p(out), N * H * W, Ca, Cb, _dt(a), _stream())
    _lib.check(rc, "tmx_cat_channels_fwd")
    return out


def upsample_nearest2x(x, out=None):
    """F.interpolate(x, scale_factor=2.0, mode="nearest") of a channels_last 16-bit feature map ([D] Upsample2D)."""
    _dev(x, out)
    if not layout_supported(x):
        raise RuntimeError("tmx.upsample_nearest2x: input must be channels_last fp16 / bf16 [N, C, H, W] with C % 8 == 0")
    N, Cc, H, W = x.shape
    if out is None:
        out = torch.empty((N, Cc, 2 * H, 2 * W), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    assert out.shape == (N, Cc, 2 * H, 2 * W) and out.dtype == x.dtype and out.is_contiguous(memory_format=torch.channels_last)
    with _Launch("upsample", 1, f"C{Cc}_HW{H * W}", 5.0 * x.numel() * x.element_size()):
        rc = _lib.load().tmx_upsample_nearest2x_fwd(_p(x), _p(out), N, H, W, Cc, _dt(x), _stream())
    _lib.check(rc, "tmx_upsample_nearest2x_fwd")
    return out


def bias_residual_add(a, bias, b=None, inv_scale: float = 1.0, out=None):
    """(a + bias[c] (+ b)) * inv_scale on channels_last 4-D (or [..., C] dense) 16-bit tensors; bias fp32 [C]."""
    _dev(a, b, bias, out)
    if a.dim() == 4:
        if not a.is_contiguous(memory_format=torch.channels_last):
            raise RuntimeError("tmx.bias_residual_add: 4-D input must be channels_last")
        Cc = a.shape[1]
    else:
        assert a.is_contiguous()
        Cc = a.shape[-1]
    assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == Cc
    if b is not None:
        assert b.dtype == a.dtype and b.shape == a.shape and b.stride() == a.stride()
    if out is None:
        out = torch.empty_like(a)
    assert out.stride() == a.stride()
    rows = a.numel() // Cc
    with _Launch("resadd", 1, "bias", (3.0 if b is not None else 2.0) * a.numel() * a.element_size()):
        rc = _lib.load().tmx_bias_resadd_fwd(_p(a), _p(b), _p(bias), _p(out), rows, Cc, float(inv_scale), _dt(a), _stream())
    _lib.check(rc, "tmx_bias_resadd_fwd")
    return out


def residual_add_layer_norm(a, b, gamma, beta, eps: float, h_out=None, n_out=None):
    """h = a + b (stored, may alias a or b); n = LayerNorm(h) over the last dim.  Returns (h, n)."""
    _dev(a, b, gamma, beta, h_out, n_out)
    assert a.is_contiguous() and b.is_contiguous() and a.shape == b.shape and a.dtype == b.dtype
    assert gamma.dtype == torch.float32 and beta.dtype == torch.float32
    D = a.shape[-1]
    rows = a.numel() // D
    if h_out is None:
        h_out = torch.empty_like(a)
    if n_out is None:
        n_out = torch.empty_like(a)
    with _Launch("resadd_ln", 1, f"D{D}", 4.0 * a.numel() * a.element_size()):
        rc = _lib.load().tmx_resadd_layernorm_fwd(_p(a), _p(b), _p(gamma), _p(beta), _p(h_out), _p(n_out), rows, D, float(eps), _dt(a), _stream())
    _lib.check(rc, "tmx_resadd_layernorm_fwd")
    return h_out, n_out


def layer_norm(x, gamma, beta, eps: float, out=None):
    """LayerNorm over the last dim of a contiguous fp16 / bf16 tensor; gamma / beta fp32 [D]."""
    _dev(x, gamma, beta, out)
    assert x.is_contiguous() and gamma.dtype == torch.float32 and beta.dtype == torch.float32
    D = x.shape[-1]
    rows = x.numel() // D
    if out is None:
        out = torch.empty_like(x)
    with _Launch("layernorm", 1, f"D{D}", 2.0 * x.numel() * x.element_size()):
        rc = _lib.load().tmx_layernorm_fwd(_p(x), _p(gamma), _p(beta), _p(out), rows, D, float(eps), _dt(x), _stream())
    _lib.check(rc, "tmx_layernorm_fwd")
    return out


def geglu(x, out=None):
    """x [..., 2F] -> x[..., :F] * gelu(x[..., F:]) (exact erf GELU); fp16 / bf16."""
    _dev(x, out)
    assert x.is_contiguous()
    F2 = x.shape[-1]
    assert F2 % 2 == 0
    rows = x.numel() // F2
    if out is None:
        out = torch.empty(*x.shape[:-1], F2 // 2, dtype=x.dtype, device=x.device)
    with _Launch("geglu", 1, "", 1.5 * x.numel() * x.element_size()):
        rc = _lib.load().tmx_geglu_fwd(_p(x), _p(out), rows, F2 // 2, _dt(x), _stream())
    _lib.check(rc, "tmx_geglu_fwd")
    return out


# ------------------------------------------------------------------------------------------ k3

MAX_ROUTED_ROWS = 16


def routed_linear(x, weights=None, lora_down=None, lora_up=None, *, nseg: int = 1, out=None):
    """Per-row routed projection (``utils_custom.py:64-82``, ``utils_lora.py:65-79,113-119``).

    ``x`` [B, M, Kin].  ``weights``: list of B tensors [Nout, Kin] -> ``out[b] = x[b] @ weights[b].T`` (grouped tcgen05
    GEMM); ``None`` -> ``out`` must already hold the shared-weight projection.  ``lora_down`` / ``lora_up``: lists of B
    entries, each a tensor ([nseg*r, Kin] / [Nout, r]) or ``None`` (row not routed) -> the rank-r deltas are ADDED to
    ``out``; output column n uses segment ``n // (Nout // nseg)`` (packed q|k|v: nseg = 3)."""
    _dev(x, out)
    assert x.dim() == 3 and x.is_contiguous(), "x must be a contiguous [B, M, Kin] tensor"
    B, M, Kin = x.shape
    if weights is not None:
        assert len(weights) == B
        Nout = weights[0].shape[0]
        for w in weights:
            assert w.dtype == x.dtype and w.is_cuda and w.is_contiguous() and tuple(w.shape) == (Nout, Kin)
        if out is None:
            out = torch.empty((B, M, Nout), dtype=x.dtype, device=x.device)
    else:
        assert out is not None and lora_down is not None, "without weights the call accumulates LoRA deltas into `out`"
        Nout = out.shape[-1]
    assert out.is_contiguous() and tuple(out.shape) == (B, M, Nout) and out.dtype == x.dtype
    rank = 0
    if lora_down is not None:
        assert lora_up is not None and len(lora_down) == B and len(lora_up) == B
        for d, u in zip(lora_down, lora_up):
            assert (d is None) == (u is None)
            if d is not None:
                rank = d.shape[0] // nseg
                assert d.dtype == x.dtype and u.dtype == x.dtype and d.is_contiguous() and u.is_contiguous()
                assert tuple(d.shape) == (nseg * rank, Kin) and tuple(u.shape) == (Nout, rank), (d.shape, u.shape, nseg, Nout)
    launches = (1 if weights is not None else 0) + (1 if rank else 0)
    for b0 in range(0, B, MAX_ROUTED_ROWS):                    # the ABI takes at most 16 batch rows per call
        nb = min(MAX_ROUTED_ROWS, B - b0)
        arr = lambda ts: (C.c_void_p * nb)(*[None if t is None else t.data_ptr() for t in ts[b0:b0 + nb]]) if ts is not None else None
        wp, dp, up = arr(weights), arr(lora_down), arr(lora_up)
        if weights is None and all(t is None for t in lora_down[b0:b0 + nb]):
            continue
        with _Launch("routed_linear", launches, f"M{M}_K{Kin}_N{Nout}", 2.0 * nb * M * Kin * Nout if weights is not None else 0.0):
            rc = _lib.load().tmx_routed_linear_fwd(_p(x[b0:]), wp, dp, up, _p(out[b0:]), nb, M, Kin, Nout, int(rank), int(nseg), _dt(x), _stream())
        _lib.check(rc, "tmx_routed_linear_fwd")
    return out


# ------------------------------------------------------------------------------------------ k10

# Which GEMMs of the transformer blocks run in the persistent tcgen05 kernel (k10) and which go to the library:
#   'auto' (default) — k10 where it is at least as fast as cuBLAS + the stand-alone epilogue kernel on this box
#                      (profiles/r02j_kbench_linear.txt): the GEGLU projection (value * gelu(gate) in the epilogue: 79.5 vs
#                      102.9 us at d = 1280, 111 vs 131 us at d = 640) and every projection that carries a LoRA tail (the
#                      rank-r delta rides as one extra MMA step instead of a separate read-modify-write pass); the plain
#                      projections, where cuBLAS' tile shapes fill the 148 SMs better (out|q 18.4 vs 21.5 us, ff2 43 vs 52 us),
#                      stay on cuBLAS with the fused residual-add + LayerNorm kernel behind them;
#   'tmx'            — every projection / feed-forward GEMM in k10 (residual adds in its epilogue, one-pass LayerNorm after);
#   'cublas'         — library GEMMs + stand-alone GEGLU / LoRA-delta / add+LayerNorm kernels (A/B measurements).
GEMM_IMPL = os.environ.get("TMX_GEMM", "auto")


def gemm_in_k10(kind: str) -> bool:
    """Policy: does a GEMM of this kind ('plain', 'geglu', 'lora') run in k10?"""
    if GEMM_IMPL == "tmx":
        return True
    return GEMM_IMPL == "auto" and kind in ("geglu", "lora")


_lin_ws = {}


def _linear_workspace(dev: int) -> torch.Tensor:
    """Split-K tail workspace of k10 (counters + fp32 partial tiles), one per device, allocated on first use; the kernels
    leave the counters at zero, launches on one stream are serialised, so one buffer serves every call."""
    ws = _lin_ws.get(dev)
    if ws is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("tmx: the GEMM workspace must be allocated before CUDA-graph capture; run the op once eagerly first")
        ws = _lin_ws[dev] = torch.zeros(_lib.load().tmx_linear_workspace_bytes(), dtype=torch.uint8, device=f"cuda:{dev}")
    return ws


def linear_supported(x, w) -> bool:
    """Shapes on the kernel's tile grid (K % 64 == 0, N % 8 == 0, 16-bit CUDA tensors, uniform row stride)."""
    return (GEMM_IMPL != "cublas" and x.is_cuda and x.dtype in (torch.float16, torch.bfloat16) and w.dtype == x.dtype
            and x.shape[-1] % 64 == 0 and w.shape[0] % 8 == 0 and x.is_contiguous() and w.is_contiguous())


def geglu_interleave_index(F: int, device=None) -> torch.Tensor:
    """Row permutation of a [2F, K] GEGLU projection (value rows then gate rows) into the kernel's layout: blocks of 64 rows =
    32 value rows followed by their 32 gate rows."""
    assert F % 32 == 0
    j = torch.arange(F // 32, device=device)[:, None] * 32 + torch.arange(32, device=device)[None, :]      # [F/32, 32]
    return torch.cat([j, j + F], dim=1).reshape(-1)


def linear(x, w, bias=None, *, residual=None, geglu=False, lora_tail=None, out=None):
    """y = x @ w.T (+ bias) (+ residual)   or, with ``geglu``, value * gelu(gate) of an INTERLEAVED projection (see
    ``geglu_interleave_index``).  x [..., K] contiguous, w [N, K], bias fp32 [N]; residual / out [..., N_out].
    ``lora_tail`` = (t [M, 64], ups: list of [N, 64] tensors or None per batch row, rows_per_batch)."""
    _dev(x, w, bias, residual, out)
    K = x.shape[-1]
    N = w.shape[0]
    M = x.numel() // K
    n_out = N // 2 if geglu else N
    assert w.shape[1] == K and x.is_contiguous() and w.is_contiguous() and w.dtype == x.dtype
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == N
    if out is None:
        out = torch.empty(*x.shape[:-1], n_out, dtype=x.dtype, device=x.device)
    assert out.is_contiguous() and out.numel() == M * n_out and out.dtype == x.dtype
    if residual is not None:
        assert residual.dtype == x.dtype and residual.is_contiguous() and residual.numel() == M * N
    t_ptr, up_arr, rpb, nb = None, None, 0, 0
    if lora_tail is not None:
        t, ups, rpb = lora_tail
        nb = len(ups)
        assert t.dtype == x.dtype and t.is_contiguous() and t.shape[-1] == 64 and t.numel() >= M * 64 and rpb * nb == M
        for u in ups:
            assert u is None or (u.dtype == x.dtype and u.is_contiguous() and tuple(u.shape) == (N, 64))
        up_arr = (C.c_void_p * nb)(*[None if u is None else u.data_ptr() for u in ups])
        t_ptr = t.data_ptr()
    tag = f"M{M}_N{N}_K{K}" + ("_geglu" if geglu else "") + ("_res" if residual is not None else "") + ("_lora" if lora_tail is not None else "")
    ws = _linear_workspace(x.device.index)
    with _Launch("linear", 1, tag, 2.0 * M * N * K):
        rc = _lib.load().tmx_linear_fwd(_p(x), _p(w), _p(bias), _p(residual), _p(out), M, N, K, K, N, n_out,
                                        _lib.EPI_GEGLU if geglu else _lib.EPI_NONE, t_ptr, up_arr, int(rpb), int(nb),
                                        _p(ws), _dt(x), _stream())
    _lib.check(rc, "tmx_linear_fwd")
    return out


_lora_t_buf = {}


def lora_t(x, downs, sr: int):
    """t = x[b] @ down[b].T per routed batch row into the shared [B*M, 64] operand buffer of the LoRA tail (columns beyond
    ``sr`` stay zero; un-routed rows keep stale values the GEMM never reads)."""
    _dev(x)
    B, M, K = x.shape
    assert x.is_contiguous() and len(downs) == B
    key = (x.device.index, x.dtype, B * M)
    t = _lora_t_buf.get(key)
    if t is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("tmx: the LoRA tail buffer must be allocated before CUDA-graph capture; run the op once eagerly first")
        t = _lora_t_buf[key] = torch.zeros(B * M, 64, dtype=x.dtype, device=x.device)
    for d in downs:
        assert d is None or (d.dtype == x.dtype and d.is_contiguous() and tuple(d.shape) == (sr, K))
    for b0 in range(0, B, MAX_ROUTED_ROWS):
        nb = min(MAX_ROUTED_ROWS, B - b0)
        arr = (C.c_void_p * nb)(*[None if d is None else d.data_ptr() for d in downs[b0:b0 + nb]])
        with _Launch("lora_t", 1, f"M{M}_K{K}", 0.0):
            rc = _lib.load().tmx_lora_t_fwd(_p(x[b0:]), arr, t[b0 * M:].data_ptr(), nb, M, K, K, int(sr), _dt(x), _stream())
        _lib.check(rc, "tmx_lora_t_fwd")
    return t


# ------------------------------------------------------------------------------------------ k11 / k12 (video loop)

def vpred_cfg_ddim(x, v_uncond, v_cond, a_t: float, a_next: float, g: float, *, out=None, x0_out=None, ref_rounding=False):
    """CFG + v-prediction Tweedie x0 + DDIM update of the video latents (video_gen/pipeline_i2vgen_xl.py:694-713); any shape,
    all tensors contiguous and of one dtype (fp16 / bf16 / fp32).  Returns x_next."""
    _dev(x, v_uncond, v_cond, out, x0_out)
    assert x.is_contiguous() and v_uncond.is_contiguous() and v_cond.is_contiguous()
    assert x.dtype == v_uncond.dtype == v_cond.dtype and x.shape == v_uncond.shape == v_cond.shape
    if out is None:
        out = torch.empty_like(x)
    assert out.is_contiguous() and out.dtype == x.dtype and (x0_out is None or (x0_out.is_contiguous() and x0_out.dtype == x.dtype))
    n = x.numel()
    with _Launch("vpred", 1, f"n{n}", (4 + (1 if x0_out is not None else 0)) * n * x.element_size()):
        rc = _lib.load().tmx_vpred_cfg_ddim_fwd(_p(x), _p(v_uncond), _p(v_cond), _p(out), _p(x0_out), n, float(a_t), float(a_next), float(g),
                                                _dt(x), _lib.ROUND_REF if ref_rounding else _lib.ROUND_FP32, _stream())
    _lib.check(rc, "tmx_vpred_cfg_ddim_fwd")
    return out


def frame_inject(y, groups: int, frames: int, interp: float = 1.0, *, ref_rounding=False):
    """In place on a dense [(groups frames), ...] 16-bit tensor: frames 1.. of every group become
    interp * frame 0 + (1 - interp) * themselves (video_gen/utils_attn.py:433-456)."""
    _dev(y)
    if y.shape[0] != groups * frames:
        raise RuntimeError(f"tmx.frame_inject: leading dim {y.shape[0]} is not groups x frames = {groups} x {frames}")
    dense = y.is_contiguous() or (y.dim() == 4 and y.is_contiguous(memory_format=torch.channels_last))
    if not dense:
        raise RuntimeError("tmx.frame_inject: y must be dense (contiguous or channels_last)")
    with _Launch("frame_inject", 1, "", 2.0 * y.numel() * y.element_size()):
        rc = _lib.load().tmx_frame_inject_fwd(_p(y), int(groups), int(frames), y.numel() // (groups * frames), float(interp), _dt(y),
                                              _lib.ROUND_REF if ref_rounding else _lib.ROUND_FP32, _stream())
    _lib.check(rc, "tmx_frame_inject_fwd")
    return y


# ------------------------------------------------------------------------------------------ k1/k2

def attention(q, k, v, heads: int, scale: float | None = None, out=None):
    """softmax(scale * q k^T) v per head, head_dim 64.  q [B,Nq,H*64], k/v [B,Nk,H*64] — the
    diffusers layout *before* head_to_batch_dim (utils_custom.py:73-91); last dim contiguous,
    token stride free (so q/k/v may be column slices of one fused projection)."""
    _dev(q, k, v, out)
    B, Nq, HD = q.shape
    Nk = k.shape[1]
    D = HD // heads
    assert q.dtype == k.dtype == v.dtype and q.stride(2) == k.stride(2) == v.stride(2) == 1
    for t, n in ((q, Nq), (k, Nk), (v, Nk)):
        assert t.stride(0) == n * t.stride(1), "batch stride must equal N * token stride"
    if out is None:
        out = torch.empty((B, Nq, HD), dtype=q.dtype, device=q.device)
    if scale is None:
        scale = D ** -0.5
    with _Launch("attention", 1, f"Nq{Nq}_Nk{Nk}_H{heads}", 4.0 * B * heads * Nq * Nk * D):
        rc = _lib.load().tmx_attn_fwd(_p(q), _p(k), _p(v), _p(out), B, heads, Nq, Nk, D,
                                      q.stride(1), k.stride(1), v.stride(1), out.stride(1),
                                      float(scale), _dt(q), _stream())
    _lib.check(rc, "tmx_attn_fwd")
    return out
