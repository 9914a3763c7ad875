"""-m gpu: k10 (persistent tcgen05 GEMM with fused epilogues, csrc/linear.cu) through the C ABI against an fp32 matmul of the
same 16-bit inputs.  Tolerance: fp32 accumulation + ONE rounding of the output dtype (rtol 2^-7 bf16 / 2^-10 fp16; atol a
few output ulps at the magnitude of the result).  Both tile widths (128 / 256 columns) are forced in turn; shapes cover the
SDXL sites of one fused step, M / N tails that do not fill a tile, a single-tile problem and more tiles than SMs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

RTOL = {torch.bfloat16: 2 ** -7, torch.float16: 2 ** -10}


def _ops():
    from tweediemix_b200 import build, ops
    build.build()
    return ops


def _variant(bn):
    from tweediemix_b200 import _lib
    assert _lib.load().tmx_linear_set_variant(bn) == 0


def _mk(shape, g, dtype, scale=1.0):
    return (torch.randn(shape, generator=g) * scale).to(dtype).cuda()


SHAPES = [  # M, N, K
    (4096, 1280, 5120), (4096, 3840, 1280), (4096, 1280, 1280), (4096, 1280, 5120), (16384, 1920, 640), (16384, 640, 640), (16384, 640, 2560),
    (1024, 1280, 1280), (128, 128, 64), (200, 136, 128), (77, 2560, 2048), (300, 72, 192), (20000, 256, 64),
]


@pytest.mark.parametrize("bn", [0, 128, 192, 256, 320, 1000, 1256])      # + 1000: split-K tail disabled
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", SHAPES)
def test_linear_bias_residual(shape, dtype, bn):
    o = _ops()
    _variant(bn)
    try:
        M, N, K = shape
        g = torch.Generator().manual_seed(M + N + K)
        x, w = _mk((M, K), g, dtype), _mk((N, K), g, dtype, K ** -0.5)
        bias = torch.randn(N, generator=g).cuda()
        res = _mk((M, N), g, dtype)
        want = x.float() @ w.float().t()
        atol = 4 * RTOL[dtype]
        torch.testing.assert_close(o.linear(x, w).float(), want, rtol=RTOL[dtype], atol=atol)
        torch.testing.assert_close(o.linear(x, w, bias).float(), want + bias, rtol=RTOL[dtype], atol=atol)
        torch.testing.assert_close(o.linear(x, w, bias, residual=res).float(), want + bias + res.float(), rtol=RTOL[dtype], atol=2 * atol)
        torch.testing.assert_close(o.linear(x, w, None, residual=res).float(), want + res.float(), rtol=RTOL[dtype], atol=2 * atol)
    finally:
        _variant(0)


@pytest.mark.parametrize("bn", [0, 128, 256, 1256])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(4096, 10240, 1280), (16384, 5120, 640), (256, 1024, 128), (100, 192, 64), (1024, 10240, 1280)])
def test_linear_geglu_epilogue(shape, dtype, bn):
    """value * gelu_erf(gate) of an interleaved projection == the un-fused reference (projection rounded to 16 bits, then
    gated): [D] GEGLU.forward."""
    o = _ops()
    _variant(bn)
    try:
        M, N, K = shape
        Fh = N // 2
        g = torch.Generator().manual_seed(N)
        x, w = _mk((M, K), g, dtype), _mk((N, K), g, dtype, K ** -0.5)
        bias = (torch.randn(N, generator=g) * 0.1).cuda()
        idx = o.geglu_interleave_index(Fh, "cuda")
        got = o.linear(x, w[idx].contiguous(), bias[idx].contiguous(), geglu=True)
        proj = (x.float() @ w.float().t() + bias).to(dtype).float()
        want = proj[:, :Fh] * F.gelu(proj[:, Fh:])
        assert got.shape == (M, Fh)
        torch.testing.assert_close(got.float(), want, rtol=2 * RTOL[dtype], atol=8 * RTOL[dtype])
    finally:
        _variant(0)


@pytest.mark.parametrize("bn", [0, 128, 192, 256, 320, 1256])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,Mb,N,K,nseg", [(4, 1024, 3840, 1280, 3), (4, 1024, 1280, 1280, 1), (3, 4096, 1920, 640, 3), (2, 128, 64, 64, 1), (1, 1024, 1280, 1280, 1)])
def test_linear_lora_tail(B, Mb, N, K, nseg, dtype, bn):
    """Rank-4 LoRA deltas as a K = 16 tail step of the GEMM (+ the skinny t = x . down^T kernel) == utils_lora.py:65-79:
    y[b] = x[b] W^T + per segment (x[b] down_s^T) up_s^T for routed rows, row 0 untouched."""
    o = _ops()
    _variant(bn)
    try:
        r = 4
        g = torch.Generator().manual_seed(B * N)
        x, w = _mk((B, Mb, K), g, dtype), _mk((N, K), g, dtype, K ** -0.5)
        first = [None] if B > 1 else []                     # (B == 1: a rank that owns one ROUTED row, concept-parallel)
        downs = first + [_mk((nseg * r, K), g, dtype, 1.0 / r) for _ in range(B - len(first))]
        ups = first + [_mk((N, r), g, dtype, 0.05) for _ in range(B - len(first))]
        bias = torch.randn(N, generator=g).cuda()
        res = _mk((B, Mb, N), g, dtype)
        from tweediemix_b200.routing import LoRARouting
        rt = LoRARouting.__new__(LoRARouting)
        rt.rows, rt._lists, rt._subsets, rt.cache_tag = [None] * B, {"w": (downs, ups)}, {}, 0
        tail = rt.tail("w", nseg, x)
        assert tail is not None and tail[2] == Mb
        got = o.linear(x, w, bias, residual=res, lora_tail=tail)
        seg = N // nseg
        want = x.float() @ w.float().t() + bias + res.float()
        for b in range(len(first), B):
            t = (x[b].float() @ downs[b].float().t()).to(dtype).float()          # the tail consumes t rounded to 16 bits
            for s in range(nseg):
                want[b, :, s * seg:(s + 1) * seg] += t[:, s * r:(s + 1) * r] @ ups[b][s * seg:(s + 1) * seg].float().t()
        torch.testing.assert_close(got.float(), want, rtol=RTOL[dtype], atol=8 * RTOL[dtype])
    finally:
        _variant(0)


@pytest.mark.parametrize("bn", [0, 128, 256, 2256])
@pytest.mark.parametrize("shape", [(4096, 1280, 5120), (1024, 1280, 5120), (16384, 640, 2560), (2048, 1280, 5120), (4096, 1280, 1280)])
def test_split_k_tail_is_deterministic_and_rearms(shape, bn):
    """Shapes whose tile count is not a multiple of the SM count take the split-K tail (cooperative launch, fp32 partials summed in
    slice order): repeated launches are bit-identical (fixed summation order, counters re-armed) and agree with the unsplit kernel
    to fp32 accumulation-order round-off."""
    o = _ops()
    M, N, K = shape
    g = torch.Generator().manual_seed(9)
    x, w = _mk((M, K), g, torch.bfloat16), _mk((N, K), g, torch.bfloat16, K ** -0.5)
    bias, res = torch.randn(N, generator=g).cuda(), _mk((M, N), g, torch.bfloat16)
    _variant(bn)
    runs = [o.linear(x, w, bias, residual=res) for _ in range(4)]
    assert all(torch.equal(runs[0], r) for r in runs[1:])
    _variant(1000 + bn % 1000)
    try:
        whole = o.linear(x, w, bias, residual=res)
    finally:
        _variant(0)
    torch.testing.assert_close(runs[0].float(), whole.float(), rtol=2 ** -7, atol=2 ** -6)
    torch.testing.assert_close(runs[0].float(), x.float() @ w.float().t() + bias + res.float(), rtol=2 ** -7, atol=2 ** -5)


def test_linear_rejects_bad_arguments():
    o = _ops()
    z = lambda *s: torch.zeros(*s, dtype=torch.bfloat16).cuda()
    with pytest.raises(RuntimeError, match="multiple of 64"):
        o.linear(z(8, 48), z(8, 48))
    with pytest.raises(RuntimeError, match="multiple of 8"):
        o.linear(z(8, 64), z(12, 64))
    with pytest.raises(RuntimeError, match="GEGLU epilogue takes no residual"):
        o.linear(z(8, 64), z(64, 64), geglu=True, residual=z(8, 64))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        o.linear(torch.zeros(8, 64, dtype=torch.bfloat16), z(8, 64))
