"""-m gpu: parity of the bandwidth kernels (k4/k5 GroupNorm+SiLU, k6 residual add, k7 fused
CFG/Tweedie/blend/DDIM) against the CPU oracle, through the C ABI.

Tolerances (stated per test):
  k7, TMX_ROUND_REF : <= 4 fp32 ulp of the oracle evaluated with the same eps dtype (the kernel
                      re-rounds exactly where torch's promotion does; torch-CUDA divides by
                      multiplying with a reciprocal, the CPU oracle and the kernel divide)
  k7, TMX_ROUND_FP32: 2e-5 relative vs the fp32 oracle on upcast eps
  GroupNorm         : fp32 I/O 3e-5 abs; 16-bit I/O one rounding of the output dtype
  residual add      : exact for fp32, one output rounding for 16-bit
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import step_math as sm
from oracle import synth

pytestmark = pytest.mark.gpu


def ops():
    from tweediemix_b200 import build, ops as o
    build.build()
    return o


DTYPES = [torch.float16, torch.bfloat16, torch.float32]


def _blend_inputs(K, h, w, dtype, seed, imgs=1):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(imgs, 4, h, w, generator=g)
    eps = torch.randn(imgs, K + 1, 4, h, w, generator=g).to(dtype)
    return x, eps


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("K,hw", [(3, (128, 128)), (3, (32, 32)), (8, (128, 128)), (1, (16, 8))])
def test_blend_ref_rounding_matches_oracle(dtype, K, hw):
    """TMX_ROUND_REF reproduces the reference's own mixed-precision arithmetic.  The reference runs
    fusion_sampling.py:376-386,430 with CUDA tensors and 0-dim fp32 *CPU* alphas (App. B); the oracle
    functions are device-agnostic restatements of those lines, so evaluating them on CUDA tensors here
    IS the reference arithmetic as torch executes it on this GPU (fp16 products rounded once from an
    fp32 scalar, division by a CPU scalar done as multiplication by its reciprocal).  Bit-exact up to
    fma contraction: tolerance 2 fp32 ulp of the largest value.  Against the CPU evaluation of the same
    oracle (which rounds the scalar to 16 bits first) the tolerance is one 16-bit ulp of eps amplified
    by sqrt(1-a)/sqrt(a)."""
    o = ops()
    h, w = hw
    masks = synth.fixture_masks(h, w) if K == 3 else synth.stripe_masks(K, h, w)
    x, eps = _blend_inputs(K, h, w, dtype, 11)
    ulp16 = {torch.float16: 2.0 ** -11, torch.bfloat16: 2.0 ** -8, torch.float32: 2.0 ** -24}[dtype]
    for (at, an, last) in [(0.043827, 0.051787, False), (0.005844, 0.007365, False), (0.99915, 0.99915, True)]:
        x0 = torch.empty_like(x).cuda()
        got = o.tweedie_blend_ddim(x.cuda(), eps.cuda(), masks.cuda(), at, an, 0.8, is_last=last,
                                   x0_out=x0, ref_rounding=True).cpu()
        ref_dev, ref0_dev = sm.fused_step(x.cuda(), eps[0].cuda(), masks.cuda(), at, an, 0.8, is_last=last)
        scale = max(ref_dev.abs().max().item(), 1.0)
        assert (got - ref_dev.cpu()).abs().max().item() <= 2 * 1.2e-7 * scale, (dtype, K, at, "vs torch-CUDA reference arithmetic")
        assert (x0.cpu() - ref0_dev.cpu()).abs().max().item() <= 2 * 1.2e-7 * max(ref0_dev.abs().max().item(), 1.0)
        want, _ = sm.fused_step(x, eps[0], masks, at, an, 0.8, is_last=last)          # CPU evaluation
        amp = ((1 - at) ** 0.5 / at ** 0.5) * eps.float().abs().max().item()
        assert (got - want).abs().max().item() <= 4 * ulp16 * amp + 1e-5 * scale, (dtype, K, at, "vs CPU oracle")


@pytest.mark.parametrize("dtype", DTYPES)
def test_blend_fp32_mode_and_batched_images(dtype):
    o = ops()
    K, h, w, imgs = 3, 64, 64, 5
    masks = synth.fixture_masks(h, w)
    x, eps = _blend_inputs(K, h, w, dtype, 12, imgs)
    got = o.tweedie_blend_ddim(x.cuda(), eps.cuda(), masks.cuda(), 0.3, 0.4, 0.8).cpu()
    for i in range(imgs):
        want, _ = sm.fused_step(x[i:i + 1], eps[i].float(), masks, 0.3, 0.4, 0.8)
        torch.testing.assert_close(got[i:i + 1], want, rtol=2e-5, atol=2e-5)


def test_blend_weight_and_nullmask_forms():
    """resampling x0 (:392-403), plain CFG (:425-430) and re-noise (:407-412) through the same kernel."""
    o = ops()
    K, h, w = 3, 32, 32
    x, eps = _blend_inputs(K, h, w, torch.float32, 13)
    at, an, g = 0.0058, 0.0074, 0.8
    # resample: rows [uncond, multi, single_1, single_2]; weights [K-1, -1, -1]
    x0 = torch.empty_like(x).cuda()
    got = o.tweedie_blend_ddim(x.cuda(), eps.cuda(), None, at, an, g, weights=[K - 1, -1, -1], x0_out=x0).cpu()
    want0 = sm.resample_x0(x, eps[0], at, g, K)
    torch.testing.assert_close(x0.cpu(), want0, rtol=2e-5, atol=2e-4)
    torch.testing.assert_close(got, sm.ddim_update(want0, eps[0, :1], an), rtol=2e-5, atol=2e-4)
    # plain CFG
    got = o.tweedie_blend_ddim(x.cuda(), eps[:, :2].contiguous().cuda(), None, at, an, g).cpu()
    want, _ = sm.cfg_step(x, eps[0, :2], at, an, g)
    torch.testing.assert_close(got, want, rtol=2e-5, atol=2e-4)
    # re-noise = CFG step with the alphas swapped
    got = o.tweedie_blend_ddim(x.cuda(), eps[:, :2].contiguous().cuda(), None, an, at, g).cpu()
    torch.testing.assert_close(got, sm.renoise(x, eps[0, :2], at, an, g), rtol=2e-5, atol=2e-4)


@pytest.mark.parametrize("dtype", DTYPES)
def test_blend_partial_finish_equals_fused(dtype):
    """Concept-parallel split: any assignment of rows to 'ranks' sums to the fused result."""
    o = ops()
    K, h, w = 5, 32, 32
    g_ = torch.Generator().manual_seed(3)
    masks = (torch.rand(K, 1, h, w, generator=g_) > 0.4).float()
    x, eps = _blend_inputs(K, h, w, dtype, 14)
    want, want0 = sm.fused_step(x, eps[0].float(), masks, 0.05, 0.06, 0.8)
    assignment = [[0, 3], [1, 2, 5], [4]]
    total = torch.zeros(1, 2, 4, h, w, device="cuda")
    for rows in assignment:
        acc = torch.empty(1, 2, 4, h, w, device="cuda")
        o.blend_partial(eps[:, rows].contiguous().cuda(), masks.cuda(), rows, acc)
        total += acc                                   # stands in for the all-reduce
    x0 = torch.empty_like(x).cuda()
    got = o.blend_finish(x.cuda(), total, masks.cuda(), 0.05, 0.06, 0.8, x0_out=x0).cpu()
    torch.testing.assert_close(x0.cpu(), want0, rtol=3e-5, atol=3e-4)
    torch.testing.assert_close(got, want, rtol=3e-5, atol=3e-4)


def test_blend_in_place_and_errors():
    o = ops()
    x, eps = _blend_inputs(3, 16, 16, torch.bfloat16, 15)
    masks = synth.fixture_masks(16, 16)
    want, _ = sm.fused_step(x, eps[0].float(), masks, 0.3, 0.4, 0.8)
    xc = x.cuda()
    o.tweedie_blend_ddim(xc, eps.cuda(), masks.cuda(), 0.3, 0.4, 0.8, out=xc)
    torch.testing.assert_close(xc.cpu(), want, rtol=2e-5, atol=2e-5)
    with pytest.raises(RuntimeError, match="multiple of 8"):
        o.tweedie_blend_ddim(torch.zeros(1, 4, 3, 3).cuda(), torch.zeros(4, 4, 3, 3).cuda(), None, 0.3, 0.4, 0.8)
    with pytest.raises(RuntimeError, match="alphas"):
        o.tweedie_blend_ddim(xc, eps.cuda(), masks.cuda(), 0.0, 0.4, 0.8)


GN_SHAPES = [  # (N, C, H, W) — SDXL sites at reduced spatial size plus ragged / tiny cases
    (4, 320, 32, 32), (2, 640, 16, 16), (4, 1280, 8, 8), (2, 960, 16, 24), (1, 1920, 8, 8),
    (2, 2560, 8, 8), (3, 64, 5, 7), (1, 32, 1, 8), (2, 320, 128, 128),
]


def _gn_ref(x, gamma, beta, groups, eps, add, silu):
    xf = x.float()
    if add is not None:
        xf = xf + add[:, :, None, None]
    y = F.group_norm(xf, groups, gamma, beta, eps)
    return F.silu(y) if silu else y


@pytest.fixture(params=[0, 1, 3], ids=["default", "two_pass", "fused"])
def gn_variant(request):
    """0: one CTA per (sample, group) where that slab fits in shared memory, else the single cooperative launch (CTA slabs + grid
    barrier) where it applies; 1: stats + apply launches; 3: the per-group kernel off (cooperative kernel / two launches)."""
    from tweediemix_b200 import _lib
    ops()
    _lib.load().tmx_groupnorm_set_variant(request.param)
    yield request.param
    _lib.load().tmx_groupnorm_set_variant(0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("layout", ["nhwc", "nchw"])
@pytest.mark.parametrize("shape", GN_SHAPES + [(4, 640, 64, 64), (4, 1280, 32, 32)])
def test_groupnorm_matches_fp32_reference(dtype, layout, shape, gn_variant):
    o = ops()
    N, C, H, W = shape
    if gn_variant != 0 and (layout == "nchw" or dtype == torch.float32):
        pytest.skip("variant only changes the NHWC 16-bit path")
    if layout == "nchw" and (H * W) % 8:
        pytest.skip("NCHW kernel requires HW % 8 == 0 (checked in test_groupnorm_errors)")
    g = torch.Generator().manual_seed(C + H)
    x = (torch.randn(shape, generator=g) * 1.7 + 0.6 * torch.randn(1, C, 1, 1, generator=g)).to(dtype)
    gamma, beta = 1 + 0.3 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g)
    add = 0.5 * torch.randn(N, C, generator=g)
    for silu, use_add, eps in [(True, False, 1e-5), (False, False, 1e-6), (True, True, 1e-5)]:
        want = _gn_ref(x, gamma, beta, 32, eps, add if use_add else None, silu)
        xc = x.cuda()
        if layout == "nhwc":
            xc = xc.contiguous(memory_format=torch.channels_last)
        got = o.group_norm(xc, gamma.cuda(), beta.cuda(), 32, eps, silu=silu, add=add.cuda() if use_add else None)
        assert got.stride() == xc.stride()
        if dtype == torch.float32:
            torch.testing.assert_close(got.cpu(), want, rtol=3e-5, atol=3e-5)
        else:
            # one rounding of the output dtype (+ fp32 round-off before it)
            torch.testing.assert_close(got.cpu().float(), want.to(dtype).float(),
                                       rtol=2 ** -7 if dtype == torch.bfloat16 else 2 ** -10, atol=2e-3)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(4, 1280, 1280, 32, 32), (4, 1280, 640, 32, 32), (4, 640, 320, 64, 64), (2, 640, 320, 128, 128),
                                   (4, 320, 320, 128, 128), (2, 64, 128, 9, 7), (1, 8, 56, 3, 3)])   # N, C1, C2, H, W: the SDXL up-block sites + ragged
def test_groupnorm_two_sources_equals_cat(dtype, shape, gn_variant):
    """group_norm(a, x2=b) == group_norm(cat([a, b], 1)): the up-block ResNets read (hidden, skip) without the torch.cat copy; groups may
    straddle the source boundary (1280 + 640 channels: 60 per group)."""
    o = ops()
    N, C1, C2, H, W = shape
    C = C1 + C2
    g = torch.Generator().manual_seed(C + H)
    a = (torch.randn(N, C1, H, W, generator=g) * 1.3 + 0.4).to(dtype)
    b = (torch.randn(N, C2, H, W, generator=g) * 0.7 - 0.2).to(dtype)
    gamma, beta = 1 + 0.3 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g)
    add = 0.5 * torch.randn(N, C, generator=g)
    ac, bc = (t.cuda().contiguous(memory_format=torch.channels_last) for t in (a, b))
    for silu, use_add in [(True, False), (True, True), (False, False)]:
        want = _gn_ref(torch.cat([a, b], 1), gamma, beta, 32 if C % 32 == 0 else 8, 1e-5, add if use_add else None, silu)
        got = o.group_norm(ac, gamma.cuda(), beta.cuda(), 32 if C % 32 == 0 else 8, 1e-5, silu=silu, add=add.cuda() if use_add else None, x2=bc)
        assert got.shape == (N, C, H, W) and got.is_contiguous(memory_format=torch.channels_last)
        torch.testing.assert_close(got.cpu().float(), want.to(dtype).float(), rtol=2 ** -7 if dtype == torch.bfloat16 else 2 ** -10, atol=2e-3)
        # and bit-identical to the single-source kernel on the materialised concatenation (same arithmetic, same order)
        cat = torch.cat([a, b], 1).cuda().contiguous(memory_format=torch.channels_last)
        one = o.group_norm(cat, gamma.cuda(), beta.cuda(), 32 if C % 32 == 0 else 8, 1e-5, silu=silu, add=add.cuda() if use_add else None)
        assert torch.equal(got, one)


def test_groupnorm_large_mean_is_stable_and_inplace(gn_variant):
    """|mean| >> std: the pivot-shifted sums must not cancel catastrophically."""
    o = ops()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 320, 16, 16, generator=g) * 0.01 + 100.0
    gamma, beta = torch.ones(320), torch.zeros(320)
    want = _gn_ref(x, gamma, beta, 32, 1e-5, None, False)
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    o.group_norm(xc, gamma.cuda(), beta.cuda(), 32, 1e-5, out=xc)
    torch.testing.assert_close(xc.cpu(), want, rtol=2e-3, atol=2e-3)
    # 16-bit, in place, launched back to back (the barrier words in the workspace are reused by every launch)
    xb = (torch.randn(4, 1280, 32, 32, generator=g) * 0.5 + 3.0).to(torch.bfloat16)
    wantb = _gn_ref(xb, torch.ones(1280), torch.zeros(1280), 32, 1e-5, None, True)
    for _ in range(3):
        xc = xb.cuda().contiguous(memory_format=torch.channels_last)
        o.group_norm(xc, torch.ones(1280).cuda(), torch.zeros(1280).cuda(), 32, 1e-5, silu=True, out=xc)
        torch.testing.assert_close(xc.cpu().float(), wantb.to(torch.bfloat16).float(), rtol=2 ** -7, atol=2e-3)


def test_groupnorm_errors():
    o = ops()
    with pytest.raises(RuntimeError, match="HW=35"):
        o.group_norm(torch.zeros(1, 64, 5, 7).cuda(), torch.ones(64).cuda(), torch.zeros(64).cuda(), 32, 1e-5)
    with pytest.raises(RuntimeError, match="not divisible"):
        o.group_norm(torch.zeros(1, 40, 8, 8).cuda(), torch.ones(40).cuda(), torch.zeros(40).cuda(), 32, 1e-5)


@pytest.mark.parametrize("dtype", DTYPES)
def test_residual_add(dtype):
    o = ops()
    g = torch.Generator().manual_seed(8)
    for n in [(4, 320, 32, 32), (1, 8), (3, 1280, 24)]:
        a, b = torch.randn(n, generator=g).to(dtype), torch.randn(n, generator=g).to(dtype)
        for inv in (1.0, 0.5):
            want = ((a.float() + b.float()) * inv).to(dtype)
            got = o.residual_add(a.cuda(), b.cuda(), inv).cpu()
            assert torch.equal(got, want)
    with pytest.raises(RuntimeError, match="multiple of 8"):
        o.residual_add(torch.zeros(7).cuda(), torch.zeros(7).cuda())


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(4, 4096, 640), (4, 1024, 1280), (3, 5, 64), (1, 1, 8), (2, 7, 2048), (9, 1000)])
def test_layernorm_matches_fp32_reference(dtype, shape):
    """k8 vs F.layer_norm evaluated in fp32 on the same 16-bit input: one rounding of the output dtype."""
    o = ops()
    g = torch.Generator().manual_seed(shape[-1])
    x = (torch.randn(shape, generator=g) * 2.0 + 3.0).to(dtype)
    D = shape[-1]
    gamma, beta = 1 + 0.3 * torch.randn(D, generator=g), 0.2 * torch.randn(D, generator=g)
    want = F.layer_norm(x.float(), (D,), gamma, beta, 1e-5)
    got = o.layer_norm(x.cuda(), gamma.cuda(), beta.cuda(), 1e-5)
    torch.testing.assert_close(got.cpu().float(), want.to(dtype).float(),
                               rtol=2 ** -7 if dtype == torch.bfloat16 else 2 ** -10, atol=2e-3)
    xc = x.cuda()
    o.layer_norm(xc, gamma.cuda(), beta.cuda(), 1e-5, out=xc)          # in place
    assert torch.equal(xc.cpu(), got.cpu())


def test_layernorm_errors():
    o = ops()
    with pytest.raises(RuntimeError, match="multiple of 8"):
        o.layer_norm(torch.zeros(2, 12, dtype=torch.bfloat16).cuda(), torch.ones(12).cuda(), torch.zeros(12).cuda(), 1e-5)
    with pytest.raises(RuntimeError, match="<= 2048"):
        o.layer_norm(torch.zeros(2, 4096, dtype=torch.bfloat16).cuda(), torch.ones(4096).cuda(), torch.zeros(4096).cuda(), 1e-5)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_bias_residual_add(dtype):
    """k6b: (a + bias[c] (+ b)) * inv_scale on channels_last tensors — fp32 sum, one rounding (bit-exact vs torch fp32)."""
    o = ops()
    g = torch.Generator().manual_seed(21)
    for shape in [(4, 320, 16, 16), (2, 1280, 8, 8), (1, 8, 2, 2)]:
        a = torch.randn(shape, generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
        b = torch.randn(shape, generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
        bias = torch.randn(shape[1], generator=g)
        for inv in (1.0, 0.5):
            want = ((a.float() + bias.reshape(1, -1, 1, 1) + b.float()) * inv).to(dtype)
            got = o.bias_residual_add(a.cuda(), bias.cuda(), b.cuda(), inv).cpu()
            assert torch.equal(got, want)
            want1 = ((a.float() + bias.reshape(1, -1, 1, 1)) * inv).to(dtype)
            ac = a.cuda()
            got1 = o.bias_residual_add(ac, bias.cuda(), None, inv, out=ac).cpu()        # in place, no residual
            assert torch.equal(got1, want1)
    with pytest.raises(RuntimeError, match="multiple of 8"):
        o.bias_residual_add(torch.zeros(2, 6, dtype=dtype).cuda(), torch.zeros(6).cuda())


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(4, 4096, 640), (4, 1024, 1280), (3, 5, 64), (1, 1, 8), (2, 7, 2048)])
def test_residual_add_layernorm(dtype, shape):
    """k6c: h = round(a + b) must be bit-exact; n must equal the stand-alone LayerNorm kernel applied to h
    bit-for-bit (same arithmetic on the same rounded row) and F.layer_norm in fp32 within one output rounding."""
    o = ops()
    g = torch.Generator().manual_seed(shape[-1] + 1)
    a = (torch.randn(shape, generator=g) * 2.0).to(dtype)
    b = (torch.randn(shape, generator=g) + 1.0).to(dtype)
    D = shape[-1]
    gamma, beta = 1 + 0.3 * torch.randn(D, generator=g), 0.2 * torch.randn(D, generator=g)
    h_want = (a.float() + b.float()).to(dtype)
    n_want = F.layer_norm(h_want.float(), (D,), gamma, beta, 1e-5)
    ac = a.cuda()
    h, n = o.residual_add_layer_norm(ac, b.cuda(), gamma.cuda(), beta.cuda(), 1e-5, h_out=ac)   # h aliases a
    assert h.data_ptr() == ac.data_ptr()
    assert torch.equal(h.cpu(), h_want)
    torch.testing.assert_close(n.cpu().float(), n_want.to(dtype).float(), rtol=2 ** -7 if dtype == torch.bfloat16 else 2 ** -10, atol=2e-3)
    assert torch.equal(n.cpu(), o.layer_norm(h_want.cuda(), gamma.cuda(), beta.cuda(), 1e-5).cpu())


def test_blend_full_size_properties_2048_images():
    """BASELINE-size properties of k7 that need no oracle run (2048 stacked 128x128 images, 2.1 GB):
    (1) per-image independence — every image of the stack equals the single-image launch on its slice;
    (2) partition property — when all concept rows carry the SAME eps, a perfect partition of masks blends to the
        plain single-prompt CFG step (fusion_sampling.py:376-386 collapses to :425-430);
    (3) t == 1 returns x0 (:471-472)."""
    o = ops()
    K, h, w, imgs = 3, 128, 128, 2048
    masks = synth.fixture_masks(h, w).cuda()
    assert torch.equal(masks.sum(0), torch.ones_like(masks[0]))          # the example masks are a perfect partition
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(imgs, 4, h, w, generator=g, device="cuda")
    eps = torch.randn(imgs, K + 1, 4, h, w, generator=g, device="cuda").to(torch.bfloat16)
    got = o.tweedie_blend_ddim(x, eps, masks, 0.0438, 0.0518, 0.8)
    for i in (0, 777, imgs - 1):
        one = o.tweedie_blend_ddim(x[i:i + 1].contiguous(), eps[i:i + 1].contiguous(), masks, 0.0438, 0.0518, 0.8)
        assert torch.equal(got[i:i + 1], one)
    same = eps.clone()
    same[:, 2:] = same[:, 1:2]
    blended = o.tweedie_blend_ddim(x, same, masks, 0.0438, 0.0518, 0.8)
    plain = o.tweedie_blend_ddim(x, same[:, :2].contiguous(), None, 0.0438, 0.0518, 0.8)
    torch.testing.assert_close(blended, plain, rtol=2e-5, atol=2e-4)
    x0 = torch.empty_like(x)
    last = o.tweedie_blend_ddim(x, eps, masks, 0.99915, 0.99915, 0.8, is_last=True, x0_out=x0)
    assert torch.equal(last, x0)


# ------------------------------------------------------------------------------------------ k3
def _routed_ref(x, weights, downs, ups, nseg, base=None):
    """fp32 restatement of utils_custom.py:64-82 (per-row weights) / utils_lora.py:65-79 (rank-r deltas per row)."""
    B = x.shape[0]
    y = torch.stack([x[b].float() @ weights[b].float().t() for b in range(B)]) if weights is not None else base.float().clone()
    if downs is not None:
        seg = y.shape[-1] // nseg
        for b in range(B):
            if downs[b] is None:
                continue
            r = downs[b].shape[0] // nseg
            t = x[b].float() @ downs[b].float().t()
            for s in range(nseg):
                y[b, :, s * seg:(s + 1) * seg] += t[:, s * r:(s + 1) * r] @ ups[b][s * seg:(s + 1) * seg].float().t()
    return y


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(4, 77, 2048, 2560), (4, 77, 2048, 1280), (2, 5, 64, 8), (3, 130, 128, 136), (1, 256, 640, 640),
                                   (9, 77, 2048, 2560)])   # B, M, Kin, Nout; the first two are the SDXL cross-attention K/V sites
def test_routed_linear_grouped_gemm(dtype, shape):
    """k3 grouped tcgen05 GEMM: one weight matrix per batch row, vs fp32 matmul of the same 16-bit inputs.
    Tolerance: fp32 accumulation, one rounding of the output dtype."""
    o = ops()
    B, M, Kin, Nout = shape
    g = torch.Generator().manual_seed(M + Nout)
    x = torch.randn(B, M, Kin, generator=g).to(dtype).cuda()
    ws = [(torch.randn(Nout, Kin, generator=g) / Kin ** 0.5).to(dtype).cuda() for _ in range(B)]
    got = o.routed_linear(x, ws)
    want = _routed_ref(x, ws, None, None, 1)
    assert got.shape == (B, M, Nout) and torch.isfinite(got.float()).all()
    torch.testing.assert_close(got.float(), want.to(dtype).float(), rtol=2 ** -7 if dtype == torch.bfloat16 else 2 ** -10, atol=2e-3)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(4, 1024, 1280, 3840, 3), (4, 4096, 640, 1920, 3), (4, 77, 2048, 2560, 2), (4, 1024, 1280, 1280, 1),
                                   (2, 37, 64, 64, 1), (3, 5, 128, 48, 3)])   # B, M, Kin, Nout, nseg — rank 4 (model_lora.py:28-48)
def test_routed_linear_lora_delta(dtype, shape):
    """k3 rank-r delta kernel: row 0 (unconditional) untouched, rows 1.. get segment-wise (x down^T) up^T added in place."""
    o = ops()
    B, M, Kin, Nout, nseg = shape
    r = 4
    g = torch.Generator().manual_seed(M + Nout + nseg)
    x = torch.randn(B, M, Kin, generator=g).to(dtype).cuda()
    base = torch.randn(B, M, Nout, generator=g).to(dtype).cuda()
    downs = [None] + [(torch.randn(nseg * r, Kin, generator=g) / r).to(dtype).cuda() for _ in range(B - 1)]
    ups = [None] + [(torch.randn(Nout, r, generator=g) * 0.05).to(dtype).cuda() for _ in range(B - 1)]
    y = base.clone()
    got = o.routed_linear(x, None, downs, ups, nseg=nseg, out=y)
    assert got.data_ptr() == y.data_ptr() and torch.equal(y[0], base[0])
    want = _routed_ref(x, None, downs, ups, nseg, base=base)
    torch.testing.assert_close(y.float(), want.to(dtype).float(), rtol=2 ** -7 if dtype == torch.bfloat16 else 2 ** -10, atol=4e-3)


def test_routed_linear_weights_plus_lora_and_errors():
    o = ops()
    g = torch.Generator().manual_seed(2)
    B, M, Kin, Nout, r = 3, 77, 128, 256, 4
    x = torch.randn(B, M, Kin, generator=g).to(torch.bfloat16).cuda()
    ws = [(torch.randn(Nout, Kin, generator=g) / Kin ** 0.5).to(torch.bfloat16).cuda() for _ in range(B)]
    downs = [None] + [(torch.randn(2 * r, Kin, generator=g) / r).to(torch.bfloat16).cuda() for _ in range(B - 1)]
    ups = [None] + [(torch.randn(Nout, r, generator=g) * 0.05).to(torch.bfloat16).cuda() for _ in range(B - 1)]
    got = o.routed_linear(x, ws, downs, ups, nseg=2)
    # the GEMM result is rounded to bf16 before the delta is added (two kernels): allow two output roundings
    want = _routed_ref(x, ws, downs, ups, 2)
    torch.testing.assert_close(got.float(), want, rtol=2 ** -6, atol=6e-3)
    with pytest.raises(RuntimeError, match="multiple of 64"):
        o.routed_linear(torch.zeros(1, 4, 48, dtype=torch.bfloat16).cuda(), [torch.zeros(8, 48, dtype=torch.bfloat16).cuda()])
    # more than 16 batch rows (K = 8 concepts x image batch 4 on one GPU = 36 rows): the front-end splits the batch into ABI
    # calls of <= 16 rows; the ABI itself still refuses B = 17
    B2 = 19
    x2 = torch.randn(B2, 40, 64, generator=g).to(torch.bfloat16).cuda()
    ws2 = [(torch.randn(24, 64, generator=g) / 8).to(torch.bfloat16).cuda() for _ in range(B2)]
    d2 = [None if b % 3 == 0 else (torch.randn(r, 64, generator=g) / r).to(torch.bfloat16).cuda() for b in range(B2)]
    u2 = [None if b % 3 == 0 else (torch.randn(24, r, generator=g) * 0.05).to(torch.bfloat16).cuda() for b in range(B2)]
    torch.testing.assert_close(o.routed_linear(x2, ws2, d2, u2, nseg=1).float(), _routed_ref(x2, ws2, d2, u2, 1), rtol=2 ** -6, atol=6e-3)
    from tweediemix_b200 import _lib
    import ctypes as C
    ptrs = (C.c_void_p * 17)(*[ws2[0].data_ptr()] * 17)
    rc = _lib.load().tmx_routed_linear_fwd(x2.data_ptr(), ptrs, None, None, x2.data_ptr(), 17, 4, 64, 24, 0, 1, _lib.BF16, None)
    assert rc == -2 and "B=17" in _lib.last_error()


# ------------------------------------------------------------------------------------------ k13 / k14 (layout kernels)
LAYOUT_CAT = [(4, 1280, 1280, 32, 32), (4, 1280, 640, 32, 32), (4, 640, 320, 64, 64), (4, 320, 320, 128, 128), (2, 640, 320, 128, 128),
              (1, 8, 24, 3, 5), (3, 16, 8, 1, 1)]                                    # N, Ca, Cb, H, W: the SDXL up-block sites + ragged


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", LAYOUT_CAT)
def test_cat_channels_is_torch_cat(dtype, shape):
    """k13 == torch.cat([a, b], 1) bit for bit on channels_last tensors ([D] up blocks: torch.cat([hidden_states, res_hidden_states], dim=1))."""
    o = ops()
    N, Ca, Cb, H, W = shape
    g = torch.Generator().manual_seed(Ca + Cb + H)
    a = torch.randn(N, Ca, H, W, generator=g).to(dtype).cuda().contiguous(memory_format=torch.channels_last)
    b = torch.randn(N, Cb, H, W, generator=g).to(dtype).cuda().contiguous(memory_format=torch.channels_last)
    got = o.cat_channels(a, b)
    assert got.shape == (N, Ca + Cb, H, W) and got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got, torch.cat([a, b], dim=1))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(4, 1280, 32, 32), (4, 640, 64, 64), (2, 320, 64, 64), (1, 8, 3, 5), (3, 24, 1, 1), (2, 16, 7, 1)])
def test_upsample_nearest2x_is_interpolate(dtype, shape):
    """k14 == F.interpolate(scale_factor=2, mode="nearest") bit for bit ([D] Upsample2D)."""
    o = ops()
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g).to(dtype).cuda().contiguous(memory_format=torch.channels_last)
    got = o.upsample_nearest2x(x)
    want = F.interpolate(x, scale_factor=2.0, mode="nearest")
    assert got.shape == want.shape and got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got, want)


def test_layout_kernels_errors():
    o = ops()
    a = torch.zeros(1, 12, 2, 2, dtype=torch.bfloat16).cuda().contiguous(memory_format=torch.channels_last)      # C % 8 != 0
    assert not o.layout_supported(a)
    with pytest.raises(RuntimeError, match="channels_last"):
        o.cat_channels(a, a)
    with pytest.raises(RuntimeError, match="channels_last"):
        o.upsample_nearest2x(torch.zeros(1, 8, 2, 2, dtype=torch.bfloat16).cuda())                               # NCHW-dense, H*W > 1
    x32 = torch.zeros(1, 8, 2, 2).cuda().contiguous(memory_format=torch.channels_last)
    assert not o.layout_supported(x32)
