"""Video stage: the product hook layer (tweediemix_b200.video_gen.utils_attn) and step (pipeline_step.VideoStepper) against
the golden vectors minted from the reference's own code (tests/golden/make_golden_video.py).

CPU: kernels replaced by the fp32 stand-ins of fake_ops (wiring).  -m gpu: the real k11 / k12 kernels through the C ABI —
TMX_ROUND_REF reproduces the reference's fp16 pipeline bit for bit, the default mode (fp32 arithmetic, one rounding) stays within one
16-bit ulp of it; the injection copy is exact, the blend within one ulp."""
import os
import types

import pytest
import torch

import fake_ops
from oracle import video_ref as V


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def _model(gd, device="cpu", dtype=torch.float32):
    unet = V.VideoUNetStub(c=16, temb=32, seed=0)
    unet.load_state_dict(gd["state_dict"])
    return types.SimpleNamespace(unet=unet.to(device, dtype))


def _check_hooks(gd, model, device, dtype, atol):
    from tweediemix_b200.video_gen import utils_attn as U
    U.register_conv_control_efficient(model, gd["schedule"], gd["interp"])
    mv = lambda t: t.to(device, dtype)
    u = model.unet
    for t in (981, 941, 1000):
        U.register_time(model, t)
        for got, key in ((u.mid_block.resnets[0].forward(mv(gd["x_mid"]), mv(gd["temb"])), f"mid0_t{t}"),
                         (u.mid_block.resnets[1].forward(mv(gd["x_mid"]), mv(gd["temb"])), f"mid1_t{t}"),
                         (u.up_blocks[1].resnets[0].forward(mv(gd["x_up"]), mv(gd["temb"])), f"up10_t{t}")):
            torch.testing.assert_close(got.float().cpu(), gd[key], atol=atol, rtol=0)
    assert sorted(n for n, m in u.named_modules() if hasattr(m, "t")) == gd["stamped"]
    m = u.up_blocks[1].resnets[0]
    assert m.injection_schedule is None and m.injection_schedule2 is gd["schedule"] and m.interp == gd["interp"]
    assert u.mid_block.resnets[0].injection_schedule is gd["schedule"] and u.mid_block.resnets[0].injection_schedule2 is None
    U.register_conv_control_efficient(model, gd["schedule"], gd["interp"])             # re-registration must not stack the injection
    U.register_time(model, 981)
    torch.testing.assert_close(u.up_blocks[1].resnets[0].forward(mv(gd["x_up"]), mv(gd["temb"])).float().cpu(), gd["up10_t981"], atol=atol, rtol=0)


@torch.no_grad()
def test_product_video_hooks_match_reference(monkeypatch, golden_dir):
    fake_ops.install(monkeypatch)
    gd = _load(golden_dir, "video_inject.pt")
    _check_hooks(gd, _model(gd), "cpu", torch.float32, 2e-5)


@torch.no_grad()
def test_video_stepper_matches_reference_lines(monkeypatch, golden_dir):
    fake_ops.install(monkeypatch)
    from tweediemix_b200.video_gen.pipeline_step import VideoStepper
    gd = _load(golden_dir, "video_step.pt")
    ts = list(range(981, 0, -20))
    st = VideoStepper(gd["alphas_cumprod"], float(gd["alphas_cumprod"][0]), ts, 9.0, injection_timestep=0.02)
    assert st.skip == 20 and st.injection_schedule == [981] and st.alpha(-19) == pytest.approx(float(gd["alphas_cumprod"][0]))
    for t in (981, 21, 1):
        rec = gd[f"f32_t{t}"]
        got = st.step(rec["latents_in"], rec["noise_pred"], t)
        torch.testing.assert_close(got, rec["latents_out"], atol=5e-6, rtol=0)


def _stub_unet_forward(model, dtype):
    """A U-Net-shaped callable over the stub's hooked blocks: [2, C, 16, H, W] latents -> v-prediction of the same shape, every hooked
    ResNet on the way (mid_block.resnets[0, 1], up_blocks[1].resnets[0, 1]), so the injection window changes the trajectory."""
    u = model.unet

    def fwd(x5, t):
        b, c, f, h, w = x5.shape
        x = x5.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
        temb = torch.sin(float(t) * 0.01 * torch.arange(1, 33, dtype=torch.float32)).to(x.device, dtype)[None].expand(b * f, -1)
        y = u.mid_block.resnets[1].forward(u.mid_block.resnets[0].forward(x, temb), temb)
        y = u.up_blocks[1].resnets[0].forward(torch.cat([y, x], 1), temb)
        y = u.up_blocks[1].resnets[1].forward(y, temb)
        return (0.5 * y).reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)
    return fwd


def _loop_pair(gd_inject, gd_step, device, dtype, n_steps=10, ref_rounding=False):
    """Product loop (product hooks + VideoStepper.denoise_loop) and oracle loop (oracle/video_ref.video_loop_ref) on two copies of the stub."""
    from tweediemix_b200.video_gen.pipeline_step import VideoStepper
    ac = gd_step["alphas_cumprod"]
    ts = list(range(1000 - 1000 // n_steps + 1, 0, -(1000 // n_steps)))              # leading spacing, offset 1: 901, 801, ... 1
    g = torch.Generator().manual_seed(3)
    lat = torch.randn(1, 16, 16, 6, 5, generator=g)
    prod, orac = _model(gd_inject, device, dtype), _model(gd_inject, "cpu", torch.float32)
    st = VideoStepper(ac, float(ac[0]), ts, 9.0, injection_timestep=0.25, ref_rounding=ref_rounding)
    assert st.injection_schedule == ts[:2]
    seen = []
    got = st.denoise_loop(prod, lat.to(device, dtype), _stub_unet_forward(prod, dtype), interp_ratio=0.7, callback=lambda i, t, x: seen.append(t))
    want = V.video_loop_ref(orac, lat, _stub_unet_forward(orac, torch.float32), ac, ac[0], ts, 9.0, 0.25, 0.7)
    assert seen == ts and [m.t for m in (prod.unet.mid_block.resnets[0], prod.unet.up_blocks[1].resnets[1])] == [1, 1]
    # and the window matters: the same loop without injection lands elsewhere
    none = V.video_loop_ref(_model(gd_inject, "cpu", torch.float32), lat, _stub_unet_forward(_model(gd_inject, "cpu", torch.float32), torch.float32), ac, ac[0], ts, 9.0, 0.0, 0.7)
    return got.float().cpu(), want, none


@torch.no_grad()
def test_video_denoise_loop_matches_oracle_loop(monkeypatch, golden_dir):
    """pipeline_i2vgen_xl.py:655-656,677-719 as one call: hooks installed for the window, register_time per step, CFG halves, fused step."""
    fake_ops.install(monkeypatch)
    got, want, none = _loop_pair(_load(golden_dir, "video_inject.pt"), _load(golden_dir, "video_step.pt"), "cpu", torch.float32)
    torch.testing.assert_close(got, want, atol=2e-4, rtol=0)
    assert (want - none).abs().max() > 1e-2


# ----------------------------------------------------------------------------------------- GPU
def _ops():
    from tweediemix_b200 import build, ops
    build.build()
    return ops


@pytest.mark.gpu
@torch.no_grad()
def test_vpred_kernel_vs_reference_lines(golden_dir):
    o = _ops()
    from tweediemix_b200.video_gen.pipeline_step import VideoStepper
    gd = _load(golden_dir, "video_step.pt")
    ts = list(range(981, 0, -20))
    for t in (981, 21, 1):
        rec = gd[f"f16_t{t}"]
        lat, npd = rec["latents_in"].cuda(), rec["noise_pred"].cuda()
        # The reference runs on CUDA, where a 0-dim fp32 tensor times an fp16 tensor keeps the scalar in fp32 (on the CPU, where the golden
        # was minted, TensorIterator first rounds it to fp16): the bit-exact yardstick is the oracle's torch ops executed on the GPU, the
        # CPU golden agrees to an fp16 ulp or two.
        alphas = gd["alphas_cumprod"]
        want_gpu, x0_gpu = V.vpred_step_ref(lat, npd, alphas[t], alphas[t - 20] if t - 20 >= 0 else alphas[0], 9.0)
        want = rec["latents_out"].float()
        torch.testing.assert_close(want_gpu.float().cpu(), want, rtol=2 ** -8, atol=1e-2)      # CPU vs CUDA scalar handling: a few fp16 ulps
        ref_mode = VideoStepper(gd["alphas_cumprod"], float(gd["alphas_cumprod"][0]), ts, 9.0, ref_rounding=True).step(lat, npd, t)
        assert torch.equal(ref_mode, want_gpu.contiguous()), "TMX_ROUND_REF must reproduce the reference's fp16 roundings"
        fast = VideoStepper(gd["alphas_cumprod"], float(gd["alphas_cumprod"][0]), ts, 9.0).step(lat, npd, t)
        # fp32 arithmetic + one rounding vs the reference's eleven fp16 roundings (guidance 9 amplifies them): a few fp16 ulps
        torch.testing.assert_close(fast.float().cpu(), want, rtol=2 ** -8, atol=2e-2)
        rec32 = gd[f"f32_t{t}"]
        got32 = VideoStepper(gd["alphas_cumprod"], float(gd["alphas_cumprod"][0]), ts, 9.0).step(rec32["latents_in"].cuda(), rec32["noise_pred"].cuda(), t)
        torch.testing.assert_close(got32.cpu(), rec32["latents_out"], rtol=1e-5, atol=1e-5)
        x0 = torch.empty_like(lat)
        o.vpred_cfg_ddim(lat, npd[:1].contiguous(), npd[1:].contiguous(), float(gd["alphas_cumprod"][t]), float(gd["alphas_cumprod"][max(t - 20, 0)]), 9.0,
                         x0_out=x0, ref_rounding=True)
        assert torch.equal(x0, x0_gpu.contiguous())
    # size-independent property at the BASELINE size (1280x720, 16 frames: 4 x 16 x 90 x 160 latents): guidance 0 + equal alphas = identity
    big = torch.randn(1, 4, 16, 90, 160, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(2, 4, 16, 90, 160, device="cuda", dtype=torch.bfloat16)
    same = o.vpred_cfg_ddim(big, v[:1].contiguous(), v[1:].contiguous(), 0.5, 0.5, 0.0)
    torch.testing.assert_close(same.float(), big.float(), rtol=2 ** -7, atol=2 ** -7)
    with pytest.raises(RuntimeError, match="multiple of 8"):
        o.vpred_cfg_ddim(torch.zeros(12, device="cuda"), torch.zeros(12, device="cuda"), torch.zeros(12, device="cuda"), 0.5, 0.5, 1.0)


@pytest.mark.gpu
@torch.no_grad()
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_frame_inject_kernel(dtype):
    o = _ops()
    g = torch.Generator().manual_seed(0)
    for shape, cl in (((32, 16, 4, 4), False), ((32, 1280, 16, 16), True), ((32, 24, 3, 8), False)):
        y = torch.randn(shape, generator=g).to(dtype).cuda()
        if cl:
            y = y.contiguous(memory_format=torch.channels_last)
        ref = V.inject_ref(y.float().cpu(), None)
        got = o.frame_inject(y.clone(memory_format=torch.preserve_format), 2, 16, 1.0)
        assert torch.equal(got.float().cpu(), ref)                                        # plain copy of frame 0: exact
        want = V.inject_ref(y, 0.7).float().cpu()                                         # torch's own 16-bit roundings, on the GPU like the reference
        got_ref = o.frame_inject(y.clone(memory_format=torch.preserve_format), 2, 16, 0.7, ref_rounding=True)
        assert torch.equal(got_ref.float().cpu(), want)
        got_fast = o.frame_inject(y.clone(memory_format=torch.preserve_format), 2, 16, 0.7)
        # fp32 blend + one rounding vs torch's three 16-bit roundings: within one ulp at the magnitude of the data (|y| < 8)
        torch.testing.assert_close(got_fast.float().cpu(), want, rtol=0, atol=2 ** -5 if dtype == torch.bfloat16 else 2 ** -8)
    with pytest.raises(RuntimeError, match="groups x frames"):
        o.frame_inject(torch.zeros(30, 8, 2, 2, dtype=dtype).cuda(), 2, 16)


@pytest.mark.gpu
@torch.no_grad()
def test_product_video_hooks_match_reference_gpu(golden_dir):
    _ops()
    gd = _load(golden_dir, "video_inject.pt")
    _check_hooks(gd, _model(gd, "cuda", torch.float16), "cuda", torch.float16, 2e-2)


@pytest.mark.gpu
@torch.no_grad()
def test_video_denoise_loop_matches_oracle_loop_gpu(golden_dir):
    """The loop on the real k11 / k12 kernels (fp16 U-Net stand-in on CUDA) against the fp32 oracle loop: 10 chained steps stay within
    fp16 rounding of the blocks' outputs (|latent| ~ 1)."""
    _ops()
    got, want, none = _loop_pair(_load(golden_dir, "video_inject.pt"), _load(golden_dir, "video_step.pt"), "cuda", torch.float16)
    assert torch.isfinite(got).all()
    scale = want.abs().max().item()
    assert (got - want).abs().max().item() <= 3e-2 * scale, ((got - want).abs().max().item(), scale)
    assert (want - none).abs().max() > 1e-2
