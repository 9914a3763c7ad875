"""The video-stage oracle (oracle/video_ref.py) against golden vectors minted from the REFERENCE's own code
(tests/golden/make_golden_video.py): the unmodified ``video_gen/utils_attn.py`` hooks, and the step lines of
``video_gen/pipeline_i2vgen_xl.py`` exec'd verbatim.  fp32 goldens: round-off only (2e-5); fp16 goldens: bit-exact (same torch ops)."""
import os
import types

import pytest
import torch

from oracle import video_ref as V


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def _model(gd):
    unet = V.VideoUNetStub(c=16, temb=32, seed=0)
    unet.load_state_dict(gd["state_dict"])
    return types.SimpleNamespace(unet=unet)


@torch.no_grad()
def test_injection_hooks_match_reference(golden_dir):
    gd = _load(golden_dir, "video_inject.pt")
    model = _model(gd)
    V.register_conv_ref(model, gd["schedule"], gd["interp"])
    for t in (981, 941, 1000):
        V.register_time_ref(model, t)
        u = model.unet
        torch.testing.assert_close(u.mid_block.resnets[0].forward(gd["x_mid"], gd["temb"]), gd[f"mid0_t{t}"], atol=2e-5, rtol=0)
        torch.testing.assert_close(u.mid_block.resnets[1].forward(gd["x_mid"], gd["temb"]), gd[f"mid1_t{t}"], atol=2e-5, rtol=0)
        torch.testing.assert_close(u.up_blocks[1].resnets[0].forward(gd["x_up"], gd["temb"]), gd[f"up10_t{t}"], atol=2e-5, rtol=0)
    assert sorted(n for n, m in model.unet.named_modules() if hasattr(m, "t")) == gd["stamped"]
    # the injection is observable: inside the window (and at t == 1000) frames 1.. of the mid blocks equal frame 0
    o = gd["mid0_t981"].reshape(2, 16, *gd["mid0_t981"].shape[1:])
    assert torch.equal(o[:, 1:], o[:, :1].expand_as(o[:, 1:])) and not torch.equal(gd["mid0_t941"], gd["mid0_t981"])
    assert torch.equal(gd["mid0_t1000"], gd["mid0_t981"])


@pytest.mark.parametrize("name,atol", [("f32", 2e-6), ("f16", 0.0)])
def test_vpred_step_matches_reference_lines(golden_dir, name, atol):
    gd = _load(golden_dir, "video_step.pt")
    alphas = gd["alphas_cumprod"]
    for t in (981, 21, 1):
        rec = gd[f"{name}_t{t}"]
        at = alphas[t]
        at_next = alphas[t - 20] if t - 20 >= 0 else alphas[0]
        got, x0 = V.vpred_step_ref(rec["latents_in"], rec["noise_pred"], at, at_next, 9.0)
        torch.testing.assert_close(got.float(), rec["latents_out"].float(), atol=atol, rtol=0)
        b, c, f, h, w = x0.shape                                   # the reference leaves `denoised_tweedie` in [(b f), c, h, w]
        torch.testing.assert_close(x0.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w).float(), rec["x0"].float(), atol=atol, rtol=0)
        assert got.shape == rec["latents_in"].shape and got.dtype == rec["latents_in"].dtype
