#!/usr/bin/env python
"""Golden vectors for the video stage, minted from the REFERENCE's own code in the build container (needs /root/reference):

  * ``video_gen/utils_attn.py`` is imported unmodified (it needs only torch + einops) and its ``register_time`` /
    ``register_conv_control_efficient`` are applied to the stand-in of ``oracle/video_ref.py`` -> outputs of the three patched
    ResNet blocks inside and outside the injection window  -> ``video_inject.pt``;
  * the guidance + v-prediction Tweedie + DDIM lines of ``I2VGenXLPipeline.__call__`` (``pipeline_i2vgen_xl.py``, from
    ``# perform guidance`` to the final reshape of ``latents``) are extracted VERBATIM from the reference file and exec'd with a stub
    ``self`` (the module itself cannot be imported: diffusers is absent)                                   -> ``video_step.pt``.

    python tests/golden/make_golden_video.py
"""
import importlib.util
import os
import sys
import textwrap
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle.video_ref import VideoUNetStub  # noqa: E402


def load_ref_utils():
    spec = importlib.util.spec_from_file_location("ref_video_utils_attn", os.path.join(REF, "video_gen", "utils_attn.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@torch.no_grad()
def main():
    ref = load_ref_utils()
    unet = VideoUNetStub(c=16, temb=32, seed=0)
    model = types.SimpleNamespace(unet=unet)
    schedule = torch.tensor([981, 961])
    interp = 0.7
    ref.register_conv_control_efficient(model, schedule, interp)
    g = torch.Generator().manual_seed(1)
    x_mid = torch.randn(32, 16, 4, 4, generator=g)
    x_up = torch.randn(32, 32, 4, 4, generator=g)
    temb = torch.randn(32, 32, generator=g)
    out = {"x_mid": x_mid, "x_up": x_up, "temb": temb, "schedule": schedule, "interp": interp, "state_dict": unet.state_dict()}
    for t in (981, 941, 1000):
        ref.register_time(model, t)
        out[f"mid0_t{t}"] = unet.mid_block.resnets[0].forward(x_mid, temb)
        out[f"mid1_t{t}"] = unet.mid_block.resnets[1].forward(x_mid, temb)
        out[f"up10_t{t}"] = unet.up_blocks[1].resnets[0].forward(x_up, temb)
    out["stamped"] = sorted(n for n, m in unet.named_modules() if hasattr(m, "t"))
    torch.save(out, os.path.join(HERE, "video_inject.pt"))

    # ---- the step: the reference's own lines, exec'd
    src = open(os.path.join(REF, "video_gen", "pipeline_i2vgen_xl.py")).read().split("\n")
    start = next(i for i, l in enumerate(src) if "# perform guidance" in l)
    end = next(i for i, l in enumerate(src) if "latents = latents[None, :].reshape(batch_size, frames, channel, width, height)" in l)
    block = textwrap.dedent("\n".join(src[start:end + 1]))
    alphas = torch.cumprod(1.0 - torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000) ** 2, 0)
    stub = types.SimpleNamespace(do_classifier_free_guidance=True, skip=20, final_alpha_cumprod=alphas[0])
    stub.alpha = lambda t: alphas[t] if t >= 0 else stub.final_alpha_cumprod
    stub.scheduler = types.SimpleNamespace(step=lambda *a, **k: types.SimpleNamespace(pred_original_sample=None))
    res = {"alphas_cumprod": alphas, "block_source_lines": (start + 1, end + 1)}
    for dtype, name in ((torch.float32, "f32"), (torch.float16, "f16")):
        g = torch.Generator().manual_seed(2)
        latents = torch.randn(1, 4, 16, 8, 8, generator=g).to(dtype)
        noise_pred = torch.randn(2, 4, 16, 8, 8, generator=g).to(dtype)
        for t in (981, 21, 1):
            ns = {"self": stub, "noise_pred": noise_pred.clone(), "latents": latents.clone(), "guidance_scale": 9.0, "t": torch.tensor(t),
                  "extra_step_kwargs": {}, "torch": torch}
            exec(block, ns)
            res[f"{name}_t{t}"] = {"latents_in": latents, "noise_pred": noise_pred, "latents_out": ns["latents"], "x0": ns["denoised_tweedie"]}
    torch.save(res, os.path.join(HERE, "video_step.pt"))
    print("reference step lines", res["block_source_lines"])
    for f in ("video_inject.pt", "video_step.pt"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
