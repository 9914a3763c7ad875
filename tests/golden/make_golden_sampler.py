#!/usr/bin/env python
"""Mint sampler-level golden vectors by running the REFERENCE's own ``Tweediemix`` methods.

    python tests/golden/make_golden_sampler.py      # needs /root/reference (build container only)

``fusion_generation/fusion_sampling.py`` / ``fusion_sampling_lora.py`` cannot be imported as they are
(``diffusers`` and ``sentence_transformers`` are absent, ``model_lora`` needs a removed transformers
class), but nothing on the hot path needs those packages: they only provide the pipeline loaders that
``Tweediemix.__init__`` calls.  This script therefore
  * registers EMPTY stub modules under those names so that the reference files import,
  * builds a ``Tweediemix`` object WITHOUT running ``__init__`` and fills in exactly the attributes
    ``__init__`` would leave behind (``fusion_sampling.py:97-223``): the U-Net (the diffusers-shaped
    stand-in, seeded weights), the per-concept U-Nets, the DDIM tables (scaled_linear / leading /
    offset 1, restated here from the SDXL scheduler config, then shifted as ``:218`` does), ``skip``,
    ``final_alpha_cumprod``, text embeddings, ``add_time_ids``,
  * and then calls the reference's UNMODIFIED ``init_fusion`` (which installs the reference's own
    hooks), ``alpha`` and ``denoise_step`` for every timestep, inside the same ``torch.autocast``
    context ``sample_loop`` opens (``:491-494``; a no-op on this CPU-only box, so everything is fp32).
The segmentation subprocess (``os.system`` at ``:457``) is replaced by a no-op and the mask files it
would have written are the reference's own ``example_results/test_out`` JPEGs, placed in
``output_path`` beforehand, so the reference's ``preprocess_mask`` (``:81-89``) ingests real files;
``decode_latent`` (VAE, out of scope) returns a blank image.

Fixtures written (inputs are re-derived from seeds by the tests; only outputs are stored):
  sampler_custom_n5.pt    config 1: 256x256, 5 DDIM steps, K=3 (cat+dog+bg), custom variant — latent after every step
  sampler_lora_n10.pt     LoRA variant, 10 steps, t_stop 0.8 (exercises the t_stop window and quirk 7)
  step_math_ref.pt        single ``denoise_step`` calls in each of the four phases on a closed-form "U-Net"
"""
import importlib.util
import os
import shutil
import sys
import tempfile
import types
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("TMX_REFERENCE", "/root/reference")
REF_FG = os.path.join(REF, "fusion_generation")

from oracle import synth  # noqa: E402
from oracle.hooks_ref import make_lora_set  # noqa: E402
from oracle.unet_ref import UNetConfig, transformer_blocks_in_hook_order  # noqa: E402

CFG = UNetConfig.tiny()
K = 3
BASE_SEED, TEXT_SEED = 4321, 78


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m


def load_reference_sampler(lora: bool):
    """Import the unmodified reference file with its unavailable third-party imports stubbed out."""
    _stub("diffusers", DDIMScheduler=object, StableDiffusionXLPipeline=object, UNet2DConditionModel=object, AutoencoderKL=object)
    _stub("diffusers.image_processor", VaeImageProcessor=object)
    _stub("sentence_transformers")
    _stub("sentence_transformers.util", semantic_search=None, dot_score=None, normalize_embeddings=None)
    _stub("model_lora", create_lora_diffusion_base=None)
    if REF_FG not in sys.path:
        sys.path.insert(0, REF_FG)            # `from utils_custom import *` / `from utils_lora import *` -> the real files
    name = "fusion_sampling_lora" if lora else "fusion_sampling"
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF_FG, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ddim_tables(n):
    """SDXL scheduler config [D]: scaled_linear betas 0.00085..0.012 over 1000 steps, leading spacing, steps_offset 1."""
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
    acp = torch.cumprod(1.0 - betas, dim=0)
    ratio = 1000 // n
    timesteps = (torch.arange(0, n) * ratio).round().flip(0).to(torch.int64) + 1
    return acp, timesteps


def build(ref, base, concept_unets, n, res, outdir, lora, t_stop=None, resampling=2, jumping=2):
    tw = object.__new__(ref.Tweediemix)
    torch.nn.Module.__init__(tw)
    base.device = torch.device("cpu")                      # diffusers ModelMixin.device
    tw.unet = base
    for i, u in enumerate(concept_unets):
        setattr(tw, f"unet_{i}", u)
    acp, timesteps = ddim_tables(n)
    tw.scheduler = types.SimpleNamespace(timesteps=timesteps, alphas_cumprod=acp, final_alpha_cumprod=acp[0],
                                         init_noise_sigma=1.0)
    tw.skip = 1000 // n                                                         # :213-216
    tw.final_alpha_cumprod = tw.scheduler.final_alpha_cumprod                   # :217
    tw.scheduler.alphas_cumprod = torch.cat([torch.tensor([1.0]), acp])         # :218
    tw.concept_num = K
    tw.add_time_ids = torch.tensor([[res, res, 0, 0, res, res]])                # compute_time_ids :70-78
    tw.text_embeds, tw.text_embeds_single = synth.make_text(CFG, K, TEXT_SEED)
    tw.masks = None                                                             # set by the reference at t_cond_prev
    tw.config = types.SimpleNamespace(guidance_scale=0.8, n_timesteps=n, t_cond=0.2, t_stop=t_stop,
                                      resampling_steps=resampling, jumping_steps=jumping,
                                      resolution_h=res, resolution_w=res, output_path=outdir, output_path_all=outdir,
                                      seg_gpu=0, seg_concepts="a cat+a dog")
    tw.decode_latent = lambda latent: torch.zeros(1, 3, 8, 8)                   # VAE is out of scope (:297-303)
    return tw


def run_loop(ref, tw, x):
    """The body of the reference's sample_loop (:491-494), without the VAE tail that follows it."""
    real_system = os.system
    os.system = lambda cmd: 0                                                   # the segmentation subprocess (:457)
    xs = []
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with torch.no_grad(), torch.autocast(device_type="cuda", dtype=torch.float16):
                for t in tw.scheduler.timesteps:
                    x = tw.denoise_step(x, t)
                    xs.append(x.clone())
    finally:
        os.system = real_system
    return xs


def place_masks(outdir):
    src = os.path.join(HERE, "masks", "test_out")
    for f in os.listdir(src):
        shutil.copy(os.path.join(src, f), os.path.join(outdir, f))


class ClosedFormUNet(torch.nn.Module):
    """A U-Net-shaped function with a known closed form, so single steps can be checked without weights."""
    device = torch.device("cpu")

    def forward(self, sample, t, encoder_hidden_states=None, added_cond_kwargs=None):
        tt = float(t) / 1000.0
        bias = encoder_hidden_states.mean(dim=(1, 2)).reshape(-1, 1, 1, 1)
        pool = added_cond_kwargs["text_embeds"].mean(dim=1).reshape(-1, 1, 1, 1)
        return {"sample": torch.sin(3.0 * sample + tt) * 0.7 + 0.3 * bias - 0.2 * pool * sample}


@torch.no_grad()
def main():
    # ------------------------------------------------------------------ custom variant, config 1
    ref = load_reference_sampler(lora=False)
    n, res = 5, 256
    base = synth.make_base_unet(CFG, BASE_SEED)
    concepts = [synth.make_concept_unet(base, 300 + i) for i in range(K)]
    with tempfile.TemporaryDirectory() as out:
        place_masks(out)
        tw = build(ref, base, concepts, n, res, out, lora=False)
        tw.init_fusion(int(n * 0.2))                                            # reference :476-483 (installs ITS hooks)
        torch.manual_seed(3821)
        x0 = torch.randn(1, 4, res // 8, res // 8)
        xs = run_loop(ref, tw, x0)
        torch.save({"x0": x0, "xs": torch.stack(xs), "masks": tw.masks.clone(), "timesteps": tw.scheduler.timesteps.clone(),
                    "alphas": torch.stack([tw.alpha(t) for t in tw.scheduler.timesteps]),
                    "t_cond": (int(tw.t_cond_prev), int(tw.t_cond_cur), int(tw.start_t)), "skip": tw.skip,
                    "meta": dict(n=n, res=res, base_seed=BASE_SEED, concept_seeds=[300, 301, 302], text_seed=TEXT_SEED,
                                 resampling_steps=2, jumping_steps=2, guidance_scale=0.8, t_cond=0.2)},
                   os.path.join(HERE, "sampler_custom_n5.pt"))
        print("custom n=5: final |x|max", xs[-1].abs().max().item())

        # -------------------------------------------------------------- single steps, closed-form U-Net, 50-step schedule
        tw = build(ref, ClosedFormUNet(), [ClosedFormUNet() for _ in range(K)], 50, 128, out, lora=False, resampling=3, jumping=0)
        ts = tw.scheduler.timesteps
        tw.t_cond = ts[10:]
        tw.t_cond_prev, tw.t_cond_cur, tw.start_t = ts[9], ts[10], ts[0]        # init_fusion :477-480 without the hook install
        sys.modules["ref_fusion_sampling"] = ref
        ref.register_time = lambda model, t: None                                # closed-form U-Net has no attention modules
        g = torch.Generator().manual_seed(11)
        x = torch.randn(1, 4, 16, 16, generator=g)
        tw.masks = synth.fixture_masks(16, 16)
        cases = {}
        for name, t in (("start", ts[0]), ("plain", ts[5]), ("fused", ts[20]), ("last", ts[49])):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                with torch.autocast(device_type="cuda", dtype=torch.float16):
                    cases[name] = {"t": int(t), "out": tw.denoise_step(x.clone(), t)}
        torch.save({"x": x, "cases": cases, "text_seed": TEXT_SEED}, os.path.join(HERE, "step_math_ref.pt"))

    # ------------------------------------------------------------------ LoRA variant
    ref = load_reference_sampler(lora=True)
    n, res = 10, 128
    base = synth.make_base_unet(CFG, BASE_SEED)
    lora_sets = [make_lora_set(base, 400 + i) for i in range(K)]
    concepts = []
    for s in lora_sets:                      # reference reads unet_i....attn{1,2}.processor.to_*_lora (utils_lora.py:139-149)
        u = synth.make_base_unet(CFG, BASE_SEED)
        for name, blk in transformer_blocks_in_hook_order(u):
            for which in ("attn1", "attn2"):
                getattr(blk, which).processor = types.SimpleNamespace(**s[f"{name}.{which}"])
        concepts.append(u)
    with tempfile.TemporaryDirectory() as out:
        place_masks(out)
        tw = build(ref, base, concepts, n, res, out, lora=True, t_stop=0.8)
        tw.init_fusion(t_cond=int(n * 0.2), t_stop=int(n * 0.8))                # reference _lora.py:476-485
        torch.manual_seed(3828)
        x0 = torch.randn(1, 4, res // 8, res // 8)
        xs = run_loop(ref, tw, x0)
        torch.save({"x0": x0, "xs": torch.stack(xs), "masks": tw.masks.clone(), "timesteps": tw.scheduler.timesteps.clone(),
                    "t_stop_cur": int(tw.t_stop_cur),
                    "meta": dict(n=n, res=res, base_seed=BASE_SEED, lora_seeds=[400, 401, 402], text_seed=TEXT_SEED,
                                 resampling_steps=2, jumping_steps=2, guidance_scale=0.8, t_cond=0.2, t_stop=0.8)},
                   os.path.join(HERE, "sampler_lora_n10.pt"))
        print("lora n=10: final |x|max", xs[-1].abs().max().item())


if __name__ == "__main__":
    main()
