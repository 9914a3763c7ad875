"""Pins the sampler-level oracle (oracle/sampler_ref.py, oracle/step_math.py, oracle/schedule.py,
oracle/masks_ref.py) against the REFERENCE's own ``Tweediemix.init_fusion / alpha / denoise_step``
(fusion_sampling.py:305-483, fusion_sampling_lora.py:309-490), which
``tests/golden/make_golden_sampler.py`` runs UNMODIFIED on CPU (third-party imports stubbed, the
reference's own hooks installed by its own init_fusion, the reference's preprocess_mask reading the
reference's example mask JPEGs) and whose latents it stores in tests/golden/sampler_*.pt and
step_math_ref.pt.  Inputs are re-derived here from the same seeds.

Tolerance: both sides are fp32 on CPU; the oracle evaluates the same formulas with a different
association in a few places, so agreement is to rounding: 2e-5 relative to max|x| per step."""
import os

import pytest
import torch

from oracle import synth
from oracle.hooks_ref import make_lora_set, register_custom_ref, register_lora_ref
from oracle.sampler_ref import RefConfig, TweediemixRef
from oracle.unet_ref import UNetConfig

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CFG = UNetConfig.tiny()
K = 3
RTOL = 2e-5


def _rel(a, b):
    return (a - b).abs().max().item() / b.abs().max().item()


@torch.no_grad()
def test_custom_config1_matches_reference_run():
    gold = torch.load(os.path.join(G, "sampler_custom_n5.pt"))
    m = gold["meta"]
    base = synth.make_base_unet(CFG, m["base_seed"])
    donors = [synth.make_concept_unet(base, s) for s in m["concept_seeds"]]
    cfg = RefConfig(n_timesteps=m["n"], resolution_h=m["res"], resolution_w=m["res"], resampling_steps=m["resampling_steps"],
                    jumping_steps=m["jumping_steps"], guidance_scale=m["guidance_scale"], t_cond=m["t_cond"], seed=3821)
    text, single = synth.make_text(CFG, K, m["text_seed"])
    masks = synth.fixture_masks(m["res"] // 8, m["res"] // 8)
    assert torch.equal(masks, gold["masks"])                       # reference preprocess_mask + bg construction (:81-89,461-469)
    s = TweediemixRef(base, text, single, masks, cfg, K, lora=False, run_jump=True)
    assert [int(t) for t in s.sched.timesteps] == [int(t) for t in gold["timesteps"]]
    assert (s.t_cond_prev, s.t_cond_cur, s.start_t) == tuple(gold["t_cond"]) and s.skip == gold["skip"]
    for t, a in zip(gold["timesteps"], gold["alphas"]):            # alpha(): the +1-shifted table (:218,305-307)
        assert abs(float(s.alpha(int(t))) - float(a)) < 1e-7
    register_custom_ref(base, donors, torch.tensor(sorted(s.hook_gate_window(), reverse=True)), K)
    x0 = s.initial_latent()
    assert torch.equal(x0, gold["x0"])                             # CPU-generator draw after manual_seed (:488,587)
    got = []
    s.sample_loop(x0.clone(), callback=lambda i, t, x: got.append(x.clone()))
    for i, (a, b) in enumerate(zip(got, gold["xs"])):
        assert _rel(a, b) < RTOL, f"step {i}: rel {_rel(a, b)}"


@torch.no_grad()
def test_lora_matches_reference_run():
    gold = torch.load(os.path.join(G, "sampler_lora_n10.pt"))
    m = gold["meta"]
    base = synth.make_base_unet(CFG, m["base_seed"])
    cfg = RefConfig(n_timesteps=m["n"], resolution_h=m["res"], resolution_w=m["res"], resampling_steps=m["resampling_steps"],
                    jumping_steps=m["jumping_steps"], guidance_scale=m["guidance_scale"], t_cond=m["t_cond"], t_stop=m["t_stop"], seed=3828)
    text, single = synth.make_text(CFG, K, m["text_seed"])
    masks = synth.fixture_masks(m["res"] // 8, m["res"] // 8)
    assert torch.equal(masks, gold["masks"])
    s = TweediemixRef(base, text, single, masks, cfg, K, lora=True, run_jump=True)
    assert s.t_stop_cur == gold["t_stop_cur"]
    register_lora_ref(base, [make_lora_set(base, sd) for sd in m["lora_seeds"]],
                      torch.tensor(sorted(s.hook_gate_window(), reverse=True)), K)
    x0 = s.initial_latent()
    assert torch.equal(x0, gold["x0"])
    got = []
    s.sample_loop(x0.clone(), callback=lambda i, t, x: got.append(x.clone()))
    for i, (a, b) in enumerate(zip(got, gold["xs"])):
        assert _rel(a, b) < RTOL, f"step {i}: rel {_rel(a, b)}"


class _ClosedFormUNet(torch.nn.Module):
    """Same closed form as tests/golden/make_golden_sampler.py::ClosedFormUNet."""

    def forward(self, sample, t, encoder_hidden_states=None, added_cond_kwargs=None):
        tt = float(t) / 1000.0
        bias = encoder_hidden_states.mean(dim=(1, 2)).reshape(-1, 1, 1, 1)
        pool = added_cond_kwargs["text_embeds"].mean(dim=1).reshape(-1, 1, 1, 1)
        return {"sample": torch.sin(3.0 * sample + tt) * 0.7 + 0.3 * bias - 0.2 * pool * sample}


@torch.no_grad()
@pytest.mark.parametrize("phase", ["start", "plain", "fused", "last"])
def test_single_step_phases_match_reference(phase, monkeypatch):
    """One reference denoise_step per phase on the 50-step schedule: start step with 3 resampling iterations
    (:388-423), plain CFG step, fused multi-concept step (:376-386), final t == 1 step (:471-472)."""
    import oracle.sampler_ref as sr
    monkeypatch.setattr(sr, "register_time_ref", lambda *a, **k: None)
    gold = torch.load(os.path.join(G, "step_math_ref.pt"))
    cfg = RefConfig(n_timesteps=50, resolution_h=128, resolution_w=128, resampling_steps=3, jumping_steps=0)
    text, single = synth.make_text(CFG, K, gold["text_seed"])
    s = TweediemixRef(_ClosedFormUNet(), text, single, synth.fixture_masks(16, 16), cfg, K, lora=False, run_jump=False)
    case = gold["cases"][phase]
    got = s.denoise_step(gold["x"].clone(), case["t"])
    assert _rel(got, case["out"]) < RTOL
