"""Sampler restatement: phase structure, sample-forward budget and config 1
(256x256, 5 DDIM steps, K=3, fp32 CPU) end to end on the tiny stand-in."""
import pytest
import torch

from oracle import synth
from oracle.hooks_ref import make_lora_set, register_custom_ref, register_lora_ref
from oracle.sampler_ref import RefConfig, TweediemixRef
from oracle.unet_ref import UNetConfig

CFG = UNetConfig.tiny()
K = 3


class CountingUNet(torch.nn.Module):
    """Shape-only stand-in: returns a deterministic function of the input (cheap)."""

    def __init__(self):
        super().__init__()
        self.calls = []

    def forward(self, sample, t, encoder_hidden_states, added_cond_kwargs):
        self.calls.append((int(t), sample.shape[0]))
        bias = encoder_hidden_states.mean(dim=(1, 2)).reshape(-1, 1, 1, 1)
        return {"sample": 0.1 * sample + 0.01 * bias}


def _mk(unet, n=50, lora=False, run_jump=True, res=128, **kw):
    cfg = RefConfig(n_timesteps=n, resolution_h=res, resolution_w=res, t_stop=0.8 if lora else None, **kw)
    text, single = synth.make_text(CFG, K, 77)
    masks = synth.fixture_masks(res // 8, res // 8)
    return TweediemixRef(unet, text, single, masks, cfg, K, lora=lora, run_jump=run_jump)


def _patch_register_time(monkeypatch):
    import oracle.sampler_ref as sr
    monkeypatch.setattr(sr, "register_time_ref", lambda *a, **k: None)


def test_forward_budget_canonical(monkeypatch):
    _patch_register_time(monkeypatch)
    u = CountingUNet()
    s = _mk(u)
    s.sample_loop(s.initial_latent())
    assert s.n_forward_rows == 252                                  # 64 + 18 + 10 + 160
    assert (s.t_cond_prev, s.t_cond_cur, s.start_t) == (801, 781, 981)
    step0 = u.calls[:21]                                          # 1 + 10 x (2-row @961, 4-row @981)
    assert sum(b for _, b in step0) == 64 and all(t in (981, 961) for t, _ in step0)
    assert [t for t, b in u.calls if b == 2 and t < 781] == [781 - 150 * i for i in range(1, 5)]  # jump: 781,631..181
    u2 = CountingUNet()
    s2 = _mk(u2, run_jump=False)
    s2.sample_loop(s2.initial_latent())
    assert s2.n_forward_rows == 242


def test_lora_window_off_by_one(monkeypatch):
    _patch_register_time(monkeypatch)
    s = _mk(CountingUNet(), lora=True, run_jump=False)
    assert s.t_stop_cur == 181
    assert s.in_fused_phase(181) and 181 not in s.hook_gate_window()      # quirk 7
    assert s.in_fused_phase(201) and 201 in s.hook_gate_window()
    assert not s.in_fused_phase(161)
    s.sample_loop(s.initial_latent())
    # 64 + 9*2 + 31*4 (steps 10..40) + 9*2 (steps 41..49)
    assert s.n_forward_rows == 64 + 18 + 31 * 4 + 18


def test_resampling_zero_raises(monkeypatch):
    _patch_register_time(monkeypatch)
    s = _mk(CountingUNet(), resampling_steps=0)
    with pytest.raises(UnboundLocalError):
        s.denoise_step(s.initial_latent(), 981)


def test_seeded_latent_is_cpu_generator_draw():
    s = _mk(CountingUNet())
    torch.manual_seed(3821)
    want = torch.randn(1, 4, 16, 16)
    assert torch.equal(s.initial_latent(), want)


@torch.no_grad()
@pytest.mark.parametrize("variant", ["custom", "lora"])
def test_config1_five_steps_end_to_end(variant):
    """BASELINE config 1 on the tiny stand-in: runs all phases with the real hooks; output finite,
    deterministic and sensitive to the concept weights (so routing really happened)."""
    def run(delta_seed):
        base = synth.make_base_unet(CFG, 1234)
        n, lora = 5, variant == "lora"
        cfg = RefConfig(n_timesteps=n, resolution_h=256, resolution_w=256, resampling_steps=1,
                        t_stop=0.8 if lora else None)
        text, single = synth.make_text(CFG, K, 77)
        s = TweediemixRef(base, text, single, synth.fixture_masks(32, 32), cfg, K, lora=lora)
        window = torch.tensor(sorted(s.hook_gate_window(), reverse=True))
        if lora:
            register_lora_ref(base, [make_lora_set(base, delta_seed + i) for i in range(K)], window, K)
        else:
            register_custom_ref(base, [synth.make_concept_unet(base, delta_seed + i) for i in range(K)], window, K)
        return s.sample_loop(s.initial_latent())
    a, b, c = run(100), run(100), run(300)
    assert torch.isfinite(a).all() and a.shape == (1, 4, 32, 32)
    assert torch.equal(a, b)
    assert (a - c).abs().max() > 1e-4
