"""CPU-only: the generalisations around the reference's single configuration (K = 3, one image) — host logic on the
plain-PyTorch kernel stand-ins of ``tests/fake_ops.py``; the same cases run on the real kernels in
``tests/test_gpu_parity.py`` (-m gpu).

  * K != 3: the routing gate is ``K + 1`` instead of the literal 4 (``utils_custom.py:61-62``, ``utils_lora.py:63``) —
    BASELINE configs[3] runs K = 8;
  * image batches: several latents share the prompt rows; every image must equal its own single-image run;
  * the mask hand-off (``fusion_sampling.py:431-469``) with a stub VAE: jump loop forward count, which x0 is decoded,
    fp16 VAE under an fp32 latent, ``tweedie.jpg`` on disk, masks picked up;
  * a scheduler object shared by two samplers is not mutated.
"""
import os

import pytest
import torch

import fake_ops
import test_host_logic as T
from oracle import synth
from oracle.hooks_ref import make_lora_set


def _extras(ref_unet, k, lora):
    return [make_lora_set(ref_unet, 20 + i, up_std=0.05) for i in range(k)] if lora else \
           [synth.make_concept_unet(ref_unet, 10 + i) for i in range(k)]


@pytest.mark.parametrize("k,lora", [(8, False), (2, True)])
def test_sampler_other_concept_counts_match_oracle(monkeypatch, k, lora):
    fake_ops.install(monkeypatch)
    n, res = 5, 64
    ref_unet = synth.make_base_unet(T.RCFG, 1)
    extra = _extras(ref_unet, k, lora)
    prod = T._product_sampler(ref_unet, extra, lora, n, res, k=k)
    orc = T._oracle_sampler(ref_unet, extra, lora, n, res, k=k)
    x0 = orc.initial_latent()
    want = orc.sample_loop(x0.clone())
    prod.init_fusion(int(n * 0.2), int(n * 0.8)) if lora else prod.init_fusion(int(n * 0.2))
    got = prod.sample_loop(x0.clone())
    assert prod.concept_num == k and prod.n_forward_rows == orc.n_forward_rows
    assert (got - want).abs().max().item() < 1e-3


@pytest.mark.parametrize("lora", [False, True])
def test_image_batch_equals_single_images(monkeypatch, lora):
    fake_ops.install(monkeypatch)
    n, res, imgs = 5, 128, 3
    ref_unet = synth.make_base_unet(T.RCFG, 1)
    extra = _extras(ref_unet, T.K, lora)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(imgs, 4, res // 8, res // 8, generator=g)
    s = T._product_sampler(ref_unet, extra, lora, n, res)
    s.init_fusion(int(n * 0.2), int(n * 0.8)) if lora else s.init_fusion(int(n * 0.2))
    one = [s.sample_loop(x[i:i + 1].clone()) for i in range(imgs)]
    rows_one = s.n_forward_rows
    s.n_forward_rows = 0
    many = s.sample_loop(x.clone())
    assert s.n_forward_rows == rows_one
    assert many.shape == x.shape
    for i in range(imgs):
        assert (many[i:i + 1] - one[i]).abs().max().item() < 1e-3      # fp32 GEMM blocking differs with the batch size


class _StubVAE(torch.nn.Module):
    """decode(z).sample = a fixed 1x1 conv of the x8-upsampled latent; fp16 weights like the reference's VAE."""

    class _Out:
        def __init__(self, sample):
            self.sample = sample

    def __init__(self, dtype):
        super().__init__()
        self.mix = torch.nn.Conv2d(4, 3, 1)
        torch.manual_seed(0)
        torch.nn.init.normal_(self.mix.weight, std=0.5)
        self.to(dtype)
        self.seen = []

    def decode(self, z):
        assert z.dtype == self.mix.weight.dtype, "latent must arrive in the VAE's dtype"
        self.seen.append(z.detach().float().clone())
        z = z.float()                                            # (CPU has no fp16 conv; the dtype contract is checked above)
        w, b = self.mix.weight.float(), self.mix.bias.float()
        return self._Out(torch.nn.functional.conv2d(torch.nn.functional.interpolate(z, scale_factor=8.0), w, b).to(self.mix.weight.dtype))


@pytest.mark.parametrize("jumping_steps", [5, 0])
def test_mask_handoff_with_stub_vae(monkeypatch, tmp_path, jumping_steps):
    fake_ops.install(monkeypatch)
    n, res = 10, 128
    ref_unet = synth.make_base_unet(T.RCFG, 1)
    extra = _extras(ref_unet, T.K, False)
    calls = []

    def provider(path, seg_concepts, h, w):
        from PIL import Image
        img = Image.open(path)
        assert img.size == (res, res) and img.mode == "RGB"      # tweedie.jpg: the decoded x0 at image resolution (:455)
        calls.append((path, seg_concepts))
        return synth.fixture_masks(h, w)

    s = T._product_sampler(ref_unet, extra, False, n, res, mask_provider=provider)
    s.masks = None                                               # no precomputed masks -> the hand-off must produce them
    s.vae = _StubVAE(torch.float16)
    s.config.output_path = str(tmp_path)
    s.config.jumping_steps = jumping_steps
    s.init_fusion(int(n * 0.2))
    orc = T._oracle_sampler(ref_unet, extra, False, n, res)
    orc.run_jump = True
    orc.config.jumping_steps = jumping_steps
    x0 = orc.initial_latent()
    want = orc.sample_loop(x0.clone())
    got = s.sample_loop(x0.clone())
    assert len(calls) == 1 and calls[0][0] == os.path.join(str(tmp_path), "tweedie.jpg") and calls[0][1] == "a+b"
    assert s.n_forward_rows == orc.n_forward_rows               # the jump's 2-row forwards are counted like the reference's
    assert (got - want).abs().max().item() < 1e-3               # the hand-off is output-neutral
    assert torch.equal(s.masks, synth.fixture_masks(res // 8, res // 8))
    # what was decoded: the jumped x0 / 0.18215 — with no jump, the Tweedie x0 of step t_cond_prev itself (:435,449-452)
    assert len(s.vae.seen) == 1
    if jumping_steps:
        want_x0 = orc.jumped_x0
    else:
        from oracle import step_math as sm
        xs = []
        orc2 = T._oracle_sampler(ref_unet, extra, False, n, res)
        orc2.sample_loop(x0.clone(), callback=lambda i, t, x: xs.append(x.clone()))
        i_prev = [int(t) for t in orc2.sched.timesteps].index(orc2.t_cond_prev)
        x_in = xs[i_prev - 1] if i_prev > 0 else x0
        E, P = orc2.text_embeds
        eps2 = orc2._unet(torch.cat([x_in, x_in]), orc2.t_cond_prev, E[:2], P[:2])
        _, want_x0 = sm.cfg_step(x_in, eps2, orc2.alpha(orc2.t_cond_prev), orc2.alpha(orc2.t_cond_prev - orc2.skip), 0.8)
    assert (s.vae.seen[0] - (want_x0 / 0.18215).half().float()).abs().max().item() < 2e-2


def test_handoff_needs_vae_and_single_process():
    ref_unet = synth.make_base_unet(T.RCFG, 1)
    s = T._product_sampler(ref_unet, _extras(ref_unet, T.K, False), False, 5, 128)
    s.masks = None
    with pytest.raises(RuntimeError, match="no VAE"):
        s._mask_handoff(torch.zeros(1, 4, 16, 16), 601, torch.zeros(1, 4, 16, 16))
    s.vae = _StubVAE(torch.float16)
    s.group_size = 2
    with pytest.raises(RuntimeError, match="single-process"):
        s._mask_handoff(torch.zeros(1, 4, 16, 16), 601, torch.zeros(1, 4, 16, 16))


def test_shared_scheduler_is_not_mutated(monkeypatch):
    fake_ops.install(monkeypatch)
    from tweediemix_b200.fusion_sampling import FusionComponents, Tweediemix
    from tweediemix_b200.schedule import DDIMSchedule
    ref_unet = synth.make_base_unet(T.RCFG, 1)
    a = T._product_sampler(ref_unet, _extras(ref_unet, T.K, False), False, 5, 128)
    sched = DDIMSchedule()
    comp = FusionComponents(unet=a.unet, concept_unets=[getattr(a, f"unet_{i}") for i in range(T.K)],
                            text_embeds=a.text_embeds, text_embeds_single=a.text_embeds_single, scheduler=sched, masks=a.masks)
    m1 = Tweediemix(T._namespace(5, 128, False), comp, use_cuda_graphs=False)
    m2 = Tweediemix(T._namespace(5, 128, False), comp, use_cuda_graphs=False)
    assert len(sched.timesteps) == 1000 and sched.alphas_cumprod.numel() == 1000
    assert m1.skip == m2.skip == 200 and m1._alpha_table == m2._alpha_table and len(m2._alpha_table) == 1001


@pytest.mark.parametrize("impl", ["cublas", "tmx"])
@pytest.mark.parametrize("lora", [False, True])
def test_gemm_policies_match_oracle(monkeypatch, lora, impl):
    """TMX_GEMM=cublas (library GEMMs + stand-alone GEGLU / add+LayerNorm / LoRA-delta kernels) and TMX_GEMM=tmx (every
    projection in k10, residual adds in its epilogue) are the same function as the default 'auto' mix (covered by
    test_host_logic.py): all match the oracle sampler."""
    fake_ops.install(monkeypatch)
    from tweediemix_b200 import ops
    monkeypatch.setattr(ops, "GEMM_IMPL", impl)
    n, res = 5, 64
    ref_unet = synth.make_base_unet(T.RCFG, 1)
    extra = _extras(ref_unet, T.K, lora)
    prod = T._product_sampler(ref_unet, extra, lora, n, res)
    orc = T._oracle_sampler(ref_unet, extra, lora, n, res)
    x0 = orc.initial_latent()
    want = orc.sample_loop(x0.clone())
    prod.init_fusion(int(n * 0.2), int(n * 0.8)) if lora else prod.init_fusion(int(n * 0.2))
    got = prod.sample_loop(x0.clone())
    assert (got - want).abs().max().item() < 1e-3


def test_cat_free_up_blocks_match_oracle(monkeypatch):
    """TMX_CAT_FREE=1: the up-block ResNets take (hidden, skip) as two sources (two-source GroupNorm, split 1x1 shortcut) instead of
    torch.cat — same function."""
    fake_ops.install(monkeypatch)
    from tweediemix_b200 import unet as U
    monkeypatch.setattr(U, "CAT_FREE", True)
    ref = synth.make_base_unet(T.RCFG, 1)
    prod = T.product_unet(ref)
    x, E, cond = T._inputs(4)
    want = ref(x, 781, E, cond)["sample"]
    calls = []
    real = U.ops.group_norm
    monkeypatch.setattr(U.ops, "group_norm", lambda *a, **k: (calls.append(k.get("x2") is not None), real(*a, **k))[1])
    got = prod(x, 781, E, cond)["sample"]
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)
    assert sum(calls) == 9                                  # the nine up-block ResNets
