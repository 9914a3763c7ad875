"""CPU-only: HOST logic of the product (U-Net wiring, hook routing, sampler phases) against the oracle,
with the tmx kernels replaced by the plain-PyTorch stand-ins of ``tests/fake_ops.py``.  The same
comparisons run on the real kernels in ``tests/test_gpu_model.py`` (-m gpu)."""
import argparse
import copy

import pytest
import torch

import fake_ops
from oracle import synth
from oracle.hooks_ref import make_lora_set, register_custom_ref, register_lora_ref, register_time_ref
from oracle.sampler_ref import RefConfig, TweediemixRef
from oracle.unet_ref import UNetConfig as RefUNetConfig

K = 3
RCFG = RefUNetConfig.tiny()


def product_unet(ref_unet):
    from tweediemix_b200.unet import TmxUNet2DConditionModel, UNetConfig
    u = TmxUNet2DConditionModel(UNetConfig.narrow())
    own = set(u.state_dict())          # a hooked oracle U-Net also lists the grafted to_k_i / to_v_i donors
    u.load_state_dict({k: v for k, v in ref_unet.state_dict().items() if k in own})
    return u.eval().requires_grad_(False).finalize()


def _inputs(batch, hw=16, seed=5):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, 4, hw, hw, generator=g).repeat(batch, 1, 1, 1)
    E = torch.randn(batch, 77, RCFG.cross_attention_dim, generator=g)
    cond = {"text_embeds": torch.randn(batch, RCFG.pooled_embed_dim, generator=g),
            "time_ids": torch.tensor([[128, 128, 0, 0, 128, 128]]).repeat(batch, 1)}
    return x, E, cond


def test_state_dict_names_match_diffusers_shaped_oracle():
    ref = synth.make_base_unet(RCFG, 1)
    prod = product_unet(ref)
    assert set(prod.state_dict()) == set(ref.state_dict())
    assert len(list(prod.attention_modules())) == 140 and len(list(prod.transformer_blocks())) == 70


def test_unet_forward_matches_oracle(monkeypatch):
    fake_ops.install(monkeypatch)
    ref = synth.make_base_unet(RCFG, 1)
    prod = product_unet(ref)
    x, E, cond = _inputs(4)
    want = ref(x, 781, E, cond)["sample"]
    got = prod(x, 781, E, cond)["sample"]
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)
    # device-scalar timestep (the graph-capturable form) gives the same result
    got2 = prod(x, torch.tensor([781.0]), E, cond)["sample"]
    torch.testing.assert_close(got2, got, rtol=0, atol=0)


class _Holder:
    pass


@pytest.mark.parametrize("t,batch", [(781, 4), (981, 4), (761, 2)])
def test_custom_hooks_match_oracle(monkeypatch, t, batch):
    fake_ops.install(monkeypatch)
    from tweediemix_b200 import utils_custom
    ref = synth.make_base_unet(RCFG, 1)
    donors = [synth.make_concept_unet(ref, 10 + i) for i in range(K)]
    prod = product_unet(ref)
    window = torch.tensor([781, 761, 741])
    h = _Holder()
    h.unet = prod
    for i, d in enumerate(donors):
        setattr(h, f"unet_{i}", d)
    utils_custom.register_attention_control_efficient(h, window, K)
    utils_custom.register_time(h, t)
    register_custom_ref(ref, donors, window, K)
    register_time_ref(ref, t)
    x, E, cond = _inputs(batch)
    want = ref(x, t, E, cond)["sample"]
    got = prod(x, t, E, cond)["sample"]
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)
    a2 = prod.mid_block.attentions[0].transformer_blocks[0].attn2
    assert a2.num_concepts == K and a2.t == t and hasattr(a2, "to_k_2") and hasattr(a2, "to_v_0")


@pytest.mark.parametrize("t,batch", [(781, 4), (981, 4), (761, 2)])
def test_lora_hooks_match_oracle(monkeypatch, t, batch):
    fake_ops.install(monkeypatch)
    from tweediemix_b200 import utils_lora
    from tweediemix_b200.synthetic import SparseUNet
    ref = synth.make_base_unet(RCFG, 1)
    prod = product_unet(ref)
    lora_sets = [make_lora_set(ref, 20 + i, up_std=0.05) for i in range(K)]
    window = torch.tensor([781, 761, 741])
    h = _Holder()
    h.unet = prod
    for i, ls in enumerate(lora_sets):                       # donors expose <attention>.processor.to_*_lora
        donor = SparseUNet()
        for path, layers in ls.items():
            proc = torch.nn.Module()
            for nm, layer in layers.items():
                proc.add_module(nm, layer)
            holder = torch.nn.Module()
            holder.add_module("processor", proc)
            donor.add_leaf(path, holder)
        setattr(h, f"unet_{i}", donor)
    # the oracle's LoRA layers are plain down/up Linear pairs: give them the .pair() accessor the hook uses
    from tweediemix_b200.model_lora import LoRALinearLayer
    for i in range(K):
        for m in getattr(h, f"unet_{i}").modules():
            if hasattr(m, "down") and hasattr(m, "up"):
                m.pair = LoRALinearLayer.pair.__get__(m)
    utils_lora.register_attention_control_efficient(h, window, K)
    utils_lora.register_time(h, t)
    register_lora_ref(ref, lora_sets, window, K)
    register_time_ref(ref, t, lora=True)
    x, E, cond = _inputs(batch)
    want = ref(x, t, E, cond)["sample"]
    got = prod(x, t, E, cond)["sample"]
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)


def _namespace(n, res, lora):
    return argparse.Namespace(guidance_scale=0.8, n_timesteps=n, t_cond=0.2, t_stop=0.8 if lora else None,
                              resampling_steps=2, jumping_steps=5, resolution_h=res, resolution_w=res,
                              crops_coords_top_left_h=0, crops_coords_top_left_w=0, seed=3821, output_path=".",
                              seg_concepts="a+b", seg_gpu=0)


def _masks_for(k, h, w):
    """The reference's shipped mask pair (+ background) for K == 3, a stripe partition otherwise."""
    return synth.fixture_masks(h, w) if k == 3 else synth.stripe_masks(k, h, w)


def _product_sampler(ref_unet, donors_or_loras, lora, n, res, pg=None, k=K, **kw):
    from tweediemix_b200.fusion_sampling import FusionComponents, Tweediemix
    from tweediemix_b200.schedule import DDIMSchedule
    from tweediemix_b200.synthetic import SparseUNet
    prod = product_unet(ref_unet)
    text, single = synth.make_text(RCFG, k, 77)
    masks = _masks_for(k, res // 8, res // 8)
    if lora:
        from tweediemix_b200.model_lora import LoRALinearLayer
        donors = []
        for ls in donors_or_loras:
            donor = SparseUNet()
            for path, layers in ls.items():
                proc = torch.nn.Module()
                for nm, layer in layers.items():
                    layer = copy.deepcopy(layer)
                    layer.pair = LoRALinearLayer.pair.__get__(layer)
                    proc.add_module(nm, layer)
                holder = torch.nn.Module()
                holder.add_module("processor", proc)
                donor.add_leaf(path, holder)
            donors.append(donor)
    else:
        donors = donors_or_loras
    comp = FusionComponents(unet=prod, concept_unets=donors, text_embeds=text, text_embeds_single=single,
                            scheduler=DDIMSchedule(), masks=masks)
    return Tweediemix(_namespace(n, res, lora), comp, variant="lora" if lora else "custom",
                      use_cuda_graphs=False, process_group=pg, **kw)


def _oracle_sampler(ref_unet, donors_or_loras, lora, n, res, k=K):
    """The oracle sampler; for K != 3 the routing gate is generalised from the reference's literal 4 (utils_custom.py:61-62,
    utils_lora.py:63) to K + 1, which is what the product does by default."""
    text, single = synth.make_text(RCFG, k, 77)
    masks = _masks_for(k, res // 8, res // 8)
    cfg = RefConfig(n_timesteps=n, resolution_h=res, resolution_w=res, t_stop=0.8 if lora else None, resampling_steps=2)
    s = TweediemixRef(ref_unet, text, single, masks, cfg, k, lora=lora, run_jump=False)
    (register_lora_ref if lora else register_custom_ref)(ref_unet, donors_or_loras, s.hook_gate_window(), k, gate=k + 1)
    return s


@pytest.mark.parametrize("lora", [False, True])
def test_sampler_matches_oracle_config1(monkeypatch, lora):
    """BASELINE config 1 shape (K=3, fp32, CPU, few DDIM steps) end to end: identical schedule, phases,
    sample-forward count and latents (1e-3 max-abs, the north star's bound; fp32 lands far inside)."""
    fake_ops.install(monkeypatch)
    n, res = 10, 128
    ref_unet = synth.make_base_unet(RCFG, 1)
    extra = [make_lora_set(ref_unet, 20 + i, up_std=0.05) for i in range(K)] if lora else \
            [synth.make_concept_unet(ref_unet, 10 + i) for i in range(K)]
    prod = _product_sampler(ref_unet, extra, lora, n, res)
    orc = _oracle_sampler(ref_unet, extra, lora, n, res)
    x0 = orc.initial_latent()
    want = orc.sample_loop(x0.clone())
    prod.init_fusion(int(n * 0.2), int(n * 0.8)) if lora else prod.init_fusion(int(n * 0.2))
    assert (prod.t_cond_prev, prod.t_cond_cur, prod.start_t, prod.skip) == (orc.t_cond_prev, orc.t_cond_cur, orc.start_t, orc.skip)
    got = prod.sample_loop(x0.clone())
    assert prod.n_forward_rows == orc.n_forward_rows
    assert (got - want).abs().max().item() < 1e-3
    assert not hasattr(prod, "unet_0")                     # fusion_sampling.py:482-483


def test_schedule_matches_oracle():
    from tweediemix_b200.schedule import DDIMSchedule
    from oracle.schedule import make_schedule
    for n in (5, 50):
        s = DDIMSchedule()
        assert len(s.timesteps) == 1000
        s.set_timesteps(n)
        o = make_schedule(n)
        assert torch.equal(s.timesteps, o.timesteps)
        assert torch.equal(torch.cat([torch.tensor([1.0]), s.alphas_cumprod]), o.alphas_cumprod)
        assert torch.equal(s.final_alpha_cumprod, o.final_alpha_cumprod)


def test_masks_match_oracle(golden_dir):
    import os
    from tweediemix_b200.masks import load_region_masks, stripe_masks
    for which, names in [("test_out", "a cat+a dog"), ("test_out_panda", "a panda+a teddybear")]:
        d = os.path.join(golden_dir, "masks", which)
        got = load_region_masks(d, names, 128, 128)
        assert torch.equal(got, synth.fixture_masks(128, 128, which))
    assert torch.equal(stripe_masks(8, 16, 24), synth.stripe_masks(8, 16, 24))


def test_assign_rows():
    from tweediemix_b200.fusion_sampling import assign_rows
    assert [assign_rows(4, 2, r) for r in range(2)] == [[0, 1], [2, 3]]
    assert [assign_rows(4, 4, r) for r in range(4)] == [[0], [1], [2], [3]]
    assert [assign_rows(2, 4, r) for r in range(4)] == [[0], [1], [], []]
    assert [assign_rows(9, 4, r) for r in range(4)] == [[0, 1, 2], [3, 4], [5, 6], [7, 8]]      # balanced: no idle rank
    assert [len(assign_rows(36, 8, r)) for r in range(8)] == [5, 5, 5, 5, 4, 4, 4, 4]           # configs[3]: 9 rows x 4 images
    assert assign_rows(4, 1, 0) == [0, 1, 2, 3]
    from tweediemix_b200.fusion_sampling import assign_units
    assert assign_units(2, 4, 2, 1) == [(1, 0), (1, 1), (1, 2), (1, 3)]                         # image-major units
    assert assign_units(4, 9, 8, 0) == [(0, 0), (0, 1), (0, 2), (0, 3), (0, 4)]


@pytest.mark.parametrize("variant", ["custom", "lora"])
def test_product_host_path_matches_reference_run(monkeypatch, variant, golden_dir):
    """The PRODUCT's Tweediemix / hook layer / U-Net wiring (tmx kernels replaced by the plain-PyTorch stand-ins, fp32,
    CPU) against the latents the REFERENCE's own unmodified init_fusion + denoise_step produced step by step
    (tests/golden/make_golden_sampler.py): config 1 (256x256, 5 steps, custom) and the 10-step LoRA run with t_stop.
    Tolerance 1e-3 max-abs — the north star's end-to-end bound; fp32 lands at ~1e-5."""
    import argparse
    import os
    from tweediemix_b200.fusion_sampling import FusionComponents, Tweediemix
    from tweediemix_b200.schedule import DDIMSchedule
    fake_ops.install(monkeypatch)
    lora = variant == "lora"
    gold = torch.load(os.path.join(golden_dir, "sampler_lora_n10.pt" if lora else "sampler_custom_n5.pt"))
    m = gold["meta"]
    ref_unet = synth.make_base_unet(RCFG, m["base_seed"])
    extra = [make_lora_set(ref_unet, s) for s in m["lora_seeds"]] if lora else \
            [synth.make_concept_unet(ref_unet, s) for s in m["concept_seeds"]]
    # same construction as _product_sampler, with the golden run's text seed / step counts
    prod = _product_sampler(ref_unet, extra, lora, m["n"], m["res"])
    text, single = synth.make_text(RCFG, K, m["text_seed"])
    ns = argparse.Namespace(guidance_scale=m["guidance_scale"], n_timesteps=m["n"], t_cond=m["t_cond"], t_stop=m.get("t_stop"),
                            resampling_steps=m["resampling_steps"], jumping_steps=m["jumping_steps"], resolution_h=m["res"],
                            resolution_w=m["res"], crops_coords_top_left_h=0, crops_coords_top_left_w=0, seed=3821, output_path=".",
                            seg_concepts="a cat+a dog", seg_gpu=0)
    comp = FusionComponents(unet=prod.unet, concept_unets=[getattr(prod, f"unet_{i}") for i in range(K)], text_embeds=text,
                            text_embeds_single=single, scheduler=DDIMSchedule(), masks=gold["masks"].clone())
    model = Tweediemix(ns, comp, variant=variant, use_cuda_graphs=False)
    if lora:
        model.init_fusion(int(m["n"] * m["t_cond"]), int(m["n"] * m["t_stop"]))
    else:
        model.init_fusion(int(m["n"] * m["t_cond"]))
    got = []
    model.sample_loop(gold["x0"].clone(), callback=lambda i, t, x: got.append(x.clone()))
    assert len(got) == len(gold["xs"])
    for i, (a, b) in enumerate(zip(got, gold["xs"])):
        assert (a - b).abs().max().item() < 1e-3, f"step {i}: {(a - b).abs().max().item()}"
