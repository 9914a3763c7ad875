"""The product hook layer applied to a DIFFUSERS-SHAPED U-Net (the tree the reference patches) instead of the product's
own ``TmxUNet2DConditionModel`` — INTEGRATION.md path B.

``register_attention_control_efficient`` / ``register_time`` of ``tweediemix_b200.utils_custom`` / ``utils_lora`` walk
``unet.{down,mid,up}_blocks[..].attentions[..].transformer_blocks[..].attn{1,2}`` like
``/root/reference/fusion_generation/utils_custom.py:113-158`` and ``utils_lora.py:134-217`` do, leave the same attributes
on the modules and assign instance-level ``forward`` closures that run the tmx kernels.  The outputs are held against the
golden vectors the REFERENCE's own unmodified hook files produced on the same seeded stand-in
(``tests/golden/make_golden.py`` -> ``hooks_{custom,lora}_{module,unet}.pt``, ``register_time.pt``).

CPU: kernels replaced by the fp32 stand-ins of ``fake_ops`` (wiring check, 5e-5).  ``-m gpu``: the real kernels through the
C ABI, fp16 weights under ``torch.autocast`` (the reference's execution mode, ``fusion_sampling.py:492``); tolerance
1.5e-2 of max|golden| (16-bit rounding through 70 transformer blocks, same bound as tests/test_gpu_model.py).
"""
import os
import types

import pytest
import torch

import fake_ops
from oracle import synth
from oracle.hooks_ref import make_lora_set
from oracle.unet_ref import UNetConfig, transformer_blocks_in_hook_order

CFG = UNetConfig.tiny()
K = 3
WINDOW = torch.tensor([781, 761, 741])


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def _holder(variant, device="cpu", dtype=torch.float32):
    """Same construction as tests/golden/make_golden.py: base U-Net + K donors in the REFERENCE's donor format."""
    base = synth.make_base_unet(CFG, 1234).to(device, dtype)
    if variant == "custom":
        donors = [synth.make_concept_unet(synth.make_base_unet(CFG, 1234), 100 + i).to(device, dtype) for i in range(K)]
    else:
        donors = []
        for s in [make_lora_set(base, 200 + i) for i in range(K)]:
            u = synth.make_base_unet(CFG, 1234)
            for name, blk in transformer_blocks_in_hook_order(u):
                for which in ("attn1", "attn2"):
                    layers = {k: v.to(device, dtype) for k, v in s[f"{name}.{which}"].items()}
                    getattr(blk, which).processor = types.SimpleNamespace(**layers)
            donors.append(u)
    m = types.SimpleNamespace(unet=base)
    for i, u in enumerate(donors):
        setattr(m, f"unet_{i}", u)
    return m


def _hooks(variant):
    from tweediemix_b200 import utils_custom, utils_lora
    return utils_lora if variant == "lora" else utils_custom


def _unet_eps(model, hooks, gd, t, device, dtype):
    (E, P), _ = synth.make_text(CFG, K, 77)
    ehs, pool = torch.cat([E[0:1], E[2:]]).to(device, dtype), torch.cat([P[0:1], P[2:]]).to(device, dtype)
    cond = {"time_ids": torch.tensor([[128, 128, 0, 0, 128, 128]]).repeat(4, 1).to(device), "text_embeds": pool}
    hooks.register_time(model, t)
    x = torch.cat([gd["latent"]] * 4).to(device, dtype)
    return model.unet(x, t, encoder_hidden_states=ehs, added_cond_kwargs=cond)["sample"]


# ----------------------------------------------------------------------------------------- CPU: wiring
@torch.no_grad()
@pytest.mark.parametrize("variant", ["custom", "lora"])
def test_product_hooks_patch_diffusers_shaped_unet(monkeypatch, golden_dir, variant):
    fake_ops.install(monkeypatch)
    hooks = _hooks(variant)
    model = _holder(variant)
    before = set(model.unet.state_dict())
    hooks.register_attention_control_efficient(model, WINDOW, K, gate=4)
    gd = _load(golden_dir, f"hooks_{variant}_unet.pt")
    torch.testing.assert_close(_unet_eps(model, hooks, gd, 761, "cpu", torch.float32), gd["eps_in_window_t761"], atol=5e-5, rtol=0)
    torch.testing.assert_close(_unet_eps(model, hooks, gd, 801, "cpu", torch.float32), gd["eps_outside_t801"], atol=5e-5, rtol=0)
    # the same modules are stamped / patched as by the reference, and no parameter was added to the tree
    stamped = sorted(n for n, m in model.unet.named_modules() if hasattr(m, "t"))
    assert stamped == sorted(_load(golden_dir, "register_time.pt")[variant])
    patched = [m for _, m in model.unet.named_modules() if "forward" in m.__dict__]
    assert len(patched) == (140 if variant == "lora" else 70)
    a2 = model.unet.down_blocks[1].attentions[0].transformer_blocks[1].attn2
    assert a2.num_concepts == K and a2.t_cond is WINDOW
    assert hasattr(a2, "to_k_2" if variant == "custom" else "to_out_2_lora")
    donor_keys = {k for k in model.unet.state_dict() if k not in before}
    assert all(("to_k_" in k or "to_v_" in k or "_lora" in k) for k in donor_keys)      # only the grafted donor leaves, as in the reference


@torch.no_grad()
def test_module_level_goldens_custom(monkeypatch, golden_dir):
    fake_ops.install(monkeypatch)
    from tweediemix_b200 import utils_custom
    model = _holder("custom")
    utils_custom.register_attention_control_efficient(model, WINDOW, K, gate=4)
    gd = _load(golden_dir, "hooks_custom_module.pt")
    mod = model.unet.down_blocks[1].attentions[0].transformer_blocks[1].attn2
    utils_custom.register_time(model, 781)
    torch.testing.assert_close(mod.forward(gd["x4"], encoder_hidden_states=gd["e4"]), gd["routed"], atol=2e-5, rtol=0)
    torch.testing.assert_close(mod.forward(gd["x5"], encoder_hidden_states=gd["e5"]), gd["batch5"], atol=2e-5, rtol=0)   # gate literal 4
    utils_custom.register_time(model, 801)
    torch.testing.assert_close(mod.forward(gd["x4"], encoder_hidden_states=gd["e4"]), gd["outside"], atol=2e-5, rtol=0)


# ----------------------------------------------------------------------------------------- GPU: real kernels
@pytest.mark.gpu
@torch.no_grad()
@pytest.mark.parametrize("variant", ["custom", "lora"])
def test_product_hooks_on_diffusers_shaped_unet_gpu(golden_dir, variant):
    from tweediemix_b200 import build, ops
    build.build()
    hooks = _hooks(variant)
    model = _holder(variant, "cuda", torch.float16)
    hooks.register_attention_control_efficient(model, WINDOW, K, gate=4)
    gd = _load(golden_dir, f"hooks_{variant}_unet.pt")
    n0 = ops.launch_count()
    with torch.autocast("cuda", dtype=torch.float16):
        got_in = _unet_eps(model, hooks, gd, 761, "cuda", torch.float16).float().cpu()
        got_out = _unet_eps(model, hooks, gd, 801, "cuda", torch.float16).float().cpu()
    assert ops.launch_count() - n0 >= 2 * (140 if variant == "lora" else 70)        # every patched attention ran a tmx kernel
    for got, want in ((got_in, gd["eps_in_window_t761"]), (got_out, gd["eps_outside_t801"])):
        rel = (got - want).abs().max().item() / want.abs().max().item()
        print(f"adapter {variant}: rel max|diff| vs reference-hook golden = {rel:.3e}")
        assert rel <= 1.5e-2
    assert (got_in - got_out).abs().max().item() > 1e-3                                # routing is observable


@pytest.mark.gpu
@torch.no_grad()
def test_module_level_goldens_gpu(golden_dir):
    from tweediemix_b200 import build, utils_custom, utils_lora
    build.build()
    for variant, hooks in (("custom", utils_custom), ("lora", utils_lora)):
        model = _holder(variant, "cuda", torch.float16)
        hooks.register_attention_control_efficient(model, WINDOW, K, gate=4)
        gd = _load(golden_dir, f"hooks_{variant}_module.pt")
        blk = model.unet.down_blocks[1].attentions[0].transformer_blocks[1]
        c = lambda t: t.to("cuda", torch.float16)
        hooks.register_time(model, 781)
        pairs = [(blk.attn2.forward(c(gd["x4"]), encoder_hidden_states=c(gd["e4"])), gd["routed" if variant == "custom" else "cross_routed"]),
                 (blk.attn2.forward(c(gd["x5"]), encoder_hidden_states=c(gd["e5"])), gd["batch5" if variant == "custom" else "cross_batch5"])]
        if variant == "lora":
            pairs.append((blk.attn1.forward(c(gd["x4"])), gd["self_routed"]))
        hooks.register_time(model, 801)
        pairs.append((blk.attn2.forward(c(gd["x4"]), encoder_hidden_states=c(gd["e4"])), gd["outside" if variant == "custom" else "cross_outside"]))
        for got, want in pairs:
            torch.testing.assert_close(got.float().cpu(), want, atol=4e-3 * want.abs().max().item() + 1e-3, rtol=0)
