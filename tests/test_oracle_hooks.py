"""Replay the golden vectors minted from the REFERENCE's own utils_custom.py / utils_lora.py
(tests/golden/make_golden.py) against the oracle restatement (oracle/hooks_ref.py).

Tolerance: the goldens were produced in fp32 on the build container's CPU; the restatement does
the same arithmetic in the same order, so agreement is to fp32 round-off.  ``ATOL`` leaves room
for a different BLAS kernel choice on another host CPU."""
import os

import pytest
import torch

from oracle import synth
from oracle.hooks_ref import (make_lora_set, register_custom_ref, register_lora_ref, register_time_ref)
from oracle.unet_ref import UNetConfig

ATOL = 2e-5
CFG = UNetConfig.tiny()
K = 3


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


@pytest.fixture(scope="module")
def custom_model():
    base = synth.make_base_unet(CFG, 1234)
    concepts = [synth.make_concept_unet(base, 100 + i) for i in range(K)]
    register_custom_ref(base, concepts, torch.tensor([781, 761, 741]), K)
    return base


@pytest.fixture(scope="module")
def lora_model():
    base = synth.make_base_unet(CFG, 1234)
    sets = [make_lora_set(base, 200 + i) for i in range(K)]
    register_lora_ref(base, sets, torch.tensor([781, 761, 741]), K)
    return base


@torch.no_grad()
def test_custom_module_matches_reference(custom_model, golden_dir):
    gd = _load(golden_dir, "hooks_custom_module.pt")
    mod = custom_model.down_blocks[1].attentions[0].transformer_blocks[1].attn2
    register_time_ref(custom_model, 781)
    torch.testing.assert_close(mod.forward(gd["x4"], encoder_hidden_states=gd["e4"]), gd["routed"], atol=ATOL, rtol=0)
    torch.testing.assert_close(mod.forward(gd["x5"], encoder_hidden_states=gd["e5"]), gd["batch5"], atol=ATOL, rtol=0)
    register_time_ref(custom_model, 801)
    torch.testing.assert_close(mod.forward(gd["x4"], encoder_hidden_states=gd["e4"]), gd["outside"], atol=ATOL, rtol=0)
    # the routing is observable: rows 1..3 differ in/out of window, row 0 does not
    d = (gd["routed"] - gd["outside"]).abs().amax(dim=(1, 2))
    assert d[0] == 0 and (d[1:] > 1e-3).all()


@torch.no_grad()
def test_lora_module_matches_reference(lora_model, golden_dir):
    gd = _load(golden_dir, "hooks_lora_module.pt")
    blk = lora_model.down_blocks[1].attentions[0].transformer_blocks[1]
    register_time_ref(lora_model, 781, lora=True)
    torch.testing.assert_close(blk.attn2.forward(gd["x4"], encoder_hidden_states=gd["e4"]), gd["cross_routed"], atol=ATOL, rtol=0)
    torch.testing.assert_close(blk.attn1.forward(gd["x4"]), gd["self_routed"], atol=ATOL, rtol=0)
    torch.testing.assert_close(blk.attn2.forward(gd["x5"], encoder_hidden_states=gd["e5"]), gd["cross_batch5"], atol=ATOL, rtol=0)
    register_time_ref(lora_model, 801, lora=True)
    torch.testing.assert_close(blk.attn2.forward(gd["x4"], encoder_hidden_states=gd["e4"]), gd["cross_outside"], atol=ATOL, rtol=0)
    torch.testing.assert_close(blk.attn1.forward(gd["x4"]), gd["self_outside"], atol=ATOL, rtol=0)
    for a, b in (("cross_routed", "cross_outside"), ("self_routed", "self_outside")):
        d = (gd[a] - gd[b]).abs().amax(dim=(1, 2))
        assert d[0] == 0 and (d[1:] > 1e-4).all()


def _unet_forward(model, gd, t, lora):
    (E, P), _ = synth.make_text(CFG, K, 77)
    ehs, pool = torch.cat([E[0:1], E[2:]]), torch.cat([P[0:1], P[2:]])
    cond = {"time_ids": torch.tensor([[128, 128, 0, 0, 128, 128]]).repeat(4, 1), "text_embeds": pool}
    register_time_ref(model, t, lora=lora)
    return model(torch.cat([gd["latent"]] * 4), t, encoder_hidden_states=ehs, added_cond_kwargs=cond)["sample"]


@torch.no_grad()
@pytest.mark.parametrize("variant", ["custom", "lora"])
def test_whole_unet_matches_reference(variant, custom_model, lora_model, golden_dir):
    gd = _load(golden_dir, f"hooks_{variant}_unet.pt")
    model = custom_model if variant == "custom" else lora_model
    got_in = _unet_forward(model, gd, 761, variant == "lora")
    got_out = _unet_forward(model, gd, 801, variant == "lora")
    torch.testing.assert_close(got_in, gd["eps_in_window_t761"], atol=5e-5, rtol=0)
    torch.testing.assert_close(got_out, gd["eps_outside_t801"], atol=5e-5, rtol=0)


def test_register_time_stamps_same_modules(custom_model, lora_model, golden_dir):
    gd = _load(golden_dir, "register_time.pt")
    register_time_ref(custom_model, 1)
    register_time_ref(lora_model, 1, lora=True)
    assert sorted(n for n, m in custom_model.named_modules() if hasattr(m, "t")) == sorted(gd["custom"])
    assert sorted(n for n, m in lora_model.named_modules() if hasattr(m, "t")) == sorted(gd["lora"])
    assert len(gd["custom"]) == 70 and len(gd["lora"]) == 140
