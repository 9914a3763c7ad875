#!/usr/bin/env python
"""Mint golden vectors by running the REFERENCE's own hook files in the build container.

    python tests/golden/make_golden.py            # needs /root/reference (build container only)

``fusion_generation/utils_custom.py`` and ``utils_lora.py`` import and execute unmodified here
(SURVEY §8c / App. D); the reference sampler itself does not (diffusers is absent).  This script
applies the reference's ``register_attention_control_efficient`` / ``register_time`` to the
diffusers-shaped stand-in (``oracle/unet_ref.py``) on seeded weights and inputs and stores the
reference's outputs as small ``.pt`` fixtures next to this file.  ``/root/reference`` does not
exist on the GPU box, so nothing at test time imports it — tests replay these files.

Fixtures written:
  hooks_custom_module.pt   one patched attn2 module: routed / out-of-window / batch-5 fall-back
  hooks_lora_module.pt     one patched attn1 + attn2 pair, LoRA variant, same three cases
  hooks_custom_unet.pt     whole tiny U-Net forward through the reference's custom hooks
  hooks_lora_unet.pt       ... through the reference's LoRA hooks
  register_time.pt         which modules the reference's register_time stamps (names)
"""
import importlib.util
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("TMX_REFERENCE", "/root/reference")

from oracle import synth  # noqa: E402
from oracle.hooks_ref import make_lora_set  # noqa: E402
from oracle.unet_ref import UNetConfig, transformer_blocks_in_hook_order  # noqa: E402


def load_ref(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, "fusion_generation", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


CFG = UNetConfig.tiny()
K = 3
BASE_SEED, TEXT_SEED = 1234, 77
WINDOW = torch.tensor([781, 761, 741])          # what init_fusion hands the hooks: a timestep tensor


def holder(base, concept_unets):
    m = types.SimpleNamespace(unet=base)
    for i, u in enumerate(concept_unets):
        setattr(m, f"unet_{i}", u)
    return m


def inputs(batch, n_tokens, gen):
    d = CFG.block_out_channels[1]
    x = torch.randn(batch, n_tokens, d, generator=gen)
    e = torch.randn(batch, 77, CFG.cross_attention_dim, generator=gen)
    return x, e


@torch.no_grad()
def main():
    torch.manual_seed(0)
    ref_custom, ref_lora = load_ref("utils_custom"), load_ref("utils_lora")

    # ---------------- custom variant ----------------
    base = synth.make_base_unet(CFG, BASE_SEED)
    concepts = [synth.make_concept_unet(base, 100 + i) for i in range(K)]
    model = holder(base, concepts)
    ref_custom.register_attention_control_efficient(model, WINDOW, K)
    g = torch.Generator().manual_seed(5)
    mod = base.down_blocks[1].attentions[0].transformer_blocks[1].attn2
    x4, e4 = inputs(4, 48, g)
    x5, e5 = inputs(5, 48, g)
    out = {"x4": x4, "e4": e4, "x5": x5, "e5": e5, "window": WINDOW}
    ref_custom.register_time(model, 781)
    out["routed"] = mod.forward(x4, encoder_hidden_states=e4)
    out["batch5"] = mod.forward(x5, encoder_hidden_states=e5)
    ref_custom.register_time(model, 801)
    out["outside"] = mod.forward(x4, encoder_hidden_states=e4)
    torch.save(out, os.path.join(HERE, "hooks_custom_module.pt"))

    stamped = [n for n, m in base.named_modules() if hasattr(m, "t")]
    torch.save({"custom": stamped}, os.path.join(HERE, "register_time.pt"))

    gl = torch.Generator().manual_seed(9)
    lat = torch.randn(1, 4, 16, 16, generator=gl)
    (E, P), _ = synth.make_text(CFG, K, TEXT_SEED)
    ehs, pool = torch.cat([E[0:1], E[2:]]), torch.cat([P[0:1], P[2:]])
    cond = {"time_ids": torch.tensor([[128, 128, 0, 0, 128, 128]]).repeat(4, 1), "text_embeds": pool}
    ref_custom.register_time(model, 761)
    eps_in = base(torch.cat([lat] * 4), 761, encoder_hidden_states=ehs, added_cond_kwargs=cond)["sample"]
    ref_custom.register_time(model, 801)
    eps_out = base(torch.cat([lat] * 4), 801, encoder_hidden_states=ehs, added_cond_kwargs=cond)["sample"]
    torch.save({"latent": lat, "eps_in_window_t761": eps_in, "eps_outside_t801": eps_out},
               os.path.join(HERE, "hooks_custom_unet.pt"))

    # ---------------- LoRA variant ----------------
    base = synth.make_base_unet(CFG, BASE_SEED)
    lora_sets = [make_lora_set(base, 200 + i) for i in range(K)]
    concepts = []
    for s in lora_sets:                      # reference reads unet_i....attn{1,2}.processor.to_*_lora
        u = synth.make_base_unet(CFG, BASE_SEED)
        for name, blk in transformer_blocks_in_hook_order(u):
            for which in ("attn1", "attn2"):
                getattr(blk, which).processor = types.SimpleNamespace(**s[f"{name}.{which}"])
        concepts.append(u)
    model = holder(base, concepts)
    ref_lora.register_attention_control_efficient(model, WINDOW, K)
    blk = base.down_blocks[1].attentions[0].transformer_blocks[1]
    g = torch.Generator().manual_seed(6)
    x4, e4 = inputs(4, 48, g)
    x5, e5 = inputs(5, 48, g)
    out = {"x4": x4, "e4": e4, "x5": x5, "e5": e5, "window": WINDOW}
    ref_lora.register_time(model, 781)
    out["cross_routed"] = blk.attn2.forward(x4, encoder_hidden_states=e4)
    out["self_routed"] = blk.attn1.forward(x4)
    out["cross_batch5"] = blk.attn2.forward(x5, encoder_hidden_states=e5)
    ref_lora.register_time(model, 801)
    out["cross_outside"] = blk.attn2.forward(x4, encoder_hidden_states=e4)
    out["self_outside"] = blk.attn1.forward(x4)
    torch.save(out, os.path.join(HERE, "hooks_lora_module.pt"))

    d = torch.load(os.path.join(HERE, "register_time.pt"))
    d["lora"] = [n for n, m in base.named_modules() if hasattr(m, "t")]
    torch.save(d, os.path.join(HERE, "register_time.pt"))

    ref_lora.register_time(model, 761)
    eps_in = base(torch.cat([lat] * 4), 761, encoder_hidden_states=ehs, added_cond_kwargs=cond)["sample"]
    ref_lora.register_time(model, 801)
    eps_out = base(torch.cat([lat] * 4), 801, encoder_hidden_states=ehs, added_cond_kwargs=cond)["sample"]
    torch.save({"latent": lat, "eps_in_window_t761": eps_in, "eps_outside_t801": eps_out},
               os.path.join(HERE, "hooks_lora_unet.pt"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
