"""Concept-checkpoint formats: the reference's ``delta-*.bin`` layout
(writers concept_training/diffusers_training_xl_new.py:41-66 and ..._xl_lora.py:43-73, reader
fusion_sampling.py:139-210 / fusion_sampling_lora.py:203-210) read and written by tweediemix_b200.checkpoints,
and the prompt / modifier-token plumbing of fusion_sampling.py:139-190."""
import pytest
import torch

from oracle import synth as osynth
from oracle.unet_ref import UNetConfig as RefUNetConfig
from tweediemix_b200 import checkpoints as ck
from tweediemix_b200.synthetic import make_custom_concept, make_lora_concept
from tweediemix_b200.unet import TmxUNet2DConditionModel, UNetConfig


@pytest.fixture(scope="module")
def unet():
    torch.manual_seed(0)
    return TmxUNet2DConditionModel(UNetConfig.narrow()).requires_grad_(False)


def _reference_style_custom_delta(ref_unet, freeze_model="crossattn_kv"):
    """What the reference writer stores: a walk over ``unet.named_parameters()`` of the DIFFUSERS-shaped tree
    (the oracle stand-in has diffusers' module names) keeping the cross-attention K/V (or all attn2) weights."""
    d = {"unet": {}, "modifier_token": {"<cat1>": torch.randn(768)}, "modifier_token_2": {"<cat1>": torch.randn(1280)}}
    for name, p in ref_unet.named_parameters():
        if freeze_model == "crossattn" and "attn2" in name:
            d["unet"][name] = p.detach().clone()
        elif freeze_model == "crossattn_kv" and ("attn2.to_k" in name or "attn2.to_v" in name):
            d["unet"][name] = p.detach().clone()
    return d


def test_reads_reference_layout_written_from_diffusers_shaped_tree(tmp_path):
    """Key names produced by the reference writer on a diffusers-shaped U-Net address the same leaves in ours."""
    rcfg = RefUNetConfig.tiny()
    ref = osynth.make_concept_unet(osynth.make_base_unet(rcfg, 3), 9)
    ours = TmxUNet2DConditionModel(UNetConfig.narrow())        # same topology / widths as RefUNetConfig.tiny()
    for flavour in ("crossattn_kv", "crossattn"):
        path = tmp_path / f"delta_{flavour}.bin"
        torch.save(_reference_style_custom_delta(ref, flavour), path)
        delta = ck.load_delta(str(path))
        donor = ck.custom_concept_from_delta(ours, delta)
        n = 0
        for name, blk in ours.transformer_blocks():
            for which in ("to_k", "to_v"):
                want = delta["unet"][f"{name}.attn2.{which}.weight"]
                got = getattr(donor.get_submodule(name + ".attn2"), which).weight
                assert torch.equal(got, want.to(got.dtype))
                n += 1
        assert n == 2 * len(list(ours.transformer_blocks())) and n > 0


def test_custom_round_trip_fallback_and_errors(unet, tmp_path):
    donor = make_custom_concept(unet, seed=5)
    path = str(tmp_path / "delta-200.bin")
    ck.save_custom_delta(path, donor, {"<dog1>": torch.ones(768)}, {"<dog1>": torch.ones(1280)})
    st = ck.load_delta(path)
    assert set(st) == {"unet", "modifier_token", "modifier_token_2"}
    assert all(k.endswith(("attn2.to_k.weight", "attn2.to_v.weight")) for k in st["unet"])
    assert len(st["unet"]) == 2 * len(list(unet.transformer_blocks()))
    back = ck.custom_concept_from_delta(unet, st)
    for (n1, p1), (n2, p2) in zip(donor.named_parameters(), back.named_parameters()):
        assert n1 == n2 and torch.equal(p1, p2)
    # a checkpoint that lacks some entries: the base weight stands in (the reference copies only the names it finds)
    some = dict(list(st["unet"].items())[:4])
    part = ck.custom_concept_from_delta(unet, {"unet": some, "modifier_token": {}, "modifier_token_2": {}})
    name, blk = list(unet.transformer_blocks())[-1]
    assert torch.equal(part.get_submodule(name + ".attn2").to_k.weight, blk.attn2.to_k.weight)
    # stray names / wrong shapes / wrong kind of checkpoint are loud
    bad = dict(st["unet"]); bad["down_blocks.9.attentions.0.transformer_blocks.0.attn2.to_k.weight"] = torch.zeros(2, 2)
    with pytest.raises(ValueError, match="does not have"):
        ck.custom_concept_from_delta(unet, {"unet": bad})
    k0 = next(iter(st["unet"]))
    with pytest.raises(ValueError, match="shape"):
        ck.custom_concept_from_delta(unet, {"unet": {**st["unet"], k0: torch.zeros(3, 3)}})
    with pytest.raises(ValueError, match="no attn2"):
        ck.custom_concept_from_delta(unet, {"unet": {}})
    torch.save([1, 2, 3], str(tmp_path / "junk.bin"))
    with pytest.raises(ValueError, match="not a concept checkpoint"):
        ck.load_delta(str(tmp_path / "junk.bin"))


def test_lora_round_trip_and_key_names(unet, tmp_path):
    donor = make_lora_concept(unet, seed=7)
    path = str(tmp_path / "delta-1000.bin")
    ck.save_lora_delta(path, donor)
    st = ck.load_delta(path)
    n_attn = len(list(unet.attention_modules()))
    assert len(st["unet"]) == n_attn * 4 * 2                         # q,k,v,out x down,up on every attention (attn1 AND attn2)
    k = next(iter(st["unet"]))
    assert ".processor.to_" in k and k.endswith((".down.weight", ".up.weight")) and (".attn1." in k or ".attn2." in k)
    back = ck.lora_concept_from_delta(unet, st)
    for (n1, p1), (n2, p2) in zip(donor.named_parameters(), back.named_parameters()):
        assert n1 == n2 and torch.equal(p1, p2)
    # rank-4 layout of model_lora.py:28-48: down [4, in], up [out, 4]
    name, attn = next(iter(unet.attention_modules()))
    assert st["unet"][f"{name}.processor.to_q_lora.down.weight"].shape == (4, attn.to_q.in_features)
    assert st["unet"][f"{name}.processor.to_out_lora.up.weight"].shape == (attn.to_q.out_features, 4)
    missing = dict(st["unet"]); missing.pop(k)
    with pytest.raises(ValueError, match="missing"):
        ck.lora_concept_from_delta(unet, {"unet": missing})
    with pytest.raises(ValueError, match="no to_"):
        ck.save_lora_delta(path, make_custom_concept(unet, seed=1))


def test_prompt_splice_matches_reference_script_flags():
    """sample_catdog.sh:10-20 through fusion_sampling.py:139-154."""
    prompts, single = ck.splice_modifier_prompts(
        "photo of a cat and a dog running, mountain background",
        "photo of a cat running, mountain background+photo of a dog running, mountain background+mountain background",
        "cat+dog+mountain", "<cat1>+<dog1>+<mountain1>")
    assert prompts == ["photo of a cat and a dog running, mountain background",
                       "photo of a <cat1> cat running, mountain background",
                       "photo of a <dog1> dog running, mountain background",
                       "<mountain1> mountain background"]
    assert single == ["photo of a cat running, mountain background", "photo of a dog running, mountain background"]
    # quirk 15: concept word absent -> str.find == -1 -> the token lands before the LAST character, silently
    prompts, _ = ck.splice_modifier_prompts("x", "photo of a bird+sky", "cat+sky", "<c>+<s>")
    assert prompts[1] == "photo of a bir<c> d" and prompts[2] == "<s> sky"


def test_modifier_embedding_pairing():
    d = [{"modifier_token": {"<a>": torch.full((768,), float(i))}, "modifier_token_2": {"<a>": torch.full((1280,), 10.0 + i)}, "unet": {}}
         for i in range(3)]
    out = ck.modifier_embeddings(d, ["<cat1>", "<dog1>", "<mountain1>"])
    assert [t for t, _, _ in out] == ["<cat1>", "<dog1>", "<mountain1>"]
    assert [float(e1[0]) for _, e1, _ in out] == [0.0, 1.0, 2.0] and [float(e2[0]) for _, _, e2 in out] == [10.0, 11.0, 12.0]
