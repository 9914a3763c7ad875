"""Concept-checkpoint formats either side of the hot path and prompt plumbing.

A single-concept training run of the reference writes ``delta-<steps>.bin`` with ``torch.save``:

    {'unet': {<parameter name>: tensor, ...},
     'modifier_token':   {'<new1>': tensor[768]},      # CLIP-L input embedding of the modifier token
     'modifier_token_2': {'<new1>': tensor[1280]}}     # OpenCLIP-bigG input embedding

* Custom-Diffusion checkpoints (``concept_training/diffusers_training_xl_new.py:41-66``, ``freeze_model ==
  'crossattn_kv'``) carry ``<block path>.attn2.to_k.weight`` / ``.attn2.to_v.weight`` ([d, 2048]) for the 70
  cross-attentions; ``'crossattn'`` checkpoints carry every ``attn2`` parameter.  The sampler copies whatever
  ``attn2`` parameters it finds into a full U-Net copy (``fusion_sampling.py:203-210``) and the hook then routes
  ONLY ``to_k`` / ``to_v`` per batch row (``utils_custom.py:64-82,125-133``) — so that is all this loader keeps.
* LoRA checkpoints (``concept_training/diffusers_training_xl_lora.py:43-73``) carry
  ``<attention path>.processor.to_{q,k,v,out}_lora.{down,up}.weight`` for all 140 attentions, rank 4
  (``model_lora.py:28-48,104-116``); the sampler copies them into ``create_lora_diffusion_base(unet_i)``
  (``fusion_sampling_lora.py:203-210``).

``custom_concept_from_delta`` / ``lora_concept_from_delta`` turn such a dict into the sparse donor module the hook
API reads (``model.unet_{i}``), without duplicating the 5 GB base U-Net per concept; ``save_*_delta`` write the same
format (used by the tests and to export synthetic concepts).  ``splice_modifier_prompts`` is the prompt construction
of ``fusion_sampling.py:139-154`` (quirk 15: ``str.find`` splice, first occurrence, silently mangles on a miss).
``components_from_diffusers`` is the reference constructor's loading path (``:119-210``); it needs diffusers,
transformers and local SDXL weights, none of which exist in the offline build image, so it is import-gated.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn as nn

from .model_lora import LoRAAttnProcessor_base
from .synthetic import SparseUNet, _KV, _Proc

LORA_LAYERS = ("to_q_lora", "to_k_lora", "to_v_lora", "to_out_lora")


# ------------------------------------------------------------------------------------------ file format
def load_delta(path: str) -> Dict[str, dict]:
    """``torch.load`` of a reference ``delta-*.bin`` with the structure checked (``fusion_sampling.py:157-158``)."""
    st = torch.load(path, map_location="cpu")
    if not isinstance(st, dict) or "unet" not in st or not isinstance(st["unet"], dict):
        raise ValueError(f"{path}: not a concept checkpoint (expected a dict with a 'unet' entry)")
    for k in ("modifier_token", "modifier_token_2"):
        st.setdefault(k, {})
    return st


def _save(path, unet_entries, modifier_token, modifier_token_2):
    torch.save({"unet": {k: v.detach().cpu().clone() for k, v in unet_entries.items()},
                "modifier_token": dict(modifier_token or {}), "modifier_token_2": dict(modifier_token_2 or {})}, path)


def save_custom_delta(path: str, donor: nn.Module, modifier_token=None, modifier_token_2=None):
    """Write a Custom-Diffusion checkpoint ('crossattn_kv' flavour) from a donor whose ``...attn2.to_k/to_v`` exist."""
    entries = {n: p for n, p in donor.named_parameters() if "attn2.to_k" in n or "attn2.to_v" in n}
    if not entries:
        raise ValueError("donor has no attn2.to_k / attn2.to_v parameters")
    _save(path, entries, modifier_token, modifier_token_2)


def save_lora_delta(path: str, donor: nn.Module, modifier_token=None, modifier_token_2=None):
    """Write a LoRA checkpoint from a donor whose attentions carry ``.processor.to_*_lora`` layers."""
    entries = {n: p for n, p in donor.named_parameters() if any(l in n for l in LORA_LAYERS)}
    if not entries:
        raise ValueError("donor has no to_{q,k,v,out}_lora parameters")
    _save(path, entries, modifier_token, modifier_token_2)


# ------------------------------------------------------------------------------------------ donors
def custom_concept_from_delta(unet, delta: Dict[str, dict], strict: bool = True) -> SparseUNet:
    """Sparse ``unet_i`` for the custom hook: ``<block>.attn2.to_k/to_v`` from the checkpoint, falling back to the base
    weight where the checkpoint has no entry (the reference copies only the names it finds, ``:206-209``).
    ``strict`` rejects checkpoints that name parameters the U-Net does not have or whose shapes differ."""
    entries = delta["unet"]
    used = set()
    donor = SparseUNet()
    for name, blk in unet.transformer_blocks():
        pair = []
        for which in ("to_k", "to_v"):
            base = getattr(blk.attn2, which)
            key = f"{name}.attn2.{which}.weight"
            w = entries.get(key)
            lin = nn.Linear(base.in_features, base.out_features, bias=False, device=base.weight.device, dtype=base.weight.dtype)
            if w is None:
                lin.weight.data.copy_(base.weight.data)
            else:
                if tuple(w.shape) != tuple(base.weight.shape):
                    raise ValueError(f"{key}: checkpoint shape {tuple(w.shape)} != U-Net shape {tuple(base.weight.shape)}")
                lin.weight.data.copy_(w.to(base.weight.dtype))
                used.add(key)
            pair.append(lin)
        donor.add_leaf(name + ".attn2", _KV(*pair))
    if strict:
        stray = [k for k in entries if k not in used and ("attn2.to_k" in k or "attn2.to_v" in k)]
        if stray:
            raise ValueError(f"checkpoint names {len(stray)} attn2 K/V parameters this U-Net does not have, e.g. {stray[0]}")
    if not used:
        raise ValueError("checkpoint holds no attn2.to_k / attn2.to_v weights (is it a LoRA checkpoint?)")
    return donor.requires_grad_(False)


def lora_concept_from_delta(unet, delta: Dict[str, dict], rank: int = 4) -> SparseUNet:
    """Sparse ``unet_i`` for the LoRA hook: ``<attention>.processor.to_{q,k,v,out}_lora`` on all attentions."""
    entries = delta["unet"]
    donor = SparseUNet()
    n_used = 0
    for name, attn in unet.attention_modules():
        hidden = attn.to_q.out_features
        proc = LoRAAttnProcessor_base(hidden, attn.to_k.in_features if attn.is_cross else None, rank)
        for layer_name in LORA_LAYERS:
            layer = getattr(proc, layer_name)
            for part in ("down", "up"):
                key = f"{name}.processor.{layer_name}.{part}.weight"
                if key not in entries:
                    raise ValueError(f"LoRA checkpoint is missing {key}")
                w = entries[key]
                dst = getattr(layer, part).weight
                if tuple(w.shape) != tuple(dst.shape):
                    raise ValueError(f"{key}: checkpoint shape {tuple(w.shape)} != expected {tuple(dst.shape)} (rank {rank})")
                dst.data.copy_(w)
                n_used += 1
        donor.add_leaf(name, _Proc(proc.to(attn.to_q.weight.device, attn.to_q.weight.dtype)))
    if n_used != len([k for k in entries if any(l in k for l in LORA_LAYERS)]):
        raise ValueError("LoRA checkpoint names attention modules this U-Net does not have")
    return donor.requires_grad_(False)


def concept_from_checkpoint(unet, path: str, variant: str) -> SparseUNet:
    delta = load_delta(path)
    return lora_concept_from_delta(unet, delta) if variant == "lora" else custom_concept_from_delta(unet, delta)


# ------------------------------------------------------------------------------------------ prompts
def splice_modifier_prompts(prompt_orig: str, prompt: str, concepts: str, modifier_token: str) -> Tuple[List[str], List[str]]:
    """``fusion_sampling.py:139-154``: ``+``-separated CLI strings -> (prompts, prompts_single).

    prompts        = [multi-concept prompt, concept prompt 1 with its modifier token spliced in, ...]  (K+1 entries)
    prompts_single = the first K-1 concept prompts WITHOUT modifier tokens (quirk 9)
    The modifier token goes in front of the FIRST occurrence of the concept word; if the word is absent ``str.find``
    returns -1 and the reference silently builds ``prompt[:-1] + token + ' ' + prompt[-1:]`` (quirk 15) — reproduced."""
    prompt_sep = prompt.split("+")
    concept_words = concepts.split("+")
    tokens = modifier_token.split("+")
    prompts = [prompt_orig.split("+")[0]]
    k = len(concept_words)
    prompts_single = prompt_sep[:k - 1]
    for i, wd in enumerate(concept_words):
        index = prompt_sep[i].find(wd)
        prompts.append(prompt_sep[i][:index] + tokens[i] + " " + prompt_sep[i][index:])
    return prompts, prompts_single


def modifier_embeddings(deltas: Sequence[Dict[str, dict]], modifier_tokens_user: Sequence[str]):
    """``fusion_sampling.py:160-188``: per user modifier token i, the (CLIP-L, bigG) input embeddings stored by checkpoint
    i under ITS OWN first token name (the reference indexes ``modifier_tokens[i]`` — the concatenated key lists — so
    checkpoint i must hold exactly one token for the pairing to be the intended one)."""
    names, names2 = [], []
    for st in deltas:
        names += list(st["modifier_token"].keys())
        names2 += list(st["modifier_token_2"].keys())
    out = []
    for i, tok in enumerate(modifier_tokens_user):
        out.append((tok, deltas[i]["modifier_token"][names[i]], deltas[i]["modifier_token_2"][names2[i]]))
    return out


# ------------------------------------------------------------------------------------------ diffusers-gated loader
def components_from_diffusers(config, variant: str):
    """The reference constructor's loading path (``fusion_sampling.py:119-223``): SDXL pipeline -> TmxUNet2DConditionModel
    (same state-dict keys), concept checkpoints -> sparse donors, modifier tokens -> tokenizer / text-encoder rows, prompts
    -> text embeddings (``get_text_embeds`` / ``encode_prompt`` ``:33-68,225-233``).  Needs diffusers + SDXL weights on disk."""
    import diffusers  # noqa: F401  (ImportError is turned into a RuntimeError by the caller)
    from diffusers import AutoencoderKL, StableDiffusionXLPipeline

    from .fusion_sampling import FusionComponents
    from .schedule import DDIMSchedule
    from .unet import TmxUNet2DConditionModel, UNetConfig

    dtype = torch.bfloat16 if getattr(config, "dtype", "bf16") == "bf16" else torch.float16
    model_key = getattr(config, "model_key", "stabilityai/stable-diffusion-xl-base-1.0")
    pipe = StableDiffusionXLPipeline.from_pretrained(model_key, torch_dtype=torch.float16, variant="fp16")
    vae = AutoencoderKL.from_pretrained("madebyollin/sdxl-vae-fp16-fix", torch_dtype=torch.float16).to("cuda")
    with torch.device("cuda"):
        unet = TmxUNet2DConditionModel(UNetConfig.sdxl_base()).to(dtype)
    missing, unexpected = unet.load_state_dict({k: v.to(dtype) for k, v in pipe.unet.state_dict().items()}, strict=False)
    if unexpected or [m for m in missing if "packed" not in m]:
        raise RuntimeError(f"SDXL U-Net state dict does not match: missing {missing[:3]}, unexpected {unexpected[:3]}")
    unet.requires_grad_(False).eval().finalize()

    deltas = [load_delta(p) for p in config.personal_checkpoint.split("+")]
    donors = [lora_concept_from_delta(unet, d) if variant == "lora" else custom_concept_from_delta(unet, d) for d in deltas]
    prompts, prompts_single = splice_modifier_prompts(config.prompt_orig, config.prompt, config.concepts, config.modifier_token)

    tokenizers, encoders = [pipe.tokenizer, pipe.tokenizer_2], [pipe.text_encoder.to("cuda"), pipe.text_encoder_2.to("cuda")]
    if deltas[0]["modifier_token"]:
        for tok, emb1, emb2 in modifier_embeddings(deltas, config.modifier_token.split("+")):
            for tk, enc, emb in ((tokenizers[0], encoders[0], emb1), (tokenizers[1], encoders[1], emb2)):
                tk.add_tokens(tok)
                enc.resize_token_embeddings(len(tk))
                enc.get_input_embeddings().weight.data[tk.convert_tokens_to_ids(tok)] = emb.to(enc.dtype)

    @torch.no_grad()
    def encode(texts):                                   # encode_prompt :33-68 — penultimate hidden states, pooled from encoder 2
        embeds, pooled = [], None
        for tk, enc in zip(tokenizers, encoders):
            ids = tk(texts, padding="max_length", max_length=tk.model_max_length, truncation=True, return_tensors="pt").input_ids
            out = enc(ids.to("cuda"), output_hidden_states=True)
            pooled = out[0]
            embeds.append(out.hidden_states[-2])
        return torch.cat(embeds, dim=-1).to(dtype), pooled.reshape(len(texts), -1).to(dtype)

    null = [config.negative_prompt]
    text = tuple(torch.cat(p) for p in zip(encode(null), encode(prompts)))                 # [uncond, multi, c_1..c_K]
    single = tuple(torch.cat(p) for p in zip(encode(null), encode(prompts_single)))       # [uncond, single_1..single_{K-1}]
    return FusionComponents(unet=unet, concept_unets=donors, text_embeds=text, text_embeds_single=single,
                            scheduler=DDIMSchedule(), masks=None, vae=vae)
