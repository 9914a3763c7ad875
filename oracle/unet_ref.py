"""SDXL ``UNet2DConditionModel`` stand-in in plain PyTorch (oracle; test infrastructure only).

[D] Restates diffusers==0.29.2 (pinned at the reference's ``requirements.txt:4``;
not vendored there and not installed in this image) for the one configuration the
reference drives — ``stabilityai/stable-diffusion-xl-base-1.0`` ``unet/config.json``
(SURVEY App. A): module tree, state-dict key names and forward arithmetic of
``models/unets/unet_2d_condition.py``, ``unets/unet_2d_blocks.py``, ``models/resnet.py``
(body mirrored in the reference at ``video_gen/utils_attn.py:391-431``),
``models/transformers/transformer_2d.py``, ``models/attention.py``,
``models/attention_processor.py`` and ``models/embeddings.py``.

The tree is name-compatible with what the reference's hook files walk
(``fusion_generation/utils_custom.py:113-157``): ``unet.{down_blocks[1,2],mid_block,
up_blocks[0,1]}.attentions[j].transformer_blocks[k].attn{1,2}`` with transformer depths
2/10, so those files patch this stand-in unmodified (``tests/golden/make_golden.py``).

Widths are configurable so the same topology runs at toy size on CPU; ``UNetConfig.sdxl()``
is the real thing (2.57 B parameters).  The call surface is the one the reference uses at
``fusion_sampling.py:340``: ``unet(sample, t, encoder_hidden_states=E,
added_cond_kwargs={"text_embeds":…, "time_ids":…})['sample']``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280)
    layers_per_block: int = 2
    transformer_layers_per_block: Tuple[int, ...] = (1, 2, 10)
    num_heads: Tuple[int, ...] = (5, 10, 20)          # config key "attention_head_dim" (= #heads)
    cross_attention_dim: int = 2048
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    addition_time_embed_dim: int = 256
    pooled_embed_dim: int = 1280                        # text_embeds width
    down_has_attn: Tuple[bool, ...] = (False, True, True)
    up_has_attn: Tuple[bool, ...] = (True, True, False)

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4

    @property
    def projection_class_embeddings_input_dim(self) -> int:
        return self.pooled_embed_dim + 6 * self.addition_time_embed_dim   # 2816 for SDXL

    @staticmethod
    def sdxl() -> "UNetConfig":
        return UNetConfig()

    @staticmethod
    def tiny(width: int = 64, cross_dim: int = 128, pooled: int = 64, add_dim: int = 32) -> "UNetConfig":
        """Same topology (depths 1/2/10, 17 resnets, 70 transformer blocks), head_dim 64, narrow."""
        return UNetConfig(block_out_channels=(width, 2 * width, 4 * width),
                          num_heads=(width // 64, 2 * width // 64, 4 * width // 64),
                          cross_attention_dim=cross_dim, pooled_embed_dim=pooled,
                          addition_time_embed_dim=add_dim)


# ------------------------------------------------------------------ embeddings [D] models/embeddings.py

def sinusoidal_embedding(timesteps: torch.Tensor, dim: int) -> torch.Tensor:
    """``get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0)`` -> [cos | sin]."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
    arg = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim: int, out_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, out_dim)
        self.linear_2 = nn.Linear(out_dim, out_dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


# ------------------------------------------------------------------ attention [D] models/attention_processor.py

class Attention(nn.Module):
    """diffusers ``Attention`` surface the hooks rely on: to_q/to_k/to_v/to_out, heads, scale,
    head_to_batch_dim / batch_to_head_dim (``utils_custom.py:64-106``)."""

    def __init__(self, query_dim: int, heads: int, dim_head: int = 64, cross_attention_dim: int | None = None):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_v = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(0.0)])

    def head_to_batch_dim(self, t):
        b, n, c = t.shape
        return t.reshape(b, n, self.heads, c // self.heads).permute(0, 2, 1, 3).reshape(b * self.heads, n, c // self.heads)

    def batch_to_head_dim(self, t):
        bh, n, d = t.shape
        b = bh // self.heads
        return t.reshape(b, self.heads, n, d).permute(0, 2, 1, 3).reshape(b, n, self.heads * d)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None):
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q = self.head_to_batch_dim(self.to_q(hidden_states))
        k = self.head_to_batch_dim(self.to_k(ctx))
        v = self.head_to_batch_dim(self.to_v(ctx))
        probs = (torch.bmm(q, k.transpose(1, 2)) * self.scale).softmax(dim=-1)
        out = self.batch_to_head_dim(torch.bmm(probs.to(v.dtype), v))
        return self.to_out[1](self.to_out[0](out))


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, cross_attention_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, heads, 64, None)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, heads, 64, cross_attention_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, h, encoder_hidden_states):
        h = self.attn1(self.norm1(h), encoder_hidden_states=None, attention_mask=None) + h
        h = self.attn2(self.norm2(h), encoder_hidden_states=encoder_hidden_states, attention_mask=None) + h
        return self.ff(self.norm3(h)) + h


class Transformer2DModel(nn.Module):
    """use_linear_projection=True variant."""

    def __init__(self, channels: int, heads: int, depth: int, cross_attention_dim: int, groups: int):
        super().__init__()
        self.norm = nn.GroupNorm(groups, channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(channels, channels)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(channels, heads, cross_attention_dim) for _ in range(depth)])
        self.proj_out = nn.Linear(channels, channels)

    def forward(self, x, encoder_hidden_states):
        b, c, hh, ww = x.shape
        res = x
        h = self.norm(x).permute(0, 2, 3, 1).reshape(b, hh * ww, c)
        h = self.proj_in(h)
        for blk in self.transformer_blocks:
            h = blk(h, encoder_hidden_states)
        h = self.proj_out(h)
        return h.reshape(b, hh, ww, c).permute(0, 3, 1, 2) + res


# ------------------------------------------------------------------ resnet [D] models/resnet.py

class ResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, temb_dim: int, groups: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps, affine=True)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None
        self.output_scale_factor = 1.0

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(self.dropout(F.silu(self.norm2(h))))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return (x + h) / self.output_scale_factor


class Downsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


# ------------------------------------------------------------------ blocks [D] unets/unet_2d_blocks.py

class DownBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, cin, cout, depth, heads, has_attn, add_down):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(cin if i == 0 else cout, cout, cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps)
            for i in range(cfg.layers_per_block)])
        self.attentions = nn.ModuleList([
            Transformer2DModel(cout, heads, depth, cfg.cross_attention_dim, cfg.norm_num_groups)
            for _ in range(cfg.layers_per_block)]) if has_attn else None
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, h, temb, ehs):
        outs = []
        for i, res in enumerate(self.resnets):
            h = res(h, temb)
            if self.attentions is not None:
                h = self.attentions[i](h, ehs)
            outs.append(h)
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
            outs.append(h)
        return h, outs


class MidBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, c, depth, heads):
        super().__init__()
        mk = lambda: ResnetBlock2D(c, c, cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps)
        self.resnets = nn.ModuleList([mk(), mk()])
        self.attentions = nn.ModuleList([Transformer2DModel(c, heads, depth, cfg.cross_attention_dim, cfg.norm_num_groups)])

    def forward(self, h, temb, ehs):
        h = self.resnets[0](h, temb)
        h = self.attentions[0](h, ehs)
        return self.resnets[1](h, temb)


class UpBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, cin, cout, cprev, depth, heads, has_attn, add_up):
        super().__init__()
        n = cfg.layers_per_block + 1
        res = []
        for i in range(n):
            skip = cin if i == n - 1 else cout
            rin = cprev if i == 0 else cout
            res.append(ResnetBlock2D(rin + skip, cout, cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps))
        self.resnets = nn.ModuleList(res)
        self.attentions = nn.ModuleList([
            Transformer2DModel(cout, heads, depth, cfg.cross_attention_dim, cfg.norm_num_groups)
            for _ in range(n)]) if has_attn else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, h, skips, temb, ehs):
        for i, res in enumerate(self.resnets):
            h = res(torch.cat([h, skips.pop()], dim=1), temb)
            if self.attentions is not None:
                h = self.attentions[i](h, ehs)
        if self.upsamplers is not None:
            h = self.upsamplers[0](h)
        return h


# ------------------------------------------------------------------ the U-Net [D] unets/unet_2d_condition.py

class UNet2DConditionModelRef(nn.Module):
    def __init__(self, cfg: UNetConfig | None = None):
        super().__init__()
        cfg = cfg or UNetConfig.sdxl()
        self.cfg = cfg
        boc = cfg.block_out_channels
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], cfg.time_embed_dim)
        self.add_embedding = TimestepEmbedding(cfg.projection_class_embeddings_input_dim, cfg.time_embed_dim)

        downs, cout = [], boc[0]
        for i, c in enumerate(boc):
            cin, cout = cout, c
            downs.append(DownBlock(cfg, cin, cout, cfg.transformer_layers_per_block[i], cfg.num_heads[i],
                                   cfg.down_has_attn[i], add_down=(i != len(boc) - 1)))
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = MidBlock(cfg, boc[-1], cfg.transformer_layers_per_block[-1], cfg.num_heads[-1])

        rev = list(reversed(boc))
        rev_depth = list(reversed(cfg.transformer_layers_per_block))
        rev_heads = list(reversed(cfg.num_heads))
        ups, cout = [], rev[0]
        for i, c in enumerate(rev):
            cprev, cout = cout, c
            cin = rev[min(i + 1, len(boc) - 1)]
            ups.append(UpBlock(cfg, cin, cout, cprev, rev_depth[i], rev_heads[i], cfg.up_has_attn[i],
                               add_up=(i != len(boc) - 1)))
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, boc[0], eps=cfg.norm_eps)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    def embed(self, sample, timestep, added_cond_kwargs):
        b = sample.shape[0]
        t = torch.as_tensor(timestep, device=sample.device).reshape(-1).expand(b)
        dt = self.conv_in.weight.dtype
        emb = self.time_embedding(sinusoidal_embedding(t, self.cfg.block_out_channels[0]).to(dt))
        time_ids = added_cond_kwargs["time_ids"]
        tid = sinusoidal_embedding(time_ids.flatten(), self.cfg.addition_time_embed_dim).reshape(b, -1)
        add = torch.cat([added_cond_kwargs["text_embeds"].to(dt), tid.to(dt)], dim=-1)
        return emb + self.add_embedding(add)

    def forward(self, sample, timestep, encoder_hidden_states, added_cond_kwargs):
        dt = self.conv_in.weight.dtype
        temb = self.embed(sample, timestep, added_cond_kwargs)
        ehs = encoder_hidden_states.to(dt)
        h = self.conv_in(sample.to(dt))
        skips = [h]
        for blk in self.down_blocks:
            h, outs = blk(h, temb, ehs)
            skips += outs
        h = self.mid_block(h, temb, ehs)
        for blk in self.up_blocks:
            h = blk(h, skips, temb, ehs)
        h = self.conv_out(F.silu(self.conv_norm_out(h)))
        return {"sample": h}


def transformer_blocks_in_hook_order(unet):
    """The 70 BasicTransformerBlocks, in the order the reference's hook files visit them
    (up, then down, then mid: ``utils_custom.py:120-157``).  Yields (name, block)."""
    for res in (0, 1):
        for j, tr in enumerate(unet.up_blocks[res].attentions):
            for k, blk in enumerate(tr.transformer_blocks):
                yield f"up_blocks.{res}.attentions.{j}.transformer_blocks.{k}", blk
    for res in (1, 2):
        for j, tr in enumerate(unet.down_blocks[res].attentions):
            for k, blk in enumerate(tr.transformer_blocks):
                yield f"down_blocks.{res}.attentions.{j}.transformer_blocks.{k}", blk
    for k, blk in enumerate(unet.mid_block.attentions[0].transformer_blocks):
        yield f"mid_block.attentions.0.transformer_blocks.{k}", blk


def seeded_init_(unet: nn.Module, seed: int, branch_damp: float = 0.3) -> nn.Module:
    """Deterministic random init that keeps activations O(1) through 70 transformer blocks
    (no SDXL checkpoint is reachable offline — SURVEY §7 "hard parts").  Fan-in scaled normals;
    the output projection of every residual branch (attn ``to_out.0``, ``ff.net.2``, ``conv2``,
    ``proj_out``) is damped so the residual stream does not blow up under 50 steps."""
    g = torch.Generator().manual_seed(seed)
    for name, p in sorted(unet.named_parameters(), key=lambda kv: kv[0]):
        with torch.no_grad():
            if p.ndim == 1:
                if name.endswith("bias"):
                    p.copy_(torch.randn(p.shape, generator=g) * 0.02)
                else:                                   # norm weights
                    p.copy_(1.0 + torch.randn(p.shape, generator=g) * 0.05)
                continue
            fan_in = p[0].numel()
            std = fan_in ** -0.5
            if any(name.endswith(s) for s in ("to_out.0.weight", "ff.net.2.weight", "conv2.weight", "proj_out.weight")):
                std *= branch_damp
            p.copy_((torch.randn(p.shape, generator=g) * std).to(p.dtype))
    return unet
