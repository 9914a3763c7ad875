"""``Tweediemix`` sampler + CLI — drop-in for ``fusion_generation/fusion_sampling.py`` (and, with
``variant="lora"``, ``fusion_sampling_lora.py``) on the B200 kernels.

Same public surface as the reference class (``fusion_sampling.py:97-530``): ``Tweediemix(config)``,
``.alpha(t)``, ``.denoise_step(x, t)``, ``.init_fusion(t_cond[, t_stop])``, ``.run_fusion()``,
``.sample_loop(x)``, attributes ``.unet .unet_{i} .scheduler .masks .text_embeds
.text_embeds_single .concept_num .skip .t_cond .t_cond_prev .t_cond_cur .start_t .add_time_ids``,
and the same 19(+1) argparse flags.  What differs is *how* a step executes:

  * every phase of ``denoise_step`` (fused ``:376-386``, start-step resampling ``:388-423``, plain CFG
    ``:424-430``, last step ``:471-472``) ends in ONE launch of the fused k7 kernel
    (``tmx_tweedie_blend_ddim_fwd``) instead of ~9 eager elementwise kernels per concept;
  * ``alpha(t)`` values are host floats (the reference indexes a CPU table with a CUDA scalar and
    ``.item()``s ``t`` — >= 73 device syncs per step); the fusion window is a host set;
  * the U-Net forward of each (phase, routed) combination is captured once in a CUDA graph and
    replayed; the timestep lives in a device scalar so one graph serves all steps of a phase;
  * cross-attention K/V are projected once per prompt set (see ``unet.py``);
  * concept-parallel multi-GPU (``process_group``): the K+1 batch rows are block-distributed over the
    ranks of the group, each rank runs the U-Net on its rows only, reduces its masked partial with
    ``tmx_blend_partial_fwd``, ONE all-reduce of a 512 KiB fp32 buffer per step crosses NVLink, and
    every rank finishes with ``tmx_blend_finish_fwd`` — bit-identical latents on all ranks, no
    broadcast.

Text encoders, VAE and the GroundingDINO/SAM segmentation subprocess sit outside the hot path
(SURVEY §8 scope): the sampler takes text embeddings and (optionally) precomputed region masks as
inputs via ``FusionComponents``; with diffusers installed ``load_components_from_diffusers`` builds
them the way the reference constructor does.
"""
from __future__ import annotations

import argparse
import copy
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops
from . import utils_custom, utils_lora
from .masks import load_region_masks
from .schedule import DDIMSchedule

opt = None          # module-global parsed flags, like the reference (read by compute_time_ids)


def save_image(img: torch.Tensor, path: str) -> None:
    """[3,H,W] in [0,1] -> 8-bit RGB file (what ``T.ToPILImage()(img).save(path)`` writes, ``fusion_sampling.py:455,522``)."""
    import numpy as np
    from PIL import Image
    arr = (img.detach().float().cpu().clamp(0, 1) * 255.0).to(torch.uint8).permute(1, 2, 0).numpy()   # ToPILImage: mul(255).byte()
    Image.fromarray(np.ascontiguousarray(arr)).save(path)


def compute_time_ids(config=None):
    """``fusion_sampling.py:70-78``: [[H, W, crop_top, crop_left, H, W]] int64."""
    c = config if config is not None else opt
    return torch.tensor([[c.resolution_h, c.resolution_w, c.crops_coords_top_left_h, c.crops_coords_top_left_w,
                          c.resolution_h, c.resolution_w]])


@dataclass
class FusionComponents:
    unet: nn.Module
    concept_unets: Sequence[nn.Module]                     # become model.unet_{i}
    text_embeds: Tuple[torch.Tensor, torch.Tensor]         # ([K+2,77,D], [K+2,P]) = [uncond, multi, c_1..c_K]
    text_embeds_single: Tuple[torch.Tensor, torch.Tensor]  # ([K,77,D], [K,P])     = [uncond, single_1..single_{K-1}]
    scheduler: DDIMSchedule = field(default_factory=DDIMSchedule)
    masks: Optional[torch.Tensor] = None                   # precomputed [K,1,h,w] fp32 (north star: masks are inputs)
    vae: Optional[nn.Module] = None


class _RowSet:
    """Static device buffers of one phase.  A *unit* is one (image, prompt row) pair = one batch row of the U-Net; this
    rank owns the units ``local_units`` (image-major, so the rows of one image are contiguous in its eps tensor)."""

    def __init__(self, name, ehs, pooled, time_ids, local_units, n_rows, n_imgs, latent_shape, dtype, device):
        self.name, self.n_rows, self.n_imgs = name, n_rows, n_imgs
        self.local_units = list(local_units)
        self.local_ids = [r for _, r in self.local_units]            # prompt row of every local unit
        self.local_imgs = [i for i, _ in self.local_units]
        n = len(self.local_units)
        self.ehs = ehs[self.local_ids].to(device=device, dtype=dtype).contiguous() if n else None
        self.pooled = pooled[self.local_ids].to(device=device, dtype=dtype).contiguous() if n else None
        self.time_ids = time_ids.repeat(n, 1).to(device) if n else None
        self.latent = torch.empty((n,) + tuple(latent_shape[1:]), dtype=dtype, device=device) if n else None
        self.img_index = torch.tensor(self.local_imgs, dtype=torch.long, device=device) if n and n_imgs > 1 else None
        self.graphs = {}          # routed flag -> (CUDAGraph, eps tensor, launches per replay)

    def cond(self):
        return {"time_ids": self.time_ids, "text_embeds": self.pooled}

    def rows_of_image(self, img: int):
        """(first local index, prompt rows) of this rank's units of image ``img`` (contiguous by construction)."""
        idx = [j for j, i in enumerate(self.local_imgs) if i == img]
        return (idx[0] if idx else 0), [self.local_ids[j] for j in idx]


def assign_rows(n_rows: int, group_size: int, rank: int) -> List[int]:
    """Balanced contiguous distribution of ``n_rows`` work units over the ranks of a concept-parallel group: the first
    ``n_rows % group_size`` ranks own one unit more; with fewer units than ranks the trailing ranks own nothing."""
    q, rem = divmod(n_rows, group_size)
    lo = rank * q + min(rank, rem)
    return list(range(lo, lo + q + (1 if rank < rem else 0)))


def assign_units(n_imgs: int, n_rows: int, group_size: int, rank: int) -> List[Tuple[int, int]]:
    """(image, prompt row) units of this rank: the ``n_imgs * n_rows`` units, image-major, split by ``assign_rows``."""
    return [divmod(u, n_rows) for u in assign_rows(n_imgs * n_rows, group_size, rank)]


def make_concept_groups(world: int, rank: int, n_rows: int, max_group: Optional[int] = None):
    """Concept-parallel process groups.  The ``n_rows`` = K+1 batch rows of one image are the only work that
    shards, so ranks form ``n_groups = world // G`` groups of ``G = min(world, n_rows[, max_group])`` ranks; a group
    block-distributes the rows of ONE image and all-reduces once per step, different groups sample different images.
    Every rank creates every group (``dist.new_group`` is collective).  Returns (G, n_groups, my_group, my process group or
    None when G == 1 or world == 1)."""
    group_size = max(1, min(world, n_rows, max_group if max_group else world))
    n_groups = max(world // group_size, 1)
    my_group = min(rank // group_size, n_groups - 1)
    pg = None
    if world > 1 and group_size > 1:
        import torch.distributed as dist
        for gi in range(n_groups):
            g = dist.new_group(list(range(gi * group_size, (gi + 1) * group_size)))
            if gi == my_group and rank < n_groups * group_size:
                pg = g
    return group_size, n_groups, my_group, pg


class Tweediemix(nn.Module):
    def __init__(self, config, components: Optional[FusionComponents] = None, *, variant: str = "custom",
                 use_cuda_graphs: Optional[bool] = None, process_group=None, gate: Optional[int] = None,
                 ref_rounding: bool = False, mask_provider=None):
        super().__init__()
        self.mask_provider = mask_provider          # callable(image_path, seg_concepts, h, w) -> [K,1,h,w]; replaces the os.system hand-off
        if variant not in ("custom", "lora"):
            raise ValueError("variant must be 'custom' or 'lora'")
        self.config = config
        self.variant = variant
        self.hooks = utils_lora if variant == "lora" else utils_custom
        self.gate = gate
        self.ref_rounding = ref_rounding
        if components is None:
            components = load_components_from_diffusers(config, variant)
        self.unet = components.unet
        self.vae = components.vae
        self.device = self.unet.device
        self.masks = None if components.masks is None else components.masks.to(self.device, torch.float32).contiguous()
        self.text_embeds = components.text_embeds
        self.text_embeds_single = components.text_embeds_single
        self.concept_num = self.text_embeds[0].shape[0] - 2
        for i, u in enumerate(components.concept_unets):
            setattr(self, f"unet_{i}", u)
        self.num_routed_concepts = len(components.concept_unets)

        # fusion_sampling.py:212-218 — N_ts is read BEFORE set_timesteps; the cumprod table gets a 1.0 prepended
        # (the caller's scheduler object is left untouched: a private copy takes the reference's in-place edits, so a
        # second Tweediemix built from the same components sees the same table)
        self.scheduler = copy.deepcopy(components.scheduler)
        n_ts = len(self.scheduler.timesteps)
        self.scheduler.set_timesteps(config.n_timesteps, device=self.device)
        self.skip = n_ts // config.n_timesteps
        self.final_alpha_cumprod = self.scheduler.final_alpha_cumprod.to(self.device)
        self.scheduler.alphas_cumprod = torch.cat([torch.tensor([1.0]), self.scheduler.alphas_cumprod.cpu()])
        self._alpha_table = [float(v) for v in self.scheduler.alphas_cumprod.tolist()]
        self._final_alpha = float(self.scheduler.final_alpha_cumprod)
        self._timesteps = [int(v) for v in self.scheduler.timesteps.tolist()]
        self.add_time_ids = compute_time_ids(config).to(self.device)

        self.pg = process_group
        if process_group is not None:
            import torch.distributed as dist
            self.group_size, self.group_rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        else:
            self.group_size, self.group_rank = 1, 0
        self.use_cuda_graphs = (self.device.type == "cuda") if use_cuda_graphs is None else use_cuda_graphs
        self._rowsets = {}
        self._t_dev = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._acc = None
        self._hook_t = None           # the time the hooks were last told (register_time): decides routing, not the forward's t
        self.n_forward_rows = 0
        self.t_stop_cur = None

    # ------------------------------------------------------------------ schedule helpers
    def alpha(self, t):
        """``fusion_sampling.py:305-307`` (shifted table; final alpha below zero)."""
        t = int(t)
        return self.scheduler.alphas_cumprod[t] if t >= 0 else self.final_alpha_cumprod

    def _alpha_f(self, t: int) -> float:
        return self._alpha_table[t] if t >= 0 else self._final_alpha

    def init_fusion(self, t_cond, t_stop=None):
        ts = self.scheduler.timesteps
        if self.variant == "lora":
            # fusion_sampling_lora.py:476-483 — the hook window stops ONE step before the sampler window (quirk ⑦)
            self.t_cond = ts[t_cond:t_stop] if t_cond >= 0 else []
            self.t_stop_cur = int(ts[t_stop])
        else:
            self.t_cond = ts[t_cond:] if t_cond >= 0 else []          # fusion_sampling.py:477
        self.t_cond_prev = int(ts[t_cond - 1])
        self.t_cond_cur = int(ts[t_cond])
        self.start_t = int(ts[0])
        self._window = self.hooks.as_window(self.t_cond)
        self.hooks.register_attention_control_efficient(self, self.t_cond, self.num_routed_concepts, gate=self.gate)
        for i in range(self.num_routed_concepts):                      # fusion_sampling.py:482-483
            delattr(self, f"unet_{i}")
        self._rowsets.clear()

    # ------------------------------------------------------------------ row sets / U-Net forward
    def _prompt_rows(self, name: str):
        E, P = self.text_embeds
        if name == "fused":                   # :325-336  [uncond, c_1 .. c_K]
            return torch.cat([E[0:1], E[2:]]), torch.cat([P[0:1], P[2:]])
        if name == "start":                   # :347-359  [uncond, multi, single_1 .. single_{K-1}]
            Es, Ps = self.text_embeds_single
            return torch.cat([E[0:1], E[1:2], Es[1:]]), torch.cat([P[0:1], P[1:2], Ps[1:]])
        return E[:2], P[:2]                   # :362-366  [uncond, multi]

    def _rowset(self, name: str, like: torch.Tensor) -> _RowSet:
        n_imgs = like.shape[0]
        rs = self._rowsets.get((name, n_imgs))
        if rs is not None:
            return rs
        ehs, pool = self._prompt_rows(name)
        n = ehs.shape[0]
        units = assign_units(n_imgs, n, self.group_size, self.group_rank)
        rs = _RowSet(name, ehs, pool, self.add_time_ids, units, n, n_imgs, like.shape, self.unet.dtype, self.device)
        self._rowsets[(name, n_imgs)] = rs
        return rs

    def _unet_call(self, rs: _RowSet):
        return self.unet(rs.latent, self._t_dev, encoder_hidden_states=rs.ehs, added_cond_kwargs=rs.cond())["sample"]

    def _forward(self, rs: _RowSet, x: torch.Tensor, t: int) -> Optional[torch.Tensor]:
        """eps of this rank's units of ``rs`` at latents ``x`` [imgs,4,h,w] / timestep ``t`` ([R_local,4,h,w]) or None."""
        if not rs.local_units:
            return None
        self.n_forward_rows += len(rs.local_units)
        if rs.img_index is None:
            rs.latent.copy_(x.expand(len(rs.local_units), -1, -1, -1))
        else:
            rs.latent.copy_(x.index_select(0, rs.img_index))
        self._t_dev.fill_(float(t))
        # the hooks route row r of a gate-sized batch; a rank that holds other rows (concept-parallel) or several
        # images says which prompt row every batch row is
        local = rs.local_ids if (self.group_size > 1 or rs.n_imgs > 1) else None
        for _, attn in self.unet.attention_modules():
            attn.local_rows = local
        if not self.use_cuda_graphs:
            return self._unet_call(rs)
        # key = the decision the hooks take: the REGISTERED time (not this forward's t, which differs inside the
        # resampling loop and the jump) in the window; the batch-size half of the gate is fixed per row set
        routed = self._hook_t in self._window
        hit = rs.graphs.get(routed)
        if hit is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):            # warm-up outside capture: lazy inits, K/V caches, workspaces
                    self._unet_call(rs)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            before = dict(ops.LAUNCHES)
            with torch.cuda.graph(graph):
                out = self._unet_call(rs)
            per_replay = {k: v - before.get(k, 0) for k, v in ops.LAUNCHES.items() if v != before.get(k, 0)}
            ops.add_launches(per_replay, -1)          # capture itself launches nothing
            ops.pin_graph_resources()                 # workspaces / cached K/V captured by address must outlive the graph
            hit = rs.graphs[routed] = (graph, out, per_replay)
        hit[0].replay()
        ops.add_launches(hit[2])
        return hit[1]

    # ------------------------------------------------------------------ the fused tail (k7)
    def _blend(self, x, eps, rs: _RowSet, use_rows: int, masks, weights, a_t, a_next, is_last, out=None, x0_out=None):
        """CFG + Tweedie + (masked / weighted) blend + DDIM over prompt rows [0, use_rows) of ``rs`` for every image.
        ``eps`` holds this rank's units (image-major).  Single GPU: one k7 launch.  Concept-parallel: partial ->
        ONE all-reduce -> finish."""
        g = float(self.config.guidance_scale)
        K = use_rows - 1
        imgs = x.shape[0]
        if self.group_size == 1:
            e = eps.reshape(imgs, rs.n_rows, *eps.shape[1:])
            if use_rows != rs.n_rows:
                e = e[:, :use_rows].contiguous() if imgs > 1 else e[:, :use_rows]
            return ops.tweedie_blend_ddim(x, e, masks, a_t, a_next, g, is_last=is_last, weights=weights, out=out,
                                          x0_out=x0_out, ref_rounding=self.ref_rounding)
        import torch.distributed as dist
        if self._acc is None or self._acc.shape[0] != imgs or self._acc.shape[2:] != x.shape[1:]:
            self._acc = torch.empty((imgs, 2) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
        for i in range(imgs):
            first, rows = rs.rows_of_image(i)
            ids = [r for r in rows if r < use_rows]                  # contiguous block distribution => a prefix of `rows`
            ops.blend_partial(eps[first:first + len(ids)] if ids else None, masks, ids, self._acc[i:i + 1], 1, K=K, weights=weights)
        dist.all_reduce(self._acc, group=self.pg)
        return ops.blend_finish(x, self._acc, masks, a_t, a_next, g, is_last=is_last, out=out, x0_out=x0_out, K=K, weights=weights)

    # ------------------------------------------------------------------ the step
    def in_fused_phase(self, t: int) -> bool:
        ok = t <= self.t_cond_cur                                         # fusion_sampling.py:324
        if self.variant == "lora":
            ok = ok and t >= self.t_stop_cur                              # fusion_sampling_lora.py:324
        return ok

    @torch.no_grad()
    def denoise_step(self, x, t, out=None):
        cfg, K = self.config, self.concept_num
        t = int(t)
        next_t = t - self.skip
        at, at_next = self._alpha_f(t), self._alpha_f(next_t)
        self.hooks.register_time(self, t)
        self._hook_t = t
        last = t == 1                                                     # :471-472

        if self.in_fused_phase(t):
            if self.masks is None:
                raise RuntimeError("region masks are not set: pass precomputed masks (FusionComponents.masks / --masks_dir) "
                                   "or provide a mask_provider for the segmentation hand-off")
            rs = self._rowset("fused", x)
            eps = self._forward(rs, x, t)
            return self._blend(x, eps, rs, K + 1, self.masks, None, at, at_next, last, out)

        if t == self.start_t:
            rs = self._rowset("start", x)
            rs2 = self._rowset("cfg", x)
            eps = self._forward(rs, x, t)
            if cfg.resampling_steps <= 0:
                # fusion_sampling.py:417 `del`s names that are only bound inside the resampling loop
                raise UnboundLocalError("the reference cannot run with resampling_steps == 0 (fusion_sampling.py:417)")
            w = [float(K - 1)] + [-1.0] * (K - 1)                         # :395-401
            for _ in range(cfg.resampling_steps):                         # :390-415
                x_low = self._blend(x, eps, rs, K + 1, None, w, at, at_next, False)
                eps_next = self._forward(rs2, x_low, next_t)
                x = self._blend(x_low, eps_next, rs2, 2, None, None, at_next, at, False)     # re-noise = CFG step with swapped alphas
                eps = self._forward(rs, x, t)
        else:
            rs = self._rowset("cfg", x)
            eps = self._forward(rs, x, t)
        handoff = t == self.t_cond_prev and self.masks is None            # :431-469 (output-neutral for x_next)
        x0 = torch.empty_like(x) if handoff else None
        x_next = self._blend(x, eps, rs, 2, None, None, at, at_next, last, out, x0_out=x0)   # :421-430
        if handoff:
            self.masks = self._mask_handoff(x_next, next_t, x0)
        return x_next

    def decode_latent(self, latent):
        """``fusion_sampling.py:297-303``: ``1/0.18215`` (quirk 13), VAE decode, to [0,1].  The latent is fp32 and the VAE
        usually fp16: cast to the VAE's dtype (the reference gets the same effect from its autocast region, ``:298``)."""
        vae_dtype = next(self.vae.parameters()).dtype if any(True for _ in self.vae.parameters()) else latent.dtype
        img = self.vae.decode((latent / 0.18215).to(vae_dtype)).sample
        return (img.float() / 2 + 0.5).clamp(0, 1)

    def _mask_handoff(self, latent, t_start, x0):
        """Jump ``jumping_steps`` x 150 timesteps ahead with plain CFG (``:436-447``), decode the jumped Tweedie x0 —
        the step's own x0 when ``jumping_steps == 0`` (``:435,449-452``) — write ``tweedie.jpg`` and run the
        segmentation subprocess (``:453-469``).  Needs a VAE and the reference's ``text_segment`` stage (or a
        ``mask_provider`` callable), both outside the hot path."""
        if self.vae is None:
            raise RuntimeError("no precomputed masks and no VAE: cannot run the segmentation hand-off")
        if self.group_size > 1:
            raise RuntimeError("the segmentation hand-off is a single-process path: pass precomputed masks "
                               "(FusionComponents.masks / --masks_dir) to a concept-parallel run")
        if latent.shape[0] != 1:
            raise RuntimeError("the segmentation hand-off handles one image (the reference's batch-of-one latent)")
        rs2 = self._rowset("cfg", latent)
        lat, t_tmp = latent, t_start
        for _ in range(max(self.config.jumping_steps, 0)):
            a_tmp = self._alpha_f(t_tmp)
            e = self._forward(rs2, lat, t_tmp)
            t_tmp -= 150                                                  # :444 (fixed stride, quirk 8)
            x0 = torch.empty_like(lat)
            lat = self._blend(lat, e, rs2, 2, None, None, a_tmp, self._alpha_f(t_tmp), False, x0_out=x0)
        img = self.decode_latent(x0)                                      # :453
        path = os.path.join(self.config.output_path, "tweedie.jpg")
        save_image(img[0], path)                                          # :455
        h, w = self.config.resolution_h // 8, self.config.resolution_w // 8
        if self.mask_provider is not None:
            return self.mask_provider(path, self.config.seg_concepts, h, w).to(self.device, torch.float32).contiguous()
        os.system(f'CUDA_VISIBLE_DEVICES={self.config.seg_gpu} python text_segment/run_expand.py --input_path={path} '
                  f'--text_condition="{self.config.seg_concepts}" --output_path={self.config.output_path}')
        return load_region_masks(self.config.output_path, self.config.seg_concepts, h, w, self.device)

    # ------------------------------------------------------------------ loops
    def initial_latent(self):
        """``:488`` — drawn on the CPU generator (seeded by ``seed_everything``), then moved."""
        h, w = self.config.resolution_h // 8, self.config.resolution_w // 8
        return torch.randn(1, 4, h, w).to(self.device) * self.scheduler.init_noise_sigma

    def run_fusion(self):
        if self.variant == "lora":
            self.init_fusion(int(self.config.n_timesteps * self.config.t_cond), int(self.config.n_timesteps * self.config.t_stop))
        else:
            self.init_fusion(int(self.config.n_timesteps * self.config.t_cond))
        return self.sample_loop(self.initial_latent())

    @torch.no_grad()
    def sample_loop(self, x, callback=None):
        x = x.to(self.device, torch.float32).contiguous().clone()
        for i, t in enumerate(self._timesteps):                           # :493-494
            x = self.denoise_step(x, t, out=x if t != self.start_t else None)
            if callback is not None:
                callback(i, t, x)
        return x

    def set_text(self, text_embeds, text_embeds_single):
        """New prompts for the next image: overwrite the static text buffers in place and re-project
        the cached cross-attention K/V (captured graphs stay valid)."""
        self.text_embeds, self.text_embeds_single = text_embeds, text_embeds_single
        for (name, _), rs in self._rowsets.items():
            if not rs.local_units:
                continue
            ehs, pool = self._prompt_rows(name)
            rs.ehs.copy_(ehs[rs.local_ids].to(rs.ehs.device), non_blocking=True)
            rs.pooled.copy_(pool[rs.local_ids].to(rs.pooled.device), non_blocking=True)
        self.unet.refresh_text_cache()

    def set_masks(self, masks):
        m = masks.to(self.device, torch.float32)
        if self.masks is not None and self.masks.shape == m.shape:
            self.masks.copy_(m)
        else:
            self.masks = m.contiguous()


def load_components_from_diffusers(config, variant):
    """The reference constructor's loading path (``fusion_sampling.py:119-210``).  Needs diffusers +
    transformers + SDXL weights, none of which exist in the offline build image."""
    try:
        import diffusers  # noqa: F401
    except ImportError as e:
        raise RuntimeError(
            "diffusers is not installed: build FusionComponents yourself (tweediemix_b200.synthetic.make_components "
            "for seeded stand-ins, or checkpoints.load_* with local SDXL weights) and pass it to Tweediemix") from e
    from .checkpoints import components_from_diffusers
    return components_from_diffusers(config, variant)


def build_parser(lora: bool = False) -> argparse.ArgumentParser:
    """The reference's flags with the reference's defaults (``fusion_sampling.py:534-585``; ``--t_stop``
    from ``fusion_sampling_lora.py:547``), plus ``--masks_dir`` / ``--synthetic`` for offline runs."""
    p = argparse.ArgumentParser()
    p.add_argument('--seed', type=int, default=182)
    p.add_argument('--device', type=str, default='cuda')
    p.add_argument('--output_path', type=str, default='results')
    p.add_argument('--output_path_all', type=str, default='results_all')
    p.add_argument('--negative_prompt', type=str, default='ugly, blurry, black, low res, unrealistic')
    p.add_argument('--sd_version', type=str, default='xl', choices=['1.4', '1.5', '2.0', '2.1', 'xl'])
    p.add_argument('--t_cond', type=float, default=0.4)
    if lora:
        p.add_argument('--t_stop', type=float, default=0.9)
    p.add_argument('--guidance_scale', type=float, default=9.0)
    p.add_argument('--n_timesteps', type=int, default=50)
    p.add_argument('--prompt', type=str, default='')
    p.add_argument('--prompt_orig', type=str, default='')
    p.add_argument('--seg_concepts', type=str, default='')
    p.add_argument('--personal_checkpoint', type=str, default='')
    p.add_argument('--concepts', type=str, default='')
    p.add_argument('--modifier_token', type=str, default='')
    p.add_argument('--resampling_steps', type=int, default=10)
    p.add_argument('--jumping_steps', type=int, default=5)
    p.add_argument('--seg_gpu', type=int, default=1)
    p.add_argument('--crops_coords_top_left_h', type=int, default=0)
    p.add_argument('--crops_coords_top_left_w', type=int, default=0)
    p.add_argument('--resolution_h', type=int, default=1024)
    p.add_argument('--resolution_w', type=int, default=1024)
    # additions (not in the reference)
    p.add_argument('--masks_dir', type=str, default='', help='directory with precomputed <seg_concept>.jpg masks; skips the segmentation subprocess')
    p.add_argument('--synthetic', action='store_true', help='seeded random weights / text embeddings (no checkpoints needed)')
    p.add_argument('--dtype', type=str, default='bf16', choices=['bf16', 'fp16'])
    return p


def main(argv=None, lora: bool = False):
    global opt
    opt = build_parser(lora).parse_args(argv)
    os.makedirs(opt.output_path, exist_ok=True)
    utils_custom.seed_everything(opt.seed)
    variant = "lora" if lora else "custom"
    components = None
    if opt.synthetic:
        from .synthetic import make_components
        k = len(opt.concepts.split('+')) if opt.concepts else 3
        components = make_components(concept_num=k, variant=variant, seed=opt.seed, device=opt.device,
                                     dtype=torch.bfloat16 if opt.dtype == 'bf16' else torch.float16,
                                     latent_hw=(opt.resolution_h // 8, opt.resolution_w // 8))
        if opt.personal_checkpoint:
            # real concept checkpoints (delta-*.bin, '+'-separated) on the seeded base U-Net: the reader of
            # fusion_sampling.py:157-158,203-210 / fusion_sampling_lora.py:203-210
            from .checkpoints import concept_from_checkpoint
            paths = opt.personal_checkpoint.split('+')
            if len(paths) != k:
                raise ValueError(f"--personal_checkpoint names {len(paths)} files for {k} concepts")
            components.concept_unets = [concept_from_checkpoint(components.unet, pth, variant) for pth in paths]
    model = Tweediemix(opt, components, variant=variant)
    if opt.masks_dir:
        model.set_masks(load_region_masks(opt.masks_dir, opt.seg_concepts, opt.resolution_h // 8, opt.resolution_w // 8))
    latent = model.run_fusion()
    prompt_orig = opt.prompt_orig.split('+')[0] if opt.prompt_orig else 'sample'
    if model.vae is not None:
        vae_dtype = next(model.vae.parameters()).dtype                    # fp32 latent, fp16 VAE (:514-520 run under autocast)
        img = model.vae.decode((latent / model.vae.config.scaling_factor).to(vae_dtype)).sample
        save_image((img[0].float() / 2 + 0.5).clamp(0, 1), os.path.join(opt.output_path, f'{prompt_orig}_{opt.seed}.png'))
    else:
        torch.save(latent.cpu(), os.path.join(opt.output_path, f'{prompt_orig}_{opt.seed}.latent.pt'))
    return latent


if __name__ == '__main__':
    main()
