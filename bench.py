#!/usr/bin/env python
"""Benchmark of the TweedieMix fusion-sampling hot path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): 1024px images/sec @ 50 DDIM steps, K concepts (+ per-denoise-step ms).
One "step" of this benchmark = ONE COMPLETE IMAGE: the reference's 50-step schedule for
BASELINE configs[1] (SDXL 1024x1024, K=2 foreground concepts + background => concept_num 3,
U-Net batch 4, guidance 0.8, t_cond 0.2, 10 resampling iterations at step 0) = 242 U-Net
sample-forwards + 71 fused CFG/Tweedie/blend/DDIM launches.  Region masks are precomputed inputs
(north star), so the output-neutral jump loop / VAE / segmentation subprocess are not run.
Weights, text embeddings and the start latent are seeded synthetic data (no network, no checkpoints).

  value : images/s with all inputs resident in HBM, CUDA-event timed, max over ranks.
  e2e   : same through the public API (Tweediemix.set_masks / set_text / sample_loop) starting from
          pinned HOST buffers: per image the masks, text embeddings and start latent are copied H2D
          (and the cross-attention K/V cache re-projected) and the final latent is read back D2H.
  roofline     : the dominant hand-written kernel (tcgen05 attention), timed per launch with CUDA
                 events in an instrumented eager pass of one fused denoise step, after the timed region.
  cpu_baseline : oracle port of the reference's PyTorch CPU path on the host cores (rank 0, N=1).

Multi-GPU (--gpus N): concept-parallel groups of G ranks share one image batch ((image, prompt row) units
sharded, one all-reduce per step); N/G groups sample different images concurrently.  G defaults to the size
that maximises images/s for the config (configs[1]: 2; --group-size overrides).  `--config 2|3` runs the
LoRA / K=8 x batch-4 BASELINE configs as the main line; on the default run they ride along as extra legs
under `other_configs` when enough GPUs are present (configs[2] at >= 4, configs[3] at 8).  Before timing,
a multi-rank run checks its sharded latents against the same model run unsharded (`concept_parallel_check`).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_CONCEPTS = 3                      # configs[1]: concept_num incl. background (BASELINE "K=2 (cat+dog)" + background)
FORWARDS_PER_IMAGE = 242            # configs[1]; other configs report the count the sampler actually executed
MASK_DIR = os.path.join(ROOT, "tests", "golden", "masks", "test_out")
MASK_NAMES = "a cat+a dog"

# BASELINE.json configs this file can run ("K" there counts foreground concepts; concept_num adds the background).
CONFIGS = {
    1: dict(name="configs[1]", variant="custom", concept_num=3, image_batch=1, group=None,
            text="SDXL-base U-Net 1024x1024 (latent 128x128), K=2 concepts + background (concept_num 3, U-Net batch 4), "
                 "50 DDIM steps, guidance 0.8, t_cond 0.2, 10 resampling iterations; masks precomputed"),
    2: dict(name="configs[2]", variant="lora", concept_num=3, image_batch=1, group=4,
            text="SDXL-base U-Net 1024x1024, K=3 LoRA concept rows + unconditional row (fusion_sampling_lora.py, t_stop 0.8), "
                 "50 DDIM steps, rank-4 LoRA on q/k/v/out of all 140 attentions; batch rows sharded over up to 4 GPUs"),
    3: dict(name="configs[3]", variant="custom", concept_num=9, image_batch=4, group=8,
            text="SDXL-base U-Net 1024x1024, K=8 concepts + background (concept_num 9, 10 prompt rows) x image batch 4 = 40 "
                 "(image, row) units per step, 50 DDIM steps; units sharded over up to 8 GPUs (5 per rank at N=8)"),
}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--config", type=int, default=1, choices=[1, 2, 3], help="BASELINE.json configs[i] to run as the main line")
    p.add_argument("--variant", default=None, choices=["custom", "lora"], help="override the config's hook variant")
    p.add_argument("--concepts", type=int, default=None, help="override concept_num (incl. background); U-Net rows = concepts + 1")
    p.add_argument("--image-batch", type=int, default=None, help="images sampled together by one concept-parallel group")
    p.add_argument("--group-size", type=int, default=None,
                   help="ranks per concept-parallel group (default: configs[1] -> 2, the size that maximises images/s; "
                        "configs[2] -> 4, configs[3] -> 8; always clamped to the GPU and unit count)")
    p.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    p.add_argument("--workload", default="image50", choices=["image50", "fused_step"],
                   help="image50 = one full 50-step image (batch) per step (the metric); fused_step = one fused-phase denoise step (profiling)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-graphs", action="store_true")
    p.add_argument("--no-extra-configs", action="store_true",
                   help="skip the extra legs (configs[2] at >= 4 GPUs, configs[3] at 8 GPUs) reported under other_configs")
    p.add_argument("--no-rank-check", action="store_true", help="skip the N-rank vs 1-rank latent check before timing")
    p.add_argument("--seed", type=int, default=3821)
    p.add_argument("--ncu-range", action="store_true",
                   help="profiling aid: after warm-up run ONE eager fused denoise step between cudaProfilerStart/Stop and exit "
                        "(use with ncu --profile-from-start off)")
    return p.parse_args()


def resolve_spec(args, world, config=None):
    """The workload of one leg: BASELINE config + command-line overrides + the group-size policy."""
    c = dict(CONFIGS[config if config is not None else args.config])
    if config is None:
        if args.variant:
            c["variant"] = args.variant
        if args.concepts:
            c["concept_num"] = args.concepts
        if args.image_batch:
            c["image_batch"] = args.image_batch
        if args.group_size:
            c["group"] = args.group_size
    units = (c["concept_num"] + 1) * c["image_batch"]
    if c["group"] is None:
        # configs[1]: 4 units per image.  Measured (profiles/README.md): a 2-rank group keeps 0.87 of the per-GPU rate
        # (18.9 vs 33.5/2 ms per fused step, both 2-row phases split evenly), a 4-rank group 0.61 (a one-row forward is
        # 12.0 ms, and the 2-row phases leave two ranks idle) -> images/s is maximised by 2-rank groups.
        c["group"] = 2
    c["group"] = max(1, min(c["group"], world, units))
    while world % c["group"]:
        c["group"] -= 1
    c["n_groups"] = world // c["group"]
    return c


def namespace(variant):
    return argparse.Namespace(guidance_scale=0.8, n_timesteps=50, t_cond=0.2, t_stop=0.8 if variant == "lora" else None,
                              resampling_steps=10, jumping_steps=5, resolution_h=1024, resolution_w=1024,
                              crops_coords_top_left_h=0, crops_coords_top_left_w=0, seed=3821, output_path=".",
                              seg_concepts=MASK_NAMES, seg_gpu=0)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            d = json.load(fh)
        return {"hbm": d["hbm_gbs"], "tensor_burst": d["bf16_tflops"], "tensor_sustained": d["bf16_tflops_sustained"], "src": "measured"}
    except Exception:
        return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        return False

    def summary(self):
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax = max(smax, float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ================================================================================ reference arm (CPU)

def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle.cpu_baseline import CpuReference
    spec = resolve_spec(args, max(args.gpus, 1))
    ref = CpuReference(res=1024)
    for _ in range(min(args.warmup, 1)):            # one warm-up forward is enough on CPU (allocator, thread pool)
        ref.forward_seconds()
    t = [ref.forward_seconds() for _ in range(args.steps)]
    sec = sum(t) / len(t)
    val = 1.0 / (forwards_per_image(spec) * sec)
    line = {"impl": "reference", "metric": "1024px images/sec @50 DDIM steps, K concepts", "value": val, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1000.0 / val,
            "higher_is_better": True, "scaling": scaling_label(spec), "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, spec),
            "gpu_launches": 0,
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": ref.threads, "kind": "port", "sample": ref.sample_description(),
                             "seconds_per_sample_forward": sec},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)
    return 0


def forwards_per_image(spec, n=50, t_cond=0.2, t_stop=0.8, resampling=10):
    """U-Net sample-forwards that shape one image (the output-neutral jump is not run, masks are inputs):
    start step (K+1)(1+R) + 2R, plain-CFG steps 2 each, fused steps K+1 each."""
    rows = spec["concept_num"] + 1
    ic = int(n * t_cond)
    fused = (int(n * t_stop) - ic + 1) if spec["variant"] == "lora" else (n - ic)
    return rows * (1 + resampling) + 2 * resampling + 2 * (n - 1 - fused) + rows * fused


def scaling_label(spec):
    """'strong' while ONE concept-parallel group shares the job (total work fixed, split over more GPUs); 'weak' once
    further GPUs add further groups that sample their own images (per-GPU work fixed)."""
    return "weak" if spec["n_groups"] > 1 else "strong"


def workload_config(args, spec):
    g, ng, ib = spec["group"], spec["n_groups"], spec["image_batch"]
    return {"workload": spec["text"],
            "baseline_config": spec["name"], "variant": spec["variant"], "concept_num": spec["concept_num"], "image_batch": ib,
            "step_unit": (f"one 50-step image batch ({ib} image(s)) per group" if args.workload == "image50" else "one fused denoise step"),
            "sample_forwards_per_image": forwards_per_image(spec),
            "parallelism": f"concept-parallel x{g} (batch rows sharded, one all-reduce per step), {ng} image group(s)",
            "group_size": g, "image_groups": ng,
            "cuda_graphs": not args.no_graphs,
            "l2": "no flush needed: every step streams 5.1 GB of bf16 weights (>> 126 MB L2)"}


# ================================================================================ our arm (GPU)

class Leg:
    """One workload (a BASELINE config) set up on this process' GPU: model, host buffers, step functions."""

    def __init__(self, args, spec, world, rank, dev, torch, dist):
        from tweediemix_b200.fusion_sampling import Tweediemix, make_concept_groups
        from tweediemix_b200.masks import load_region_masks, stripe_masks
        from tweediemix_b200.synthetic import make_components, make_text
        from tweediemix_b200.unet import UNetConfig
        self.args, self.spec, self.world, self.rank, self.dev, self.torch, self.dist = args, spec, world, rank, dev, torch, dist
        K, ib = spec["concept_num"], spec["image_batch"]
        gs, n_groups, my_group, pg = make_concept_groups(world, rank, (K + 1) * ib, spec["group"])
        assert gs == spec["group"] and n_groups == spec["n_groups"], (gs, n_groups, spec)
        self.group_size, self.n_groups, self.my_group, self.pg = gs, n_groups, my_group, pg
        dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float16
        if K == 3:
            masks = load_region_masks(MASK_DIR, MASK_NAMES, 128, 128)        # the reference's shipped cat / dog masks
        else:
            masks = stripe_masks(K, 128, 128)                               # no fixture for other K: stripe partition
        self.masks_host = masks.pin_memory()
        comp = make_components(K, spec["variant"], seed=args.seed, device=dev, dtype=dtype, masks=self.masks_host)
        self.comp = comp
        self.model = model = Tweediemix(namespace(spec["variant"]), comp, variant=spec["variant"],
                                        use_cuda_graphs=not args.no_graphs, process_group=pg if gs > 1 else None)
        ns = model.config
        if spec["variant"] == "lora":
            model.init_fusion(int(ns.n_timesteps * ns.t_cond), int(ns.n_timesteps * ns.t_stop))
        else:
            model.init_fusion(int(ns.n_timesteps * ns.t_cond))
        gen = torch.Generator().manual_seed(args.seed + 17 * my_group)
        self.x_host = torch.randn(ib, 4, 128, 128, generator=gen).pin_memory()   # CPU-generator draw, like fusion_sampling.py:488
        text_host, single_host = make_text(UNetConfig.sdxl_base(), K, args.seed + 1, dtype=dtype)
        self.text_host = tuple(t.pin_memory() for t in text_host)
        self.single_host = tuple(t.pin_memory() for t in single_host)
        self.out_host = torch.empty(ib, 4, 128, 128).pin_memory()
        self.x_dev = self.x_host.to(dev)
        self.fused_t = [t for t in model._timesteps if model.in_fused_phase(t)]
        self.h2d = self.masks_host.numel() * 4 + self.x_host.numel() * 4 + sum(t.numel() * t.element_size() for t in self.text_host + self.single_host)
        self.d2h = self.out_host.numel() * 4

    def step_resident(self):
        if self.args.workload == "image50":
            return self.model.sample_loop(self.x_dev)
        return self.model.denoise_step(self.x_dev, self.fused_t[0])

    def step_e2e(self):
        torch, dev, m = self.torch, self.dev, self.model
        m.set_masks(self.masks_host.to(dev, non_blocking=True))
        m.set_text(tuple(t.to(dev, non_blocking=True) for t in self.text_host), tuple(t.to(dev, non_blocking=True) for t in self.single_host))
        x = self.x_host.to(dev, non_blocking=True)
        y = m.sample_loop(x) if self.args.workload == "image50" else m.denoise_step(x, self.fused_t[0])
        self.out_host.copy_(y, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def sync_all(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        torch = self.torch
        self.sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def rank_check(self):
        """Concept-parallel sanity check before timing: a short chain of denoise steps that touches every phase — the start
        step (its resampling loop cut to one iteration: ten iterations at t = 981, where Tweedie's formula amplifies eps 13x,
        turn 16-bit rounding noise into O(0.1) differences), two plain-CFG steps and three fused steps — through the sharded
        path (partial -> NCCL all-reduce -> finish) against the SAME model run unsharded on this rank.  The two differ by
        16-bit rounding only (cuBLAS / cuDNN pick other kernels for a 1- or 2-row batch than for 4 rows), so the bound is
        relative: 0.2 of max|x| separates that noise (~1e-2) from a wrong row mapping, a missing partial or a diverged rank
        (O(1)); the measured value is reported in the JSON line, tests/test_gpu_nccl.py holds the tight kernel-level bound.
        The latents of the ranks of a group must be bit-identical."""
        torch, dist, m = self.torch, self.dist, self.model
        if self.group_size == 1:
            return None
        ts = m._timesteps[:3] + self.fused_t[:3]
        resampling = m.config.resampling_steps
        m.config.resampling_steps = 1

        def chain():
            x = self.x_dev.clone()
            for t in ts:
                x = m.denoise_step(x, t)
            return x

        try:
            x = chain()
            gathered = [torch.empty_like(x) for _ in range(self.group_size)]
            dist.all_gather(gathered, x, group=self.pg)
            identical = all(torch.equal(gathered[0], g) for g in gathered[1:])
            pg, gs, gr, graphs, rowsets = m.pg, m.group_size, m.group_rank, m.use_cuda_graphs, m._rowsets
            m.pg, m.group_size, m.group_rank, m.use_cuda_graphs, m._rowsets = None, 1, 0, False, {}
            try:
                y = chain()
            finally:
                m.pg, m.group_size, m.group_rank, m.use_cuda_graphs, m._rowsets = pg, gs, gr, graphs, rowsets
        finally:
            m.config.resampling_steps = resampling
        diff = float((x - y).abs().max().item())
        scale = float(y.abs().max().item())
        res = {"steps": len(ts), "max_abs_diff": diff, "max_abs_latent": scale, "rel": diff / scale, "ranks_bit_identical": identical,
               "bound_rel": 0.2}
        if not identical or not (diff <= 0.2 * scale):
            raise RuntimeError(f"concept-parallel check failed on rank {self.rank}: {res}")
        return res

    def measure(self, steps, warmup):
        """Warm-up, timed resident + e2e regions, per-denoise-step ms.  Returns a dict (identical on every rank)."""
        from tweediemix_b200 import ops
        torch = self.torch
        m = self.model
        check = None if self.args.no_rank_check else self.rank_check()
        m.n_forward_rows = 0
        for _ in range(warmup):
            self.step_resident()
        rows = torch.tensor([float(m.n_forward_rows)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(rows)
        launches0 = ops.launch_count()
        local = int(os.environ.get("LOCAL_RANK", "0"))
        with ClockSampler(local) as clk:
            total_ms = self.timed(self.step_resident, steps)
        launches = ops.launch_count() - launches0
        clocks = clk.summary()
        self.step_e2e()                                      # warm (first set_text re-projects K/V)
        e2e_ms = self.timed(self.step_e2e, steps)
        images_per_step = self.n_groups * self.spec["image_batch"]      # images finished per benchmark step, whole job
        ms_per_step = total_ms / steps
        n_f = 8
        m.denoise_step(self.x_dev, self.fused_t[0])
        fused_ms = self.timed(lambda: m.denoise_step(self.x_dev, self.fused_t[1]), n_f) / n_f
        unit = "images/s" if self.args.workload == "image50" else "steps/s"
        return {"value": images_per_step * 1000.0 / ms_per_step, "unit": unit, "ms_per_step": ms_per_step,
                "e2e": {"value": images_per_step * 1000.0 / (e2e_ms / steps), "unit": unit,
                        "h2d_bytes_per_step": self.h2d * self.world, "d2h_bytes_per_step": self.d2h * self.world,   # every rank copies its replica
                        "ms_per_step": e2e_ms / steps},
                "per_denoise_step_ms": fused_ms, "clocks": clocks, "gpu_launches": launches,
                "sample_forwards_executed_per_image": (rows.item() / max(warmup, 1)) / images_per_step if warmup and self.args.workload == "image50" else None,
                "concept_parallel_check": check}


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from tweediemix_b200 import build, ops
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()

    spec = resolve_spec(args, world)
    leg = Leg(args, spec, world, rank, dev, torch, dist)
    model, x_dev, fused_t = leg.model, leg.x_dev, leg.fused_t
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float16

    if args.ncu_range:
        model.use_cuda_graphs = False
        for _ in range(max(args.warmup, 1)):
            model.denoise_step(x_dev, fused_t[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        model.denoise_step(x_dev, fused_t[1])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return 0

    main = leg.measure(args.steps, args.warmup)

    # ---- roofline of the hand-written kernels: instrumented EAGER fused step, CUDA events per launch
    roof, others = None, []
    pk = peaks()
    graphs_were = model.use_cuda_graphs
    model.use_cuda_graphs = False
    model.denoise_step(x_dev, fused_t[1])                 # eager warm-up
    prof = ops.KernelProfile()
    ops.set_profile(prof)
    model.denoise_step(x_dev, fused_t[1])
    ops.set_profile(None)
    model.use_cuda_graphs = graphs_were
    summ = prof.summary()
    fam = {}
    for (f, tag), d in summ.items():
        fam.setdefault(f, []).append((tag, d))
    if "attention" in fam:
        self_attn = [(tag, d) for tag, d in fam["attention"] if "Nk77" not in tag]
        work = sum(d["work"] for _, d in self_attn); ms = sum(d["ms"] for _, d in self_attn); n = sum(d["launches"] for _, d in self_attn)
        if ms > 0:
            # The eager per-launch event pairs above include the host's launch gap (~3 us on a 40 us kernel).  The step itself runs as a
            # CUDA graph, so the kernel's launch duration is measured the same way: the step's attention launches (same shapes, packed
            # q|k|v views, buffers rotated beyond L2) replayed back to back in one graph, CUDA events around the replay.
            dev_ms, by_graph = 0.0, {}
            try:
                for tag, d in self_attn:
                    us = _attention_graph_time(torch, ops, tag, d, dtype, dev)
                    by_graph[tag] = {"launches": d["launches"], "avg_ms": us * 1e-3, "tflops": d["work"] / d["launches"] / (us * 1e-6) / 1e12}
                    dev_ms += us * 1e-3 * d["launches"]
            except Exception as e:                          # never let the side measurement break the headline
                by_graph, dev_ms = {"error": str(e)}, 0.0
            use_ms = dev_ms if dev_ms > 0 else ms
            ach = work / (use_ms * 1e-3) / 1e12
            roof = {"kernel": "attn_fwd_kernel (tcgen05 self-attention, all launches of one fused step)", "bound": "tensor",
                    "achieved": ach, "peak": pk["tensor_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tensor_sustained"],
                    "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({pk['src']})",
                    "timing": "CUDA events around a graph replay of the step's self-attention launches" if dev_ms > 0 else "eager per-launch CUDA events",
                    **_ncu_attention_traffic({tag: d["launches"] for tag, d in self_attn}),
                    "launches": n, "avg_launch_ms": use_ms / n, "flops_per_launch": work / n,
                    "by_shape": by_graph,
                    "by_shape_eager_events": {tag: {"launches": d["launches"], "avg_ms": d["ms"] / d["launches"],
                                                    "tflops": d["work"] / (d["ms"] * 1e-3) / 1e12} for tag, d in fam["attention"]}}
    for f in ("groupnorm", "resadd", "resadd_ln", "layernorm", "geglu", "blend"):
        if f in fam:
            work = sum(d["work"] for _, d in fam[f]); ms = sum(d["ms"] for _, d in fam[f]); n = sum(d["launches"] for _, d in fam[f])
            if ms > 0:
                gbs = work / (ms * 1e-3) / 1e9
                others.append({"kernel": f, "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                               "calls": n, "total_ms": ms})
    for f in ("linear",):
        if f in fam:
            work = sum(d["work"] for _, d in fam[f]); ms = sum(d["ms"] for _, d in fam[f]); n = sum(d["launches"] for _, d in fam[f])
            if ms > 0:
                tfs = work / (ms * 1e-3) / 1e12
                others.append({"kernel": "linear_kernel (tcgen05 GEMM + fused epilogues, all launches of one fused step)", "bound": "tensor",
                               "achieved": tfs, "peak": pk["tensor_sustained"], "unit": "TFLOP/s", "frac": tfs / pk["tensor_sustained"],
                               "calls": n, "total_ms": ms,
                               "by_shape": {tag: {"launches": d["launches"], "avg_ms": d["ms"] / d["launches"],
                                                  "tflops": d["work"] / (d["ms"] * 1e-3) / 1e12} for tag, d in fam[f]}})
    step_kernel_ms = {f: sum(d["ms"] for _, d in v) for f, v in fam.items()}
    eager_note = ("per-launch CUDA events in an eager step; with one or two batch rows per GPU the launches are host-bound and these "
                  "per-kernel times over-state the device time" if spec["group"] > 1 else "per-launch CUDA events in an eager step")

    # k7 at a size where HBM (not launch latency) is the bound: 2048 stacked images (2.5 GB)
    try:
        imgs = 2048
        Kc = spec["concept_num"]
        xb = torch.randn(imgs, 4, 128, 128, device=dev)
        eb = torch.randn(imgs, Kc + 1, 4, 128, 128, device=dev, dtype=dtype)
        mb = model.masks
        ob = torch.empty_like(xb)
        for _ in range(3):
            ops.tweedie_blend_ddim(xb, eb, mb, 0.0438, 0.0518, 0.8, out=ob)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.tweedie_blend_ddim(xb, eb, mb, 0.0438, 0.0518, 0.8, out=ob)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        nbytes = imgs * (4 * 16384 * 8 + (Kc + 1) * 4 * 16384 * eb.element_size()) + Kc * 16384 * 4
        gbs = nbytes / (ms * 1e-3) / 1e9
        others.append({"kernel": "blend (k7) on 2048 stacked images", "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
                       "frac": gbs / pk["hbm"], "avg_launch_ms": ms, "bytes_per_launch": nbytes})
        del xb, eb, ob
    except Exception as e:                                 # never let the side measurement break the headline
        others.append({"kernel": "blend (k7) on 2048 stacked images", "error": str(e)})

    # ---- extra legs: the BASELINE configs that need several GPUs ride on the driver's 4- and 8-GPU runs
    extra = []
    if not args.no_extra_configs and args.config == 1 and args.workload == "image50":
        for cfg_id, need in ((2, 4), (3, 8)):
            if world < need:
                continue
            entry = {"baseline_config": CONFIGS[cfg_id]["name"]}
            try:
                torch.cuda.empty_cache()
                sp = resolve_spec(args, world, cfg_id)
                lg = Leg(args, sp, world, rank, dev, torch, dist)
                r = lg.measure(max(1, min(args.steps, 2)), 1)
                entry.update({"config": workload_config(args, sp), "scaling": scaling_label(sp), "n_gpus": world, **r})
                del lg
            except Exception as e:                         # an extra leg must never take the headline line down with it
                entry["error"] = f"{type(e).__name__}: {e}"
                if world > 1:                              # the ranks may have diverged: do not attempt further collectives
                    extra.append(entry)
                    break
            extra.append(entry)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle.cpu_baseline import CpuReference
            torch.cuda.empty_cache()
            ref = CpuReference(res=1024)
            sec = ref.forward_seconds()
            cpu = {"value": 1.0 / (forwards_per_image(spec) * sec), "unit": "images/s", "cores": ref.threads, "kind": "port",
                   "sample": ref.sample_description(), "seconds_per_sample_forward": sec}
        except Exception as e:
            cpu = {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        metric = "1024px images/sec @50 DDIM steps, K concepts" if args.workload == "image50" else "fused denoise steps/sec (profiling workload)"
        line = {"metric": metric, "value": main["value"], "unit": main["unit"],
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"],
                "higher_is_better": True, "scaling": scaling_label(spec), "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": workload_config(args, spec),
                "per_denoise_step_ms": main["per_denoise_step_ms"], "clocks": main["clocks"],
                "e2e": main["e2e"],
                "gpu_launches": main["gpu_launches"], "gpu_launches_by_kernel": {k: v for k, v in ops.LAUNCHES.items()},
                "sample_forwards_executed_per_image": main["sample_forwards_executed_per_image"],
                "concept_parallel_check": main["concept_parallel_check"],
                "roofline": roof, "roofline_other": others, "fused_step_tmx_kernel_ms": step_kernel_ms,
                "fused_step_tmx_kernel_ms_note": eager_note,
                "other_configs": extra,
                "cpu_baseline": cpu}
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _attention_graph_time(torch, ops, tag, d, dtype, dev, reps=16):
    """Device microseconds per launch of the self-attention shape `tag` ("Nq{N}_Nk{N}_H{H}"), batch recovered from the launch's FLOPs."""
    import re
    m = re.match(r"Nq(\d+)_Nk(\d+)_H(\d+)", tag)
    N, H = int(m.group(1)), int(m.group(3))
    B = max(1, round(d["work"] / d["launches"] / (4.0 * H * N * N * 64)))
    nb = 4
    qkv = [torch.randn(B, N, 3 * H * 64, device=dev, dtype=dtype) for _ in range(nb)]
    outs = [torch.empty(B, N, H * 64, device=dev, dtype=dtype) for _ in range(nb)]
    hd = H * 64

    def run(i):
        t = qkv[i % nb]
        ops.attention(t[..., :hd], t[..., hd:2 * hd], t[..., 2 * hd:], H, out=outs[i % nb])

    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(reps):
                run(i)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


_JSON_FD = None


def _ncu_attention_traffic(launches_by_tag):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the attention kernel, from the committed
    `ncu --set full` summaries under profiles/ (one capture per self-attention shape), weighted by this step's launch mix."""
    import glob
    import re
    root = os.path.dirname(os.path.abspath(__file__))
    per_shape, used = {}, []
    for tag in launches_by_tag:
        m = re.match(r"Nq(\d+)_Nk(\d+)", tag)
        files = sorted(glob.glob(os.path.join(root, "profiles", f"r*_ncu_attn_n{m.group(1)}_final.txt"))) if m and m.group(1) == m.group(2) else []
        if not files:
            return {"traffic": None}
        tot = 0.0
        for line in open(files[-1]):
            mm = re.match(r"\s*dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
            if mm:
                tot += float(mm.group(2)) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[mm.group(3)]
        per_shape[tag] = tot
        used.append(os.path.basename(files[-1]))
    n = sum(launches_by_tag.values())
    return {"traffic": sum(per_shape[t] * c for t, c in launches_by_tag.items()) / n,
            "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, launch-weighted over " + ", ".join(used)}


def _protect_stdout():
    """The contract is ONE JSON line on stdout: send everything libraries print to fd 1 (NCCL's version banner, ...) to
    stderr and keep a private duplicate of the real stdout for the result line."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _protect_stdout()
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
