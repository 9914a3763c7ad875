"""-m gpu: the product host path (U-Net, hooks, sampler) on the REAL tmx kernels vs the fp32 CPU oracle.

Tolerances.  The oracle is fp32 end to end; the product computes in fp16 / bf16 with fp32 accumulation
(what the reference itself does under ``autocast(fp16)``), so agreement is bounded by
the 16-bit roundings of ~600 chained layers, not by the kernels (those are pinned tightly in
test_gpu_kernels.py / test_gpu_attention.py / test_gpu_linear.py).  The yardstick is MEASURED: the reference's own
execution mode (oracle U-Net + oracle hooks, fp16 weights, CUDA, torch.autocast(fp16)) deviates from the fp32 oracle by
1.16e-2 (custom) / 1.17e-2 (LoRA) of max|latent| after the 10-step run of tests/test_gpu_parity.py; the product lands at
7.6e-3 / 7.2e-3 in fp16 and 6.8e-2 / 7.0e-2 in bf16 (B200, profiles/r02b_pytest_gpu.txt).  Stated bounds, relative to
max|oracle output|:
  U-Net forward (tiny width, 70 transformer blocks): fp16 <= 1.5e-2, bf16 <= 6e-2; hooked / routed forward: same
  10-step sampler latent (max-abs and L2, relative): fp16 <= 2.5e-2 (= 2x the reference's own fp16 deviation),
  bf16 <= 1.5e-1 (13x: the bf16 mantissa is 3 bits shorter; the seeded random U-Net amplifies rounding more than real
  SDXL weights would, and the fp32-oracle comparison is the honest one — SURVEY §7 'parity budget')
CUDA-graph replay vs eager: bit-identical.  GEGLU kernel: one output rounding.
"""
import argparse

import pytest
import torch
import torch.nn.functional as F

import test_host_logic as T
from oracle import synth
from oracle.hooks_ref import make_lora_set, register_custom_ref, register_lora_ref, register_time_ref

pytestmark = pytest.mark.gpu

K = 3
TOL_FWD = {torch.float16: 1.5e-2, torch.bfloat16: 6e-2}
TOL_LOOP = {torch.float16: 2.5e-2, torch.bfloat16: 1.5e-1}


def _build():
    from tweediemix_b200 import build
    build.build()


def _gpu_unet(ref_unet, dtype):
    _build()
    return T.product_unet(ref_unet).to("cuda", dtype).finalize()


def _rel(got, want):
    return (got.float().cpu() - want).abs().max().item() / want.abs().max().item()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_geglu_kernel(dtype):
    _build()
    from tweediemix_b200 import ops
    g = torch.Generator().manual_seed(0)
    for shape in [(4, 1024, 2 * 5120), (3, 7, 16), (1, 4096, 2 * 2560)]:
        x = (torch.randn(shape, generator=g) * 2).to(dtype).cuda()
        h, gate = x.float().chunk(2, dim=-1)
        want = (h * F.gelu(gate)).to(dtype)
        got = ops.geglu(x)
        torch.testing.assert_close(got.float(), want.float(), rtol=2 ** -7 if dtype == torch.bfloat16 else 2 ** -10, atol=1e-3)
    with pytest.raises(RuntimeError, match="multiple of 8"):
        ops.geglu(torch.zeros(2, 12, dtype=dtype).cuda())


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_unet_forward_vs_oracle(dtype):
    ref = synth.make_base_unet(T.RCFG, 1)
    prod = _gpu_unet(ref, dtype)
    x, E, cond = T._inputs(4, hw=32)
    want = ref(x, 781, E, cond)["sample"]
    cc = {k: v.cuda() for k, v in cond.items()}
    got = prod(x.cuda(), 781, E.cuda(), cc)["sample"]
    assert got.dtype == dtype and got.shape == want.shape and got.is_contiguous()
    assert _rel(got, want) <= TOL_FWD[dtype]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("variant", ["custom", "lora"])
@pytest.mark.parametrize("t,batch", [(781, 4), (981, 4), (761, 2)])
def test_hooked_forward_vs_oracle(variant, t, batch, dtype):
    ref = synth.make_base_unet(T.RCFG, 1)
    lora = variant == "lora"
    extra = [make_lora_set(ref, 20 + i, up_std=0.05) for i in range(K)] if lora else \
            [synth.make_concept_unet(ref, 10 + i) for i in range(K)]
    s = T._product_sampler(ref, extra, lora, 50, 256)            # builds donors in the product's format
    _build()
    prod = s.unet.to("cuda", dtype).finalize()
    window = torch.tensor([781, 761, 741])
    hooks = s.hooks
    hooks.register_attention_control_efficient(s, window, K)
    hooks.register_time(s, t)
    (register_lora_ref if lora else register_custom_ref)(ref, extra, window, K)
    register_time_ref(ref, t, lora=lora)
    x, E, cond = T._inputs(batch, hw=32)
    want = ref(x, t, E, cond)["sample"]
    # un-routed control: with routing active the routed rows must differ from the base model
    got = prod(x.cuda(), t, E.cuda(), {k: v.cuda() for k, v in cond.items()})["sample"]
    assert _rel(got, want) <= TOL_FWD[dtype]


def _gpu_sampler(ref_unet, extra, lora, n, res, dtype, graphs):
    from tweediemix_b200.fusion_sampling import FusionComponents, Tweediemix
    s = T._product_sampler(ref_unet, extra, lora, n, res)
    comp = FusionComponents(unet=s.unet.to("cuda", dtype).finalize(),
                            concept_unets=[getattr(s, f"unet_{i}") for i in range(K)],
                            text_embeds=s.text_embeds, text_embeds_single=s.text_embeds_single, masks=s.masks)
    return Tweediemix(T._namespace(n, res, lora), comp, variant="lora" if lora else "custom", use_cuda_graphs=graphs)


@pytest.mark.parametrize("lora", [False, True])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_sampler_vs_oracle_and_graph_equals_eager(lora, dtype):
    _build()
    n, res = 10, 256
    ref_unet = synth.make_base_unet(T.RCFG, 1)
    extra = [make_lora_set(ref_unet, 20 + i, up_std=0.05) for i in range(K)] if lora else \
            [synth.make_concept_unet(ref_unet, 10 + i) for i in range(K)]
    orc = T._oracle_sampler(ref_unet, extra, lora, n, res)
    x0 = orc.initial_latent()
    want = orc.sample_loop(x0.clone())
    outs = []
    for graphs in (False, True):
        s = _gpu_sampler(ref_unet, extra, lora, n, res, dtype, graphs)
        s.init_fusion(int(n * 0.2), int(n * 0.8)) if lora else s.init_fusion(int(n * 0.2))
        got = s.sample_loop(x0.clone())
        assert s.n_forward_rows == orc.n_forward_rows
        outs.append(got.cpu())
    assert torch.equal(outs[0], outs[1]), "CUDA-graph replay must be bit-identical to eager"
    assert torch.isfinite(outs[1]).all()
    rel_max = (outs[1] - want).abs().max().item() / want.abs().max().item()
    rel_l2 = ((outs[1] - want).norm() / want.norm()).item()
    print(f"sampler parity lora={lora} {dtype}: rel_max={rel_max:.3e} rel_l2={rel_l2:.3e}")
    assert rel_max <= TOL_LOOP[dtype] and rel_l2 <= TOL_LOOP[dtype], (rel_max, rel_l2)


def test_second_image_reuses_graphs_with_new_text():
    """set_text(): new prompts are written into the static buffers and the cached cross-attention K/V
    are re-projected in place, so replayed graphs see them (compare with a fresh sampler)."""
    _build()
    n, res, dtype = 5, 256, torch.bfloat16
    ref_unet = synth.make_base_unet(T.RCFG, 1)
    extra = [synth.make_concept_unet(ref_unet, 10 + i) for i in range(K)]
    text2, single2 = synth.make_text(T.RCFG, K, 999)
    a = _gpu_sampler(ref_unet, extra, False, n, res, dtype, True)
    a.init_fusion(1)
    x0 = torch.randn(1, 4, res // 8, res // 8, generator=torch.Generator().manual_seed(1))
    first = a.sample_loop(x0.clone()).cpu()
    a.set_text(tuple(t.cuda() for t in text2), tuple(t.cuda() for t in single2))
    second = a.sample_loop(x0.clone()).cpu()
    b = _gpu_sampler(ref_unet, extra, False, n, res, dtype, True)
    b.text_embeds, b.text_embeds_single = text2, single2
    b.init_fusion(1)
    fresh = b.sample_loop(x0.clone()).cpu()
    assert torch.equal(second, fresh)
    assert not torch.equal(first, second)


@pytest.mark.parametrize("impl", ["tmx", "cublas"])
@pytest.mark.parametrize("lora", [False, True])
def test_sampler_under_other_gemm_policies(monkeypatch, lora, impl):
    """The default policy is 'auto' (k10 for the GEGLU and LoRA-tail GEMMs, cuBLAS for the plain projections); 'tmx' (all in
    k10) and 'cublas' (all in the library) must hold the same parity bound."""
    _build()
    from tweediemix_b200 import ops
    monkeypatch.setattr(ops, "GEMM_IMPL", impl)
    n, res, dtype = 10, 256, torch.float16
    ref_unet = synth.make_base_unet(T.RCFG, 1)
    extra = [make_lora_set(ref_unet, 20 + i, up_std=0.05) for i in range(K)] if lora else \
            [synth.make_concept_unet(ref_unet, 10 + i) for i in range(K)]
    orc = T._oracle_sampler(ref_unet, extra, lora, n, res)
    x0 = orc.initial_latent()
    want = orc.sample_loop(x0.clone())
    s = _gpu_sampler(ref_unet, extra, lora, n, res, dtype, True)
    s.init_fusion(int(n * 0.2), int(n * 0.8)) if lora else s.init_fusion(int(n * 0.2))
    got = s.sample_loop(x0.clone()).cpu()
    rel = (got - want).abs().max().item() / want.abs().max().item()
    print(f"sampler parity lora={lora} fp16 TMX_GEMM={impl}: rel_max={rel:.3e}")
    assert s.n_forward_rows == orc.n_forward_rows and rel <= TOL_LOOP[dtype]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_cat_free_up_blocks_gpu(monkeypatch, dtype):
    """TMX_CAT_FREE=1 (two-source GroupNorm + split shortcut GEMMs in the up blocks) vs the default torch.cat path and the fp32 oracle."""
    from tweediemix_b200 import unet as U
    ref = synth.make_base_unet(T.RCFG, 1)
    prod = _gpu_unet(ref, dtype)
    x, E, cond = T._inputs(4, hw=32)
    want = ref(x, 781, E, cond)["sample"]
    cc = {k: v.cuda() for k, v in cond.items()}
    base = prod(x.cuda(), 781, E.cuda(), cc)["sample"]
    monkeypatch.setattr(U, "CAT_FREE", True)
    got = prod(x.cuda(), 781, E.cuda(), cc)["sample"]
    assert _rel(got, want) <= TOL_FWD[dtype]
    assert (got.float() - base.float()).abs().max().item() <= 0.5 * TOL_FWD[dtype] * want.abs().max().item()
