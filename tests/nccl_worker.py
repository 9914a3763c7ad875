"""Worker of tests/test_gpu_nccl.py — launched as `python -m torch.distributed.run --nproc-per-node 2 tests/nccl_worker.py OUT`.

Checks, on real GPUs over NCCL:
  (1) kernel level: tmx_blend_partial_fwd on each rank's rows -> ncclAllReduce(sum, fp32) -> tmx_blend_finish_fwd equals the
      single-GPU fused k7 kernel on the same eps within 2e-4 (fp32 round-off of the linear form), bit-identical across ranks;
  (2) sampler level: a 10-step custom + LoRA run of the narrow SDXL-topology config, concept-parallel over the 2 ranks, against
      the SAME model run unsharded on each rank: bit-identical across ranks; vs unsharded <= 3e-2 relative (the U-Net sees 2-row
      instead of 4-row batches, so cuBLAS / cuDNN may round differently — the blend itself is pinned by (1)).
Rank 0 writes a JSON report to OUT.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def main(out_path):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from tweediemix_b200 import build, ops
    if rank == 0:
        build.build()
    dist.barrier()
    report = {"world": world}

    # ---- (1) kernel level
    from tweediemix_b200.fusion_sampling import assign_rows
    K, h, w = 3, 128, 128
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4, h, w, generator=g).to(dev)
    eps = torch.randn(1, K + 1, 4, h, w, generator=g).to(torch.bfloat16).to(dev)
    from tweediemix_b200.masks import stripe_masks
    masks = stripe_masks(K, h, w, device=dev)
    a_t, a_n, gs = 0.043827, 0.051787, 0.8
    want = ops.tweedie_blend_ddim(x, eps, masks, a_t, a_n, gs)
    mine = assign_rows(K + 1, world, rank)
    acc = torch.empty(1, 2, 4, h, w, device=dev)
    ops.blend_partial(eps[:, mine[0]:mine[-1] + 1].contiguous() if mine else None, masks, mine, acc, 1, K=K)
    dist.all_reduce(acc)
    got = ops.blend_finish(x, acc, masks, a_t, a_n, gs, K=K)
    gathered = [torch.empty_like(got) for _ in range(world)]
    dist.all_gather(gathered, got)
    report["kernel_max_abs_vs_single_gpu"] = float((got - want).abs().max().item())
    report["kernel_ranks_bit_identical"] = all(torch.equal(gathered[0], t) for t in gathered[1:])
    assert report["kernel_max_abs_vs_single_gpu"] <= 2e-4 and report["kernel_ranks_bit_identical"], report

    # ---- (2) sampler level
    import test_host_logic as T
    from oracle import synth
    from oracle.hooks_ref import make_lora_set
    from tweediemix_b200.fusion_sampling import FusionComponents, Tweediemix, make_concept_groups
    n, res = 10, 256
    for lora in (False, True):
        ref_unet = synth.make_base_unet(T.RCFG, 1)
        extra = [make_lora_set(ref_unet, 20 + i, up_std=0.05) for i in range(T.K)] if lora else \
                [synth.make_concept_unet(ref_unet, 10 + i) for i in range(T.K)]
        x0 = torch.randn(1, 4, res // 8, res // 8, generator=torch.Generator().manual_seed(3))
        outs = {}
        for sharded in (True, False):
            s = T._product_sampler(ref_unet, extra, lora, n, res)
            comp = FusionComponents(unet=s.unet.to(dev, torch.float16).finalize(), concept_unets=[getattr(s, f"unet_{i}") for i in range(T.K)],
                                    text_embeds=s.text_embeds, text_embeds_single=s.text_embeds_single, masks=s.masks)
            pg = None
            if sharded:
                _, _, _, pg = make_concept_groups(world, rank, T.K + 1)
            m = Tweediemix(T._namespace(n, res, lora), comp, variant="lora" if lora else "custom", use_cuda_graphs=True, process_group=pg)
            m.init_fusion(int(n * 0.2), int(n * 0.8)) if lora else m.init_fusion(int(n * 0.2))
            outs[sharded] = m.sample_loop(x0.clone())
            if sharded:
                rows = torch.tensor([float(m.n_forward_rows)], device=dev)
                dist.all_reduce(rows)
            else:
                assert int(rows.item()) == m.n_forward_rows, (rows.item(), m.n_forward_rows)      # every row computed exactly once
        gathered = [torch.empty_like(outs[True]) for _ in range(world)]
        dist.all_gather(gathered, outs[True])
        ident = all(torch.equal(gathered[0], t) for t in gathered[1:])
        rel = float((outs[True] - outs[False]).abs().max().item() / outs[False].abs().max().item())
        tag = "lora" if lora else "custom"
        report[f"sampler_{tag}_rel_vs_unsharded"] = rel
        report[f"sampler_{tag}_ranks_bit_identical"] = ident
        assert ident and rel <= 3e-2, report
    if rank == 0:
        with open(out_path, "w") as fh:
            json.dump(report, fh)
        print("NCCL parity:", json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
