"""CPU-only, world_size 2 over gloo: the concept-parallel path of the sampler (row assignment ->
per-rank U-Net rows -> blend partial -> ONE all-reduce -> blend finish) gives every rank the latent
the single-process sampler produces.  Kernels are the plain-PyTorch stand-ins of tests/fake_ops.py
(host logic only; the real kernels run the same path in tests/test_gpu_model.py)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, lora, out_dir, max_group=None, n_imgs=1, k=None):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fake_ops
        from tweediemix_b200 import ops
        for name in fake_ops.NAMES:
            setattr(ops, name, getattr(fake_ops, name))
        import test_host_logic as T
        from oracle import synth
        from oracle.hooks_ref import make_lora_set
        n, res = 5, 64
        k = T.K if k is None else k
        ref_unet = synth.make_base_unet(T.RCFG, 1)
        extra = [make_lora_set(ref_unet, 20 + i, up_std=0.05) for i in range(k)] if lora else \
                [synth.make_concept_unet(ref_unet, 10 + i) for i in range(k)]
        from tweediemix_b200.fusion_sampling import make_concept_groups
        gsize, n_groups, my_group, pg = make_concept_groups(world, rank, (k + 1) * n_imgs, max_group)
        s = T._product_sampler(ref_unet, extra, lora, n, res, pg=pg, k=k)
        s.init_fusion(int(n * 0.2), int(n * 0.8)) if lora else s.init_fusion(int(n * 0.2))
        torch.manual_seed(7 + my_group)                      # every image group samples its own image(s)
        x = s.sample_loop(torch.randn(n_imgs, 4, res // 8, res // 8))
        torch.save({"x": x, "rows": s.n_forward_rows, "group": my_group, "gsize": gsize, "n_groups": n_groups},
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("lora", [False, True])
def test_concept_parallel_two_ranks_equals_single(tmp_path, monkeypatch, lora):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), lora, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(tmp_path / f"rank{r}.pt") for r in range(world))
    assert torch.equal(r0["x"], r1["x"]), "ranks must end with bit-identical latents (no broadcast needed)"

    sys.path.insert(0, HERE)
    import fake_ops
    import test_host_logic as T
    from oracle import synth
    from oracle.hooks_ref import make_lora_set
    fake_ops.install(monkeypatch)
    n, res = 5, 64
    ref_unet = synth.make_base_unet(T.RCFG, 1)
    extra = [make_lora_set(ref_unet, 20 + i, up_std=0.05) for i in range(T.K)] if lora else \
            [synth.make_concept_unet(ref_unet, 10 + i) for i in range(T.K)]
    single = T._product_sampler(ref_unet, extra, lora, n, res)
    single.init_fusion(int(n * 0.2), int(n * 0.8)) if lora else single.init_fusion(int(n * 0.2))
    torch.manual_seed(7)
    want = single.sample_loop(torch.randn(1, 4, res // 8, res // 8))
    # the sharded path evaluates the CFG / blend in its linear form (fp32): equal up to round-off
    assert (r0["x"] - want).abs().max().item() < 1e-3      # fp32 round-off of the linear form (north-star bound)
    assert r0["rows"] + r1["rows"] == single.n_forward_rows      # every batch row computed exactly once


def test_two_image_groups_of_two_ranks(tmp_path, monkeypatch):
    """world_size 4 with groups capped at 2 ranks = the shape of the 8-GPU run (two 4-rank groups): each group
    block-distributes the K+1 rows of ITS image, all-reduces inside the group only, and reproduces the single-process
    latent of that image; the two groups never exchange data."""
    world = 4
    mp.spawn(_worker, args=(world, _free_port(), False, str(tmp_path), 2), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"rank{k}.pt") for k in range(world)]
    assert [x["group"] for x in r] == [0, 0, 1, 1] and all(x["gsize"] == 2 and x["n_groups"] == 2 for x in r)
    assert torch.equal(r[0]["x"], r[1]["x"]) and torch.equal(r[2]["x"], r[3]["x"])
    assert not torch.equal(r[0]["x"], r[2]["x"])

    sys.path.insert(0, HERE)
    import fake_ops
    import test_host_logic as T
    from oracle import synth
    fake_ops.install(monkeypatch)
    n, res = 5, 64
    ref_unet = synth.make_base_unet(T.RCFG, 1)
    extra = [synth.make_concept_unet(ref_unet, 10 + i) for i in range(T.K)]
    for g in range(2):
        single = T._product_sampler(ref_unet, extra, False, n, res)
        single.init_fusion(int(n * 0.2))
        torch.manual_seed(7 + g)
        want = single.sample_loop(torch.randn(1, 4, res // 8, res // 8))
        assert (r[2 * g]["x"] - want).abs().max().item() < 1e-3
        assert r[2 * g]["rows"] + r[2 * g + 1]["rows"] == single.n_forward_rows


@pytest.mark.parametrize("world,k,n_imgs", [(3, 3, 2), (2, 8, 1)])
def test_units_of_several_images_over_uneven_ranks(tmp_path, monkeypatch, world, k, n_imgs):
    """(image, prompt row) units split over ranks that do not divide them: 2 images x 4 rows over 3 ranks (3/3/2, rank 1
    holds rows of BOTH images) and K = 8 (9 rows over 2 ranks, 5/4) — the shape of BASELINE configs[3] (9 rows x 4 images
    over 8 ranks).  Every rank ends with the latents of the single-process run."""
    mp.spawn(_worker, args=(world, _free_port(), False, str(tmp_path), None, n_imgs, k), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"rank{q}.pt") for q in range(world)]
    for q in range(1, world):
        assert torch.equal(r[0]["x"], r[q]["x"])
    sys.path.insert(0, HERE)
    import fake_ops
    import test_host_logic as T
    from oracle import synth
    fake_ops.install(monkeypatch)
    n, res = 5, 64
    ref_unet = synth.make_base_unet(T.RCFG, 1)
    extra = [synth.make_concept_unet(ref_unet, 10 + i) for i in range(k)]
    single = T._product_sampler(ref_unet, extra, False, n, res, k=k)
    single.init_fusion(int(n * 0.2))
    torch.manual_seed(7)
    want = single.sample_loop(torch.randn(n_imgs, 4, res // 8, res // 8))
    assert (r[0]["x"] - want).abs().max().item() < 1e-3
    assert sum(x["rows"] for x in r) == single.n_forward_rows
    assert max(x["rows"] for x in r) - min(x["rows"] for x in r) <= single.n_forward_rows // (world * 4) + 5   # balanced


def test_make_concept_groups_single_process():
    from tweediemix_b200.fusion_sampling import make_concept_groups
    assert make_concept_groups(1, 0, 4) == (1, 1, 0, None)
