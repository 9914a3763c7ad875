"""CPU model of the stream-K schedule of the attention kernel (csrc/attention.cu: `next_step`, the split-unit dump / merge protocol and the
host-side grid / cost-model choice in `tmx_attn_fwd`).  The CUDA code cannot run here; this restates its index arithmetic in Python and checks
the properties the kernel relies on, over the SDXL shapes and a sweep of ragged ones:

  * every (unit, K/V tile) iteration is executed exactly once, by exactly one CTA, and both schedules cover every query tile;
  * a CTA dumps at most ONE partial piece (so one workspace slot per CTA and slot suffices) and it is the first step of its range;
  * every dumped piece is folded by exactly one owner — the CTA that holds the head (j0 == 0) of the unit — and the owner's loop
    `for k = c + 1; begin(k) < unit_end` (split_merge) visits exactly the CTAs that dumped pieces of that unit, in CTA order;
  * owners only wait on CTAs with a HIGHER index whose piece is their FIRST step (no circular wait);
  * the constants are the ones in the source (kept in sync by parsing them).
"""
import os
import re

import pytest

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tweediemix_b200", "csrc", "attention.cu")
KBM = KBN = 128
SMS = 148


def _consts():
    src = open(SRC).read()
    m = re.search(r"constexpr float kStepCost = ([0-9.]+)f, kLoneCost = ([0-9.]+)f, kMergeCost = ([0-9.]+)f;", src)
    assert m, "cost-model constants not found in attention.cu"
    return tuple(float(v) for v in m.groups())


def host_plan(B, H, Nq, Nk, sms=SMS, split_mode=1, tpu=2):
    """tmx_attn_fwd: (split, grid, total) — total = schedule space size (iterations if split else query tiles)."""
    step_c, lone_c, merge_c = _consts()
    QT = -(-Nq // KBM)
    T = -(-Nk // KBN)
    total_tiles = B * H * QT
    UPP = -(-QT // tpu)
    units = B * H * UPP
    split = 0
    if T > 1 and split_mode != 0:
        if split_mode == 2:
            split = 1
        else:
            unitsA = (total_tiles + 1) // 2 if tpu == 2 else total_tiles
            gridA = min(unitsA, sms)
            nt = -(-total_tiles // gridA)
            costA = ((nt // 2 + lone_c * (nt & 1)) * T + (nt // 2 + (nt & 1)) * step_c) if tpu == 2 else nt * (T + step_c)
            wantC = max(units * T // 2, units)
            gridC = min(wantC, sms)
            itersC = units * T / gridC
            costC = itersC + (itersC / T + 1.0) * step_c + merge_c
            split = 1 if costC < costA else 0
    iters = units * T if split else total_tiles
    want = max(iters // 2, units) if split else ((total_tiles + 1) // 2 if tpu == 2 else total_tiles)
    grid = min(want, sms)
    return dict(split=split, grid=grid, total=iters, QT=QT, T=T, UPP=UPP, TPU=tpu, units=units, H=H)


def cta_steps(p, c):
    """next_step() walked over CTA c's range: list of (unit or first tile id, tiles covered, j0, j1)."""
    begin, end = c * p["total"] // p["grid"], (c + 1) * p["total"] // p["grid"]
    it, out = begin, []
    T, QT, UPP, TPU = p["T"], p["QT"], p["UPP"], p["TPU"]
    while it < end:
        if p["split"]:
            u = it // T
            j0 = it - u * T
            j1 = j0 + (end - it) if (end - it) < T - j0 else T
            qt = (u % UPP) * TPU
            nslots = 2 if (TPU == 2 and qt + 1 < QT) else 1
            bh = u // UPP
            tiles = [bh * QT + qt + w for w in range(nslots)]
            out.append((u, tiles, j0, j1))
            it += j1 - j0
        else:
            nslots = 2 if (TPU == 2 and it + 1 < end and it % QT + 1 < QT) else 1
            out.append((it, [it + w for w in range(nslots)], 0, T))
            it += nslots
    return out


def check_schedule(B, H, Nq, Nk, split_mode):
    p = host_plan(B, H, Nq, Nk, split_mode=split_mode)
    T, QT = p["T"], p["QT"]
    assert 1 <= p["grid"] <= SMS and p["grid"] <= p["total"]                  # every CTA has work; all CTAs co-resident
    seen = {}                                                                   # (tile, kv tile) -> count
    dumps, owners = {}, {}
    for c in range(p["grid"]):
        steps = cta_steps(p, c)
        assert steps, "a CTA without work"
        n_dumps = 0
        for si, (u, tiles, j0, j1) in enumerate(steps):
            assert 0 <= j0 < j1 <= T
            for t in tiles:
                assert t // QT == tiles[0] // QT                                # both slots in the same (b, h)
                for j in range(j0, j1):
                    seen[(t, j)] = seen.get((t, j), 0) + 1
            if p["split"]:
                if j0 > 0:                                                      # partial piece handed to the owner
                    n_dumps += 1
                    assert si == 0, "a dumped piece must be the CTA's first step (owners rely on it being ready early)"
                    dumps.setdefault(u, []).append(c)
                elif j1 < T:                                                    # owner's head piece
                    assert si == len(steps) - 1, "the owner's head piece is the last step of its range"
                    owners[u] = c
        assert n_dumps <= 1                                                     # one workspace slot per CTA
    total_tiles = B * H * QT
    assert len(seen) == total_tiles * T and all(v == 1 for v in seen.values())  # exactly-once coverage
    if p["split"]:
        assert set(dumps) == set(owners)
        for u, c in owners.items():
            unit_end = (u + 1) * T
            k, visited = c + 1, []
            while k < p["grid"] and k * p["total"] // p["grid"] < unit_end:     # split_merge's loop
                visited.append(k)
                k += 1
            assert visited == dumps[u] and all(k > c for k in visited)
    return p


SDXL = [(4, 10, 4096, 4096), (4, 20, 1024, 1024), (2, 10, 4096, 4096), (2, 20, 1024, 1024), (1, 10, 4096, 4096), (1, 20, 1024, 1024),
        (5, 20, 1024, 1024), (5, 10, 4096, 4096)]
RAGGED = [(1, 5, 200, 333), (1, 3, 300, 700), (2, 2, 128, 2000), (1, 1, 256, 4096), (3, 7, 129, 129), (1, 1, 1, 129), (2, 3, 1000, 257),
          (1, 1, 128, 128 * 40), (7, 1, 640, 640)]


@pytest.mark.parametrize("split_mode", [0, 1, 2])
@pytest.mark.parametrize("shape", SDXL + RAGGED)
def test_schedule_covers_every_iteration_once(shape, split_mode):
    check_schedule(*shape, split_mode=split_mode)


def test_cost_model_choices_at_the_sdxl_shapes():
    """What the host picks (measured in profiles/r02p_kbench_attn_schedules.txt): stream-K wherever whole tiles quantise badly on 148 SMs,
    the whole-tile schedule for one U-Net row at N = 1024 (80 units on 148 SMs: the merge costs more than the half-empty round)."""
    pick = lambda *s: host_plan(*s)["split"]
    assert pick(4, 10, 4096, 4096) == 1 and pick(2, 10, 4096, 4096) == 1 and pick(1, 10, 4096, 4096) == 1
    assert pick(2, 20, 1024, 1024) == 1
    assert pick(1, 20, 1024, 1024) == 0
    assert pick(4, 20, 1024, 77) == 0                                            # a single K/V tile: nothing to split


def test_one_kv_tile_never_splits_and_grid_never_exceeds_units():
    for shape in [(4, 20, 1024, 77), (1, 1, 128, 1), (2, 3, 300, 128)]:
        for mode in (1, 2):
            p = check_schedule(*shape, split_mode=mode)
            assert p["split"] == 0


# ------------------------------------------------------------------------------------------------ merge arithmetic (split_merge)
def _piece(q, k, v, j0, j1, scale_log2, thr=8.0, dtype=None):
    """One piece of a unit as the softmax warps compute it: K/V tiles [j0, j1), lazy reference maximum (moved only when the tile
    maximum exceeds it by more than `thr`), P rounded to the I/O dtype before P V, l accumulated in fp32 from the unrounded P.
    Returns (O un-normalised, m_ref, l) per row."""
    import torch
    n = q.shape[0]
    m = torch.full((n,), float("-inf"))
    l = torch.zeros(n)
    o = torch.zeros(n, v.shape[1])
    for j in range(j0, j1):
        s = (q @ k[j * KBN:(j + 1) * KBN].T) * scale_log2
        mt = s.max(dim=1).values
        bump = mt > m + thr
        m_new = torch.where(bump, mt, m)
        alpha = torch.where(bump, torch.exp2(m - m_new), torch.ones(n))
        alpha = torch.where(torch.isinf(m), torch.zeros(n), alpha)
        p = torch.exp2(s - m_new[:, None])
        l = l * alpha + p.sum(dim=1)
        pr = p.to(dtype).float() if dtype is not None else p
        o = o * alpha[:, None] + pr @ v[j * KBN:(j + 1) * KBN]
        m = m_new
    return o, m, l


@pytest.mark.parametrize("cuts", [(3,), (1, 2), (2, 5, 6), (1, 2, 3, 4, 5, 6, 7)])
def test_merging_pieces_equals_the_whole_softmax(cuts):
    """The owner's fold (split_merge: m_new = max, O = O * 2^(m - m_new) + O_k * 2^(m_k - m_new), same for l, pieces in CTA order)
    over un-normalised partials with DIFFERENT lazy reference maxima reproduces softmax(Q K^T) V."""
    import torch
    torch.manual_seed(0)
    T, d = 8, 64
    q = torch.randn(128, d)
    k = torch.randn(T * KBN, d) * 2.0                       # peaked rows: the pieces see different maxima, some rescale, some do not
    v = torch.randn(T * KBN, d)
    sl2 = 0.125 * 1.4426950408889634
    want = torch.softmax((q @ k.T) * 0.125, dim=1) @ v
    for dtype, tol in ((None, 2e-5), (torch.bfloat16, 2e-2), (torch.float16, 3e-3)):
        bounds = [0, *cuts, T]
        o, m, l = _piece(q, k, v, bounds[0], bounds[1], sl2, dtype=dtype)            # the owner's head piece
        for a, b in zip(bounds[1:-1], bounds[2:]):                                  # the following CTAs' pieces, in order
            ok, mk, lk = _piece(q, k, v, a, b, sl2, dtype=dtype)
            m_new = torch.maximum(m, mk)
            a_own, a_k = torch.exp2(m - m_new), torch.exp2(mk - m_new)
            l = l * a_own + lk * a_k
            o = o * a_own[:, None] + ok * a_k[:, None]
            m = m_new
        got = o / l[:, None]
        assert (got - want).abs().max().item() <= tol
