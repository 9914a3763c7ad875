"""CPU model of the per-(sample, group) GroupNorm plan (csrc/groupnorm.cu: `plan_slab`, the thread -> (pixel, vector) map of `gn_group_slab`):
which SDXL sites take the slab kernel, with what vector width / cluster size / block size, and that the map covers every vector of the
slab exactly once with each thread keeping ONE channel vector (its gamma / beta / add stay in registers)."""
import math
import os
import re

import pytest

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tweediemix_b200", "csrc", "groupnorm.cu")


def smem_max():
    m = re.search(r"constexpr size_t kGnSlabSmemMax = (\d+) \* 1024;", open(SRC).read())
    assert m
    return int(m.group(1)) * 1024


def plan_slab(N, C, HW, G=32, clusters=True):
    cpg = C // G
    if cpg % 2 or N > 65535 or N * G < 24:
        return None
    vl = 8 if cpg % 8 == 0 else (4 if cpg % 4 == 0 else 2)
    vpp = cpg // vl
    unit = vpp // math.gcd(vpp, 32) * 32
    if unit > 1024:
        return None
    cs = 1
    while cs <= 8:
        ppc = -(-HW // cs)
        smem = cpg * ppc * 2
        if smem <= smem_max():
            if cs > 1 and (not clusters or vl != 8 or (cs - 1) * ppc >= HW):
                return None
            threads = 1024 // unit * unit
            vecs = ppc * vpp
            while threads - unit >= 256 and (threads - unit) * 4 >= vecs:
                threads -= unit
            return dict(vl=vl, vpp=vpp, cs=cs, ppc=ppc, smem=smem, threads=threads, cpg=cpg)
        cs *= 2
    return None


SITES = {  # (C, HW): expected (vector length, cluster size) at batch 4, None = falls through to the cooperative / two-launch kernels
    (320, 16384): None, (640, 16384): None, (960, 16384): None,                 # 128x128: 4- / 8-byte runs, slab > one CTA
    (320, 4096): (2, 1), (640, 4096): (4, 1), (1280, 4096): (8, 2), (1920, 4096): None, (960, 4096): None,
    (640, 1024): (4, 1), (1280, 1024): (8, 1), (2560, 1024): (8, 1), (1920, 1024): (4, 1),
}


@pytest.mark.parametrize("site", sorted(SITES))
def test_sdxl_sites_take_the_expected_path(site):
    C, HW = site
    p = plan_slab(4, C, HW)
    want = SITES[site]
    assert (p is None) == (want is None)
    if p:
        assert (p["vl"], p["cs"]) == want
        assert p["smem"] <= smem_max() and p["threads"] % 32 == 0 and p["threads"] % p["vpp"] == 0 and p["threads"] <= 1024


@pytest.mark.parametrize("shape", [(4, 1280, 1024), (4, 640, 4096), (4, 1280, 4096), (2, 2560, 1024), (1, 1920, 64), (2, 640, 256), (5, 320, 4096),
                                   (3, 64, 35), (1, 1280, 4097)])
def test_thread_map_covers_the_slab_exactly_once(shape):
    N, C, HW = shape
    p = plan_slab(N, C, HW)
    if p is None:
        pytest.skip("shape not on the slab path")
    vpp, nt = p["vpp"], p["threads"]
    pstep = nt // vpp
    for rank in range(p["cs"]):
        pbeg, pend = rank * p["ppc"], min(HW, (rank + 1) * p["ppc"])
        assert pend > pbeg                                                       # no empty rank
        seen = set()
        for tid in range(nt):
            v = tid % vpp                                                        # ONE vector per thread for the whole kernel
            pix = pbeg + tid // vpp
            while pix < pend:
                key = (pix, v)
                assert key not in seen
                seen.add(key)
                assert (pix - pbeg) * vpp + v < p["ppc"] * vpp                   # inside this CTA's shared-memory slab
                pix += pstep
        assert len(seen) == (pend - pbeg) * vpp


def test_small_batches_and_odd_groups_fall_through():
    assert plan_slab(1, 320, 4096, G=32) is not None                             # 32 (n, g) pairs: enough CTAs
    assert plan_slab(1, 64, 64, G=8) is None                                     # 8 pairs: too few, other paths take it
    assert plan_slab(4, 96, 64, G=32) is None                                    # cpg = 3: no even vector
