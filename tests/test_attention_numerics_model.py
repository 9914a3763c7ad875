"""CPU model of the attention kernel's arithmetic (csrc/attention.cu), with the constants parsed out of the .cu file so the
test cannot drift from the kernel: (1) the FMA-pipe exp2 (Cody-Waite split + degree-3 polynomial) stays within the error the
header claims over the whole range the kernel feeds it; (2) online softmax with LAZY rescaling (reference max moved only when
it grows by more than the threshold) and P rounded to 16 bits reproduces softmax(QK^T)V to the tolerance the GPU test uses."""
import os
import re

import numpy as np
import pytest
import torch

SRC = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tweediemix_b200", "csrc", "attention.cu")).read()


def _consts():
    body = SRC[SRC.index("ex2_poly2("):]
    body = body[:body.index("e1 = __int_as_float")]
    c = [float(x) for x in re.findall(r"pk2\((-?\d+\.\d*(?:e-?\d+)?)f, ", body)]
    # order in the kernel: magic, -magic, -1, c3, c2, c1, c0
    assert c[0] == 12582912.0 and c[1] == -12582912.0 and c[2] == -1.0 and len(c) == 7, c
    thr = float(re.search(r"kRescaleThreshold = (\d+\.\d+)f", SRC).group(1))
    every = int(re.search(r"#define TMX_ATTN_POLY_EVERY (\d+)", SRC).group(1))
    return c[3:], thr, every


def ex2_poly(x, coef):
    """fp32 emulation of ex2_poly2: clamp, round-to-nearest split via the 1.5*2^23 magic number, Horner, exponent add."""
    c3, c2, c1, c0 = [np.float32(v) for v in coef]
    x = np.maximum(x.astype(np.float32), np.float32(-126.0))
    magic = np.float32(12582912.0)
    xr = (x + magic).astype(np.float32)
    n = (xr - magic).astype(np.float32)
    f = (x - n).astype(np.float32)
    p = (f * c3 + c2).astype(np.float32)
    p = (p * f + c1).astype(np.float32)
    p = (p * f + c0).astype(np.float32)
    bits = p.view(np.int32) + (xr.view(np.int32) << 23)
    return bits.view(np.float32)


def test_polynomial_exp2_error_bound():
    coef, _, _ = _consts()
    x = np.concatenate([np.linspace(-120, 8.0, 2_000_001), np.array([-120.0, -0.5, 0.0, 0.5, 8.0])]).astype(np.float32)
    got = ex2_poly(x, coef).astype(np.float64)
    want = np.exp2(x.astype(np.float64))
    rel = np.abs(got / want - 1.0)
    assert rel.max() < 9e-5, rel.max()                       # header: 7.5e-5 (+ fp32 round-off of the Horner steps)
    # far below the rounding of P: 2^-9 (bf16) / 2^-12 (fp16)
    assert rel.max() < 2.0 ** -12 / 2
    # near the clamp (x -> -126, results ~1e-38, where p < 1 drops into the denormal encoding) accuracy is irrelevant but the
    # value must stay a harmless tiny non-negative number; inputs below the clamp (masked columns are -inf) give a harmless ~2^-126, never garbage from exponent wrap-around
    tiny = ex2_poly(np.array([-1e30, -np.inf, -500.0, -126.0, -125.7, -124.2], dtype=np.float32), coef)
    assert np.all(tiny >= 0) and np.all(tiny < 1e-37)


def _kernel_model(q, k, v, scale, thr, every, coef, p_dtype, tile=128):
    """One (b, h): online softmax over 128-column K/V tiles exactly as the softmax warps do it (log2 domain, lazy rescale,
    1 pair in `every` through the polynomial, P rounded to the I/O dtype before P V, fp32 accumulation)."""
    sl2 = np.float32(scale * 1.4426950408889634)
    Nq, Nk = q.shape[0], k.shape[0]
    m_ref = np.full(Nq, -np.inf, dtype=np.float32)
    l = np.zeros(Nq, dtype=np.float32)
    O = np.zeros((Nq, v.shape[1]), dtype=np.float32)
    for j0 in range(0, Nk, tile):
        S = (q @ k[j0:j0 + tile].T).astype(np.float32)                         # tcgen05 fp32 accumulator
        m_tile = S.max(axis=1) * sl2
        bump = m_tile > m_ref + np.float32(thr)
        m_new = np.where(bump, m_tile, m_ref).astype(np.float32)
        alpha = np.where(bump, np.exp2((m_ref - m_new).astype(np.float64)), 1.0).astype(np.float32)
        alpha[np.isnan(alpha)] = 0.0                                           # first tile: exp2(-inf - m) = 0
        l *= alpha
        O *= alpha[:, None]
        x = (S * sl2 - m_new[:, None]).astype(np.float32)
        P = np.exp2(x.astype(np.float64)).astype(np.float32)
        pair = (np.arange(x.shape[1]) // 2) % 16
        poly_cols = (pair % every) == every - 1 if every > 0 else np.zeros(x.shape[1], bool)
        P[:, poly_cols] = ex2_poly(x[:, poly_cols], coef)
        l += P.sum(axis=1, dtype=np.float32)
        Pr = torch.from_numpy(P).to(p_dtype).float().numpy()                   # P is handed to the tensor core in 16 bits
        O += Pr @ v[j0:j0 + tile]
        m_ref = m_new
    return O / l[:, None]


@pytest.mark.parametrize("p_dtype,tol", [(torch.bfloat16, 2e-2), (torch.float16, 4e-3)])
@pytest.mark.parametrize("peaked", [False, True])
def test_lazy_rescale_online_softmax_model(p_dtype, tol, peaked):
    coef, thr, every = _consts()
    g = np.random.default_rng(3)
    Nq, Nk, d = 64, 640, 64
    q = g.standard_normal((Nq, d)).astype(np.float32)
    k = g.standard_normal((Nk, d)).astype(np.float32)
    v = g.standard_normal((Nk, d)).astype(np.float32)
    if peaked:                                   # logits ~ +-60 with the largest ones in LATE tiles: the reference max moves late
        q *= 4.0
        k *= 4.0
        k[Nk // 2:] *= 1.5
    to16 = lambda a: torch.from_numpy(a).to(p_dtype).float().numpy()
    q, k, v = to16(q), to16(k), to16(v)
    got = _kernel_model(q, k, v, 0.125, thr, every, coef, p_dtype)
    want = torch.softmax(torch.from_numpy(q @ k.T).double() * 0.125, dim=-1).numpy() @ v.astype(np.float64)
    assert np.isfinite(got).all()
    assert np.abs(got - want).max() <= tol * (1.5 if peaked else 1.0)
    # the lazy threshold really skipped rescales (otherwise this test would not exercise the lazy path)
    assert thr >= 1.0
