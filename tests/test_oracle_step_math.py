"""Known-answer and property tests for the step arithmetic restatement (oracle/step_math.py)."""
import math

import pytest
import torch

from oracle import step_math as sm


def test_known_answer_scalar():
    # one pixel, K=2, hand-computed
    x = torch.full((1, 1, 1, 1), 2.0)
    eps = torch.tensor([1.0, 3.0, -1.0]).reshape(3, 1, 1, 1)       # uncond, c0, c1
    masks = torch.tensor([1.0, 0.0]).reshape(2, 1, 1, 1)
    at, at_next, g = 0.25, 0.64, 0.5
    # eps_c0 = 1 + .5*(3-1) = 2 ; x0 = (2 - sqrt(.75)*2)/0.5
    x0_expect = (2 - math.sqrt(0.75) * 2) / 0.5
    xn_expect = 0.8 * x0_expect + 0.6 * 1.0
    xn, x0 = sm.fused_step(x, eps, masks, at, at_next, g)
    assert float(x0) == pytest.approx(x0_expect, rel=1e-6)
    assert float(xn) == pytest.approx(xn_expect, rel=1e-6)
    xl, _ = sm.fused_step(x, eps, masks, at, at_next, g, is_last=True)
    assert float(xl) == pytest.approx(x0_expect, rel=1e-6)


def test_dtype_semantics_match_reference_promotion():
    """App. B: 0-dim fp32 alpha does not promote fp16 eps; subtraction from fp32 x does."""
    at = torch.tensor(0.3)
    eps = torch.randn(4, 4, 8, 8).half()
    x = torch.randn(1, 4, 8, 8)
    assert ((1 - at).sqrt() * eps).dtype == torch.float16
    assert sm.cfg_combine(eps[:1], eps[1:2], 0.8).dtype == torch.float16
    assert sm.tweedie_x0(x, eps[:1], at).dtype == torch.float32
    xn, x0 = sm.fused_step(x, eps, torch.ones(3, 1, 8, 8) / 3, at, torch.tensor(0.4), 0.8)
    assert xn.dtype == torch.float32 and x0.dtype == torch.float32


def test_partition_masks_reduce_to_single_cfg():
    """With a partition and identical concept rows the blend equals a plain CFG step."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4, 8, 8, generator=g)
    eu, ec = torch.randn(1, 4, 8, 8, generator=g), torch.randn(1, 4, 8, 8, generator=g)
    masks = torch.zeros(3, 1, 8, 8)
    masks[0, :, :, :3], masks[1, :, :, 3:5], masks[2, :, :, 5:] = 1, 1, 1
    a, b = sm.fused_step(x, torch.cat([eu, ec, ec, ec]), masks, 0.2, 0.3, 0.8)
    c, d = sm.cfg_step(x, torch.cat([eu, ec]), 0.2, 0.3, 0.8)
    torch.testing.assert_close(a, c, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(b, d, rtol=1e-6, atol=1e-6)


def test_linear_sharded_form_equals_direct_form():
    g = torch.Generator().manual_seed(1)
    K = 5
    x = torch.randn(1, 4, 16, 16, generator=g)
    eps = torch.randn(K + 1, 4, 16, 16, generator=g)
    masks = (torch.rand(K, 1, 16, 16, generator=g) > 0.5).float()     # overlapping, non-partition
    xn, x0 = sm.fused_step(x, eps, masks, 0.05, 0.06, 0.8)
    acc = sum(sm.blend_partial(eps[1 + c:2 + c], masks[c:c + 1]) for c in range(K))
    xn2, x02 = sm.blend_finish(x, acc, eps[:1], masks.sum(0, keepdim=True), 0.05, 0.06, 0.8)
    torch.testing.assert_close(x0, x02, rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(xn, xn2, rtol=2e-5, atol=2e-5)


def test_resample_and_renoise_identities():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, 4, 8, 8, generator=g)
    eps = torch.randn(4, 4, 8, 8, generator=g)
    K = 3
    x0 = sm.resample_x0(x, eps, 0.01, 0.8, K)
    man = 2 * sm.tweedie_x0(x, sm.cfg_combine(eps[:1], eps[1:2], 0.8), 0.01) \
        - sm.tweedie_x0(x, sm.cfg_combine(eps[:1], eps[2:3], 0.8), 0.01) \
        - sm.tweedie_x0(x, sm.cfg_combine(eps[:1], eps[3:4], 0.8), 0.01)
    torch.testing.assert_close(x0, man)
    # DDIM down then re-noise with the same eps (g=0 -> eps_u) is the identity on x
    e2 = torch.cat([eps[:1], eps[:1]])
    x0u = sm.tweedie_x0(x, eps[:1], 0.3)
    low = sm.ddim_update(x0u, eps[:1], 0.5)
    back = sm.renoise(low, e2, 0.3, 0.5, 0.0)
    torch.testing.assert_close(back, x, rtol=1e-5, atol=1e-5)
