"""Oracle schedule vs the table in SURVEY App. C (pure arithmetic; exact to fp32 print precision)."""
import pytest
import torch

from oracle.schedule import make_schedule

TABLE_50 = [  # step, t, alpha(t), next_t, alpha(next_t), sqrt(1-a)/sqrt(a)
    (0, 981, 0.005844, 961, 0.007365, 13.043),
    (9, 801, 0.036870, 781, 0.043827, 5.111),
    (10, 781, 0.043827, 761, 0.051787, 4.671),
    (40, 181, 0.784153, 161, 0.813550, 0.525),
    (48, 21, 0.981314, 1, 0.999150, 0.138),
    (49, 1, 0.999150, -19, 0.999150, 0.029),
]
TABLE_5 = [
    (0, 801, 0.036870, 601, 0.160782, 5.111),
    (1, 601, 0.160782, 401, 0.424483, 2.285),
    (4, 1, 0.999150, -199, 0.999150, 0.029),
]


@pytest.mark.parametrize("n,table", [(50, TABLE_50), (5, TABLE_5)])
def test_schedule_table(n, table):
    s = make_schedule(n)
    assert s.skip == 1000 // n
    assert len(s.timesteps) == n
    for step, t, a, nt, an, ratio in table:
        assert int(s.timesteps[step]) == t
        assert int(s.timesteps[step]) - s.skip == nt
        assert float(s.alpha(t)) == pytest.approx(a, abs=6e-7)
        assert float(s.alpha(nt)) == pytest.approx(an, abs=6e-7)
        at = s.alpha(t)
        assert float((1 - at).sqrt() / at.sqrt()) == pytest.approx(ratio, abs=6e-4)


def test_shift_and_final_alpha():
    s = make_schedule(50)
    assert s.alphas_cumprod.shape == (1001,)
    assert float(s.alphas_cumprod[0]) == 1.0                       # the prepended 1.0
    assert float(s.final_alpha_cumprod) == pytest.approx(0.99914998, abs=1e-7)
    assert torch.equal(s.alpha(-19), s.final_alpha_cumprod)
    assert torch.equal(s.alpha(1), s.alphas_cumprod[1])            # alpha(t) = orig[t-1]
    assert s.timesteps[0] == 981 and s.timesteps[-1] == 1
    assert s.alpha(5).dtype == torch.float32 and s.alpha(5).ndim == 0
