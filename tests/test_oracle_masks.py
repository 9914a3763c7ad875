"""Mask ingest restatement on the reference's own example_results mask JPEGs (SURVEY App. D stats)."""
import pytest
import torch

from oracle import synth

COVER = {"test_out": (0.1432, 0.2055, 0.6513), "test_out_lora": (0.1387, 0.1952, 0.6661),
         "test_out_panda": (0.2308, 0.2361, 0.5330), "test_out_woman": (0.1730, 0.2880, 0.5390)}


@pytest.mark.parametrize("which", sorted(COVER))
def test_fixture_masks_partition(which):
    m = synth.fixture_masks(128, 128, which)
    assert m.shape == (3, 1, 128, 128) and m.dtype == torch.float32
    assert set(m.unique().tolist()) <= {0.0, 1.0}
    assert torch.equal(m.sum(0), torch.ones(1, 128, 128))            # perfect partition, no overlap
    cover = [float(m[c].mean()) for c in range(3)]
    for got, want in zip(cover, COVER[which]):
        assert got == pytest.approx(want, abs=6e-5)


def test_small_grid_and_stripes():
    assert synth.fixture_masks(32, 32).shape == (3, 1, 32, 32)
    s = synth.stripe_masks(8, 16, 20)
    assert torch.equal(s.sum(0), torch.ones(1, 16, 20))
