"""si_sdr against the doctest values the reference ships (disco_theque/metrics.py:355-372)."""
import numpy as np
import torch

from disco_b200.post import si_sdr, snr_db


def test_si_sdr_reference_doctests():
    np.random.seed(0)
    reference = np.random.randn(100)
    assert torch.isinf(si_sdr(reference, reference * 2))
    assert abs(si_sdr(reference, np.flip(reference).copy()).item() - (-25.127672346460717)) < 1e-9
    assert abs(si_sdr(reference, reference + np.flip(reference)).item() - 0.481070445785553) < 1e-9
    assert abs(si_sdr(reference, reference + 0.5).item() - 6.3704606032577304) < 1e-9
    assert abs(si_sdr(reference, reference * 2 + 1).item() - 6.3704606032577304) < 1e-9
    assert torch.isnan(si_sdr(np.array([1.0, 0.0]), np.array([0.0, 0.0])))
    both = si_sdr(np.stack([reference, reference]), np.stack([reference * 2 + 1, reference + 0.5]))
    assert np.allclose(both.numpy(), [6.3704606, 6.3704606], atol=1e-6)
    assert abs(snr_db(np.ones(10) * 2.0, np.ones(10)).item() - 10 * np.log10(4.0)) < 1e-12
