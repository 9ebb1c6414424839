"""Host-side helpers of the compat layer against the reference's own outputs (tests/golden)."""
import numpy as np
import pytest

from conftest import load_golden


def test_list_helpers_match_reference():
    from disco_b200.compat import tango as ct
    g = load_golden("helpers_kat")
    ys, zs, zn = list(g["cat_y"]), list(g["cat_z"]), list(g["cat_zn"])
    for k in range(3):
        assert np.array_equal(ct.concatenate_signals(ys, zs, k), g["cat_%d" % k])
        assert np.array_equal(ct.concatenate_signals(ys, zs, k, g["cat_m"]), g["catm_%d" % k])
        assert np.array_equal(ct.get_z_for_mask(zs, zn, k, 3, "zs_hat"), g["zmask_zs_%d" % k])
        assert np.array_equal(ct.get_z_for_mask(zs, zn, k, 3, "zn_hat"), g["zmask_zn_%d" % k])
        assert np.array_equal(ct.get_z_for_mask(zs, zn, k, 3, ["zs_hat", "zn_hat"]), g["zmask_both_%d" % k])
    assert np.array_equal(ct.reshape_mask(g["reshape_in"], "last"), g["reshape_last"])
    assert np.array_equal(ct.reshape_mask(g["reshape_in"], "mid"), g["reshape_mid"])
    with pytest.raises(NotImplementedError):
        ct.reshape_mask(g["reshape_in"], "all")
    with pytest.raises(ValueError):
        ct.reshape_mask(g["reshape_in"], "first")
    assert (ct.N_FFT, ct.N_HOP, ct.WIN_LEN, ct.PRED_FRAME, ct.MASK_Z) == (512, 256, 21, "mid", "local")


def test_get_filter_type():
    from disco_b200.compat.internal_formulas import get_filter_type
    assert get_filter_type("gevd") == ("gevd", "Full")
    assert get_filter_type("r1-gevd") == ("gevd", 1)
    assert get_filter_type("mwf") == ("mwf", None)
