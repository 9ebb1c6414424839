"""Host-side helpers of the compat layer against the reference's own outputs (tests/golden)."""
import os
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


# ---- drop-in boundary: names, positional order and defaults of the reference's functions
OURS = {
    "disco_theque/speech_enhancement/tango.py:offline_tango": "disco_b200.tango:offline_tango",
    "disco_theque/speech_enhancement/tango.py:get_mask": "disco_b200.compat.tango:get_mask",
    "disco_theque/speech_enhancement/tango.py:concatenate_signals": "disco_b200.compat.tango:concatenate_signals",
    "disco_theque/speech_enhancement/tango.py:get_z_for_mask": "disco_b200.compat.tango:get_z_for_mask",
    "disco_theque/speech_enhancement/tango.py:reshape_mask": "disco_b200.compat.tango:reshape_mask",
    "disco_theque/speech_enhancement/get_z_signals.py:offline_tango": "disco_b200.compat.get_z_signals:offline_tango",
    "disco_theque/se_utils/internal_formulas.py:get_filter_type": "disco_b200.compat.internal_formulas:get_filter_type",
    "disco_theque/se_utils/internal_formulas.py:intern_filter": "disco_b200.compat.internal_formulas:intern_filter",
    "disco_theque/se_utils/internal_formulas.py:spatial_correlation_matrix":
        "disco_b200.compat.internal_formulas:spatial_correlation_matrix",
    "disco_theque/dnn/utils.py:tf_mask": "disco_b200.compat.sigproc_utils:tf_mask",
    "disco_theque/sigproc_utils.py:tf_mask": "disco_b200.compat.sigproc_utils:tf_mask",
    "disco_theque/sigproc_utils.py:vad_oracle_batch": "disco_b200.compat.sigproc_utils:vad_oracle_batch",
    "disco_theque/math_utils.py:my_stft": "disco_b200.compat.math_utils:my_stft",
    "disco_theque/math_utils.py:my_istft": "disco_b200.compat.math_utils:my_istft",
    "disco_theque/metrics.py:snr": "disco_b200.compat.metrics:snr",
    "disco_theque/metrics.py:delta_snr": "disco_b200.compat.metrics:delta_snr",
    "disco_theque/metrics.py:sd": "disco_b200.compat.metrics:sd",
    "disco_theque/metrics.py:fw_snr": "disco_b200.compat.metrics:fw_snr",
    "disco_theque/metrics.py:fw_sd": "disco_b200.compat.metrics:fw_sd",
    "disco_theque/metrics.py:si_sdr": "disco_b200.compat.metrics:si_sdr",
    "disco_theque/speech_enhancement/utils.py:prepare_data": "disco_b200.dnn_mask:prepare_data",
}


def test_adapter_signatures_match_the_reference(golden_dir):
    """Every reference parameter exists here at the same position, under the same name, with the same default
    (tests/golden/reference_signatures.json, extracted from the reference source by oracle/make_signatures.py);
    adapters may only ADD parameters behind them (keyword-only or defaulted: device, n_fft, ...)."""
    import importlib
    import inspect
    import json
    ref = json.load(open(os.path.join(golden_dir, "reference_signatures.json")))
    assert set(ref) == set(OURS)
    for key, target in OURS.items():
        mod, name = target.split(":")
        fn = getattr(importlib.import_module(mod), name)
        ours = list(inspect.signature(fn).parameters.values())
        for i, want in enumerate(ref[key]["params"]):
            assert i < len(ours), (key, "missing parameter", want["name"])
            got = ours[i]
            assert got.name == want["name"], (key, i, got.name, want["name"])
            assert got.kind in (got.POSITIONAL_OR_KEYWORD, got.POSITIONAL_ONLY), (key, got.name)
            if want["has_default"]:
                assert got.default is not inspect.Parameter.empty and got.default == want["default"], (key, got.name, got.default)
            else:
                assert got.default is inspect.Parameter.empty, (key, got.name)
        for extra in ours[len(ref[key]["params"]):]:            # additions must not break positional callers
            assert (extra.kind in (extra.KEYWORD_ONLY, extra.VAR_KEYWORD, extra.VAR_POSITIONAL)
                    or extra.default is not inspect.Parameter.empty), (key, extra.name)
        if ref[key]["varkw"]:
            assert any(p.kind == p.VAR_KEYWORD for p in ours), key
