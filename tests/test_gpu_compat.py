"""Reference-signature adapters (disco_b200/compat) on the GPU against the reference's outputs."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def test_tf_mask_and_vad_signatures():
    from disco_b200.compat.sigproc_utils import tf_mask, vad_oracle_batch
    g = load_golden("helpers_kat")
    for typ in ("irm1", "irm2", "ibm1", "ibm2"):
        got = tf_mask(g["s"], g["n"], type=typ)
        ref = g["sig_" + typ]
        assert got.shape == ref.shape
        if typ.startswith("ibm"):
            assert got.dtype == bool and np.array_equal(got, ref)
        else:
            assert np.max(np.abs(got - ref)) < 1e-6
    with pytest.raises(AssertionError):
        tf_mask(g["s"], g["n"][:, :5])
    with pytest.raises(ValueError):
        tf_mask(g["s"], g["n"], type="xyz1")
    assert np.array_equal(vad_oracle_batch(g["vad_x"]), g["vad_default"])
    assert np.array_equal(vad_oracle_batch(g["vad_x"], 256, 128, 0.01, 4), g["vad_256_128"])


def test_intern_filter_signature_and_kats():
    from disco_b200.compat.internal_formulas import intern_filter, spatial_correlation_matrix
    g = load_golden("intern_filter_kat")
    for i in range(0, int(g["count"]), 3):
        typ, rank, mu = str(g["cfg_%d" % i]).split("|")
        kw = {} if rank == "None" else {"rank": rank if rank == "full" else int(rank)}
        if typ == "gevd" and not kw:
            kw = {"rank": "Full"}
        W, (t1, sort_index) = intern_filter(g["Rxx_%d" % i], g["Rnn_%d" % i], mu=float(mu), type=typ, **kw)
        assert W.dtype == np.complex128 and sort_index is None
        tol = 5e-4 if g["Rxx_%d" % i].dtype == np.complex64 else 1e-5
        assert rel_l2(W, g["W_%d" % i]) < tol and rel_l2(t1, g["t1_%d" % i]) < tol
    with pytest.raises(AttributeError):
        intern_filter(g["Rxx_0"], g["Rnn_0"], type="foo")
    h = load_golden("helpers_kat")
    assert np.allclose(spatial_correlation_matrix(h["scm_R0"], h["scm_x"]), h["scm_plain"], atol=1e-14)
    assert np.allclose(spatial_correlation_matrix(h["scm_R0"], h["scm_x"], 0.9, 0.3), h["scm_masked"], atol=1e-14)


def test_my_stft_my_istft():
    from disco_b200.compat.math_utils import my_istft, my_stft
    from oracle import librosa_np
    x = np.random.default_rng(0).standard_normal(12345).astype(np.float32)
    Y = my_stft(x)
    assert Y.shape == (257, 1 + 12345 // 256) and Y.dtype == np.complex64
    assert rel_l2(Y, librosa_np.stft(x)) < 2e-6
    back = my_istft(Y, 12345)
    assert back.shape == (12345,) and np.max(np.abs(back - x)) < 5e-6


def test_get_mask_ivad_matches_reference_golden():
    from disco_b200.compat.tango import get_mask
    from oracle.make_golden import TANGO_CASES, case_inputs
    seed, chans, length, vads, mfz, _ = TANGO_CASES["tango_k2c2_ivad"]
    g = load_golden("tango_k2c2_ivad")
    y, s, n = case_inputs(seed, chans, length, vads)
    for k in range(2):
        m = get_mask(None, np.zeros((257, 33)), None, mask_type="ivad", ts=s[k][0])
        assert np.array_equal(m, g["masks_z_%d" % k])
    with pytest.raises(ValueError):
        get_mask(None, np.zeros((3, 3)), None, mask_type="nope")


def test_get_z_signals_step1_twin():
    """get_z_signals.offline_tango returns the step-1 outputs of the full Tango (same z_y, z_s, z_n, zn, masks)."""
    from disco_b200.compat.get_z_signals import offline_tango as step1_only
    from oracle.make_golden import NAMES, TANGO_CASES, case_inputs
    seed, chans, length, vads, mfz, _ = TANGO_CASES["tango_k2c3_local"]
    g = load_golden("tango_k2c3_local")
    y, s, n = case_inputs(seed, chans, length, vads)
    z_y, z_s, z_n, zn, mz = step1_only(y, s, n, "irm1", [None], "local")
    for k in range(2):
        for got, nm in ((z_y, "z_y"), (z_s, "z_s"), (z_n, "z_n"), (zn, "zn")):
            ref = g["%s_%d" % (nm, k)]
            assert np.linalg.norm(np.abs(got[k]) - np.abs(ref)) / np.linalg.norm(np.abs(ref)) < 1e-5
        assert np.max(np.abs(mz[k] - g["masks_z_%d" % k])) < 5e-6


def test_metrics_reference_signatures():
    """compat.metrics: the reference's call signatures (one 1-D NumPy signal per call) against its own outputs."""
    from disco_b200.compat import metrics as m
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "metrics_kat.npz"))
    fs = int(g["fs"])
    s, n, est, vad = g["s"], g["n"], g["est"], g["vad"]
    for i in range(3):
        fq, mean, F = m.fw_snr(s[i], n[i], fs)
        assert isinstance(mean, float) and fq.shape == (17,) and np.array_equal(F, g["F"])
        assert abs(mean - g["fw_snr_mean_%d" % i]) < 1e-6 and np.max(np.abs(fq - g["fw_snr_fq_%d" % i])) < 1e-6
        _, mean_v, _ = m.fw_snr(s[i], n[i], fs, vad_tar=vad[i], vad_noi=vad[i])
        assert abs(mean_v - g["fw_snr_vad_mean_%d" % i]) < 1e-6
        _, mean_d, _ = m.fw_sd(est[i], s[i], fs)
        assert abs(mean_d - g["fw_sd_mean_%d" % i]) < 1e-6
        assert abs(m.snr(s[i], n[i]) - g["snr_%d" % i]) < 1e-4 and abs(m.sd(est[i], s[i]) - g["sd_%d" % i]) < 1e-4
        assert abs(m.delta_snr(0.8 * s[i], 0.3 * n[i], s[i], n[i]) - g["delta_snr_%d" % i]) < 1e-4
        assert abs(m.si_sdr(s[i].astype(np.float64), est[i].astype(np.float64)) - g["si_sdr_%d" % i]) < 1e-6
    both = m.si_sdr(s[:2].astype(np.float64), est[:2].astype(np.float64))
    assert both.shape == (2,) and abs(both[1] - g["si_sdr_1"]) < 1e-6
    with pytest.raises(NotImplementedError):
        m.fw_snr(np.zeros((10, 2)), np.zeros((10, 2)), fs)
