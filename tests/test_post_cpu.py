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


def test_third_octave_filterbank_matches_reference_coefficients():
    """Own Butterworth band-pass design == the reference's sigproc_utils.third_octave_filterbank (scipy.signal.butter
    on acoustics' band edges), golden coefficients from the reference run (tests/golden/metrics_kat.npz)."""
    import os
    from disco_b200 import post
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "metrics_kat.npz"))
    F, I = post.third_octave_bands(16000)
    assert np.array_equal(F, g["F"]) and len(I) == 17
    assert np.array_equal(post.third_octave_bands(8000)[0], g["F8k"])
    b, a = post.third_octave_filterbank(F, 16000, order=4)
    for got, ref in ((b, g["bank_b4"]), (a, g["bank_a4"])):
        assert np.max(np.abs(got - ref) / np.max(np.abs(ref), axis=1, keepdims=True)) < 1e-13


def test_octave_band_restatement_known_values():
    """IEC 61260-1 base-10 third-octave bands: 1 kHz band = [10**2.95, 10**3.05] Hz, 160 Hz nominal -> 10**2.2."""
    from oracle.octave_np import OctaveBand
    ob = OctaveBand(center=1000, fraction=3)
    assert abs(ob.lower.item() - 10 ** 2.95) < 1e-9 and abs(ob.upper.item() - 10 ** 3.05) < 1e-9
    ob = OctaveBand(center=160, fraction=3)
    assert abs(ob.center.item() - 10 ** 2.2) < 1e-9
    ob = OctaveBand(center=[6300, 8000], fraction=3)
    assert np.allclose(ob.center, [10 ** 3.8, 10 ** 3.9])


def test_wav_codec_roundtrip(tmp_path):
    """disco_b200.wav_io: PCM16 -> float = x / 32768, float -> PCM16 = rint(x * 32767) (libsndfile's rules)."""
    from disco_b200 import wav_io
    x = np.array([0.0, 0.5, -0.5, 1.0, -1.0, 1.5 / 32767, 2.5 / 32767, 0.123456], dtype=np.float64)
    p = str(tmp_path / "a.wav")
    wav_io.write(p, x, 16000)
    y, fs = wav_io.read(p)
    assert fs == 16000 and y.dtype == np.float32
    want = np.array([0, 16384, -16384, 32767, -32767, 2, 2, 4045]) / 32768.0     # half to even: 1.5 -> 2, 2.5 -> 2
    assert np.array_equal(y, want.astype(np.float32))
    stereo = np.stack([x, -x], axis=1)
    wav_io.write(p, stereo, 8000)
    y2, fs2 = wav_io.read(p, dtype="float64")
    assert fs2 == 8000 and y2.shape == (8, 2) and np.array_equal(y2[:, 1], -y2[:, 0])


def test_post_generator_host_logic(tmp_path):
    """The host side of dataset_post.PostGenerator (no CUDA needed): data-set split, directory names, channel file
    ordering (numeric, Ch-10 after Ch-9) and the reference's assertions (post_generator.py:57-97)."""
    import pytest
    from disco_b200.dataset_post import PostGenerator
    from oracle.make_golden import make_post_dataset
    root = str(tmp_path)
    make_post_dataset(root, rirs=(1,), lengths=(2000,))
    gen = PostGenerator(1, 1, "living", "fs", [0, 6], root, n_samples=[10, 2, 2], device="cpu")
    assert gen.case == "train" and gen.snr_dir == "0-6" and gen.n_ch == 16
    tar, noi = gen.get_sig_lists(1)
    assert len(tar) == 16 and len(noi) == 1 and len(noi[0]) == 16
    assert [int(p.split("_Ch-")[-1].split(".wav")[0]) for p in tar] == list(range(1, 17))
    tars, nois = gen.load_sigs(tar, noi)
    assert tars.shape == (16, 2000) and nois.shape == (16, 2000) and tars.dtype == np.float32
    from disco_b200 import wav_io
    fs, long_noise = wav_io.read(noi[0][0], dtype="float32")[1], np.zeros(2001, dtype=np.float32)
    wav_io.write(noi[0][0], long_noise, fs)
    with pytest.raises(ValueError):                    # the reference's `noi_seg[:len(noi)] += noi` cannot broadcast
        gen.load_sigs(tar, noi)
    assert PostGenerator(11, 1, "living", "fs", [0, 6], root, n_samples=[10, 4, 2], device="cpu").case == "val"
    with pytest.raises(AssertionError):
        PostGenerator(9, 3, "living", "fs", [0, 6], root, n_samples=[10, 2, 2], device="cpu")      # spans two sets
    with pytest.raises(AssertionError):
        PostGenerator(0, 1, "living", "fs", [0, 6], root, n_samples=[10, 2, 2], device="cpu")
    with pytest.raises(ValueError):
        PostGenerator(1, 1, "living", "fs", [0, 6], root, n_fft=512, n_hop=128, n_samples=[10, 2, 2], device="cpu")
