"""The step right after the beamformer: back to the time domain and scoring, on
the device.  Mirrors the post-processing of disco_theque/speech_enhancement/tango.py:526-539 (six
``lb.core.istft`` calls per node) and disco_theque/metrics.py (``si_sdr`` :342-391, ``snr`` helpers).

`to_time` runs ONE batched iSTFT kernel launch for all outputs, nodes and utterances.  `fw_snr` / `fw_sd`
(metrics.py:63-128, 211-279) run the whole third-octave Butterworth bank over all signals in one launch of
the IIR filter-bank kernel (csrc/filterbank.cu) that returns only the band powers; the band weighting is a
few float64 operations on [..., 17] tensors.  `si_sdr`, `snr`, `sd` are float64 reductions (torch on the
device).  The third-party `bss_eval` / STOI scores of tango.main stay outside this repository's scope.

All metrics are batched: TIME IS THE LAST AXIS, every leading axis is a batch axis (the reference's
functions take one 1-D signal per call).
"""
import math

import numpy as np
import torch

from . import ops

# band importance function of the reference (metrics.py:81-95): ANSI S3.5 third-octave weights
_I_WIDE = np.array([83, 95, 150, 289, 440, 578, 653, 711, 818, 844, 882, 898, 868, 844, 771, 527, 364, 185]) * 1e-4
_F_WIDE = np.array([160, 200, 250, 315, 400, 500, 630, 800, 1000, 1250, 1600, 2000, 2500, 3150, 4000, 5000, 6300, 8000])
_I_NARROW = np.array([128, 320, 320, 447, 447, 639, 639, 767, 959, 1182, 1214, 1086, 1086, 757]) * 1e-4
_F_NARROW = np.array([200, 250, 315, 400, 500, 630, 800, 1000, 1250, 1600, 2000, 2500, 3150, 4000])
_G_OCTAVE = 10.0 ** 0.3          # IEC 61260-1:2014 octave ratio (what acoustics.signal.OctaveBand implements)


def to_time(outputs, length, n_fft=512, names=("yf", "z_y", "sf", "nf", "z_s", "z_n"), layout="FT"):
    """outputs: dict from tango_batched ([B, K, F, T] for layout 'FT', [B, K, T, F] for 'TF').
    Returns {name: [B, K, length] float32} for the names present (tango.py:526-539)."""
    present = [n for n in names if n in outputs]
    if not present:
        return {}
    specs = []
    for n in present:
        S = outputs[n]
        specs.append(ops.transpose_last2(S.contiguous()) if layout == "FT" else S)
    stack = torch.stack(specs).contiguous()                    # [n, B, K, T, F]
    x = ops.istft(stack, int(length), n_fft)                   # [n, B, K, length]
    return {n: x[i] for i, n in enumerate(present)}


def si_sdr(reference, estimation):
    """Scale-invariant SDR in dB over the last axis (metrics.py:342-391), float64 on the device."""
    ref = torch.as_tensor(reference).to(torch.float64)
    est = torch.as_tensor(estimation).to(device=ref.device, dtype=torch.float64)
    est, ref = torch.broadcast_tensors(est, ref)
    energy = (ref * ref).sum(-1, keepdim=True)
    proj = ((ref * est).sum(-1, keepdim=True) / energy) * ref
    noise = est - proj
    return 10.0 * torch.log10((proj * proj).sum(-1) / (noise * noise).sum(-1))


def snr_db(signal, noise):
    """10 log10 of the power ratio over the last axis (the plain SNR used around tango.py:552-593)."""
    s = torch.as_tensor(signal).to(torch.float64)
    n = torch.as_tensor(noise).to(device=s.device, dtype=torch.float64)
    return 10.0 * torch.log10((s * s).sum(-1) / (n * n).sum(-1))


# ------------------------------------------------------------------ frequency-weighted metrics
def third_octave_bands(fs):
    """Centre frequencies and importance weights the reference uses at sampling rate fs (metrics.py:80-95):
    the bands whose upper edge F * 2**(1/6) lies below fs / 2."""
    F, I = (_F_WIDE, _I_WIDE) if fs / 2 > 4500 else (_F_NARROW, _I_NARROW)
    n = int(np.sum(F * 2 ** (1 / 6) < fs / 2))
    return F[:n], I[:n]


def _butter_bandpass(order, lo, hi):
    """Digital Butterworth band-pass of prototype order `order` (filter order 2 * order), edges normalised to
    Nyquist = 1: analog prototype -> pre-warped band-pass transform -> bilinear transform -> polynomials
    (the textbook route, the one scipy.signal.butter(order, [lo, hi], 'bandpass') follows)."""
    fs = 2.0
    w = 2 * fs * np.tan(np.pi * np.array([lo, hi], dtype=np.float64) / fs)
    bw, wo = w[1] - w[0], math.sqrt(w[0] * w[1])
    m = np.arange(-order + 1, order, 2)
    p_lp = -np.exp(1j * np.pi * m / (2 * order)) * bw / 2                  # Butterworth poles, scaled to bw
    root = np.sqrt(p_lp ** 2 - wo ** 2)
    p_bp = np.concatenate((p_lp + root, p_lp - root))
    z_bp = np.zeros(order)
    k_bp = bw ** order
    fs2 = 2 * fs
    z_d = np.append((fs2 + z_bp) / (fs2 - z_bp), -np.ones(len(p_bp) - len(z_bp)))
    p_d = (fs2 + p_bp) / (fs2 - p_bp)
    k_d = k_bp * np.real(np.prod(fs2 - z_bp) / np.prod(fs2 - p_bp))
    return np.real(k_d * np.poly(z_d)), np.real(np.poly(p_d))


def third_octave_filterbank(F, fs, order=8):
    """Reference signature (sigproc_utils.py:90-116): row i = coefficients (b, a) of the order-`order`
    Butterworth band-pass over the third-octave band around F[i].  Band edges: IEC 61260-1 base-10 exact
    mid-band frequency of the band nearest to F[i], times G**(-1/6), G**(+1/6), G = 10**0.3."""
    F = np.atleast_1d(np.asarray(F, dtype=np.float64))
    b = np.zeros((len(F), 2 * order + 1))
    a = np.zeros((len(F), 2 * order + 1))
    for i, f in enumerate(F):
        idx = np.round(3 * np.log(f / 1000.0) / np.log(_G_OCTAVE))
        centre = 1000.0 * _G_OCTAVE ** (idx / 3)
        lo, hi = centre * _G_OCTAVE ** (-1 / 6), centre * _G_OCTAVE ** (1 / 6)
        b[i], a[i] = _butter_bandpass(order, lo * 2 / fs, hi * 2 / fs)
    return b, a


_bank_cache = {}


def _bank(fs, order, device):
    key = (int(fs), int(order), str(device))
    if key not in _bank_cache:
        F, I = third_octave_bands(fs)
        b, a = third_octave_filterbank(F, fs, order=order)
        ba = torch.from_numpy(np.stack([b, a], axis=1)).to(device)          # [N, 2, 2*order+1]
        _bank_cache[key] = (F, torch.from_numpy(I / np.sum(I)).to(device), ba)
    return _bank_cache[key]


def band_powers_db(x, fs, order=4, vad=None):
    """10 log10 of the variance of every third-octave band of x ([..., L] float32 on the device) over the
    non-zero filter outputs, or over the samples with vad != 0 (metrics.py:96-107) -> [..., N] float64."""
    _, _, ba = _bank(fs, order, x.device)
    st = ops.band_stats(x, ba, sel=None if vad is None else vad.to(torch.float32))
    cnt, sm, sq = st[..., 0], st[..., 1], st[..., 2]
    mean = sm / cnt
    return 10.0 * torch.log10(sq / cnt - mean * mean)


def fw_snr(s, n, fs, vad_tar=None, vad_noi=None, clipping=1, db=True):
    """Frequency-weighted SNR (reference metrics.py:63-128), batched over the leading axes.
    Returns (fqwt_snr [..., N], fw_snr_mean [...], F)."""
    F, w, _ = _bank(fs, 4, s.device)
    snr_var = band_powers_db(s, fs, 4, vad_tar) - band_powers_db(n, fs, 4, vad_noi)
    if clipping:
        snr_var = snr_var.clamp(-15.0, 25.0)
    fq = w * snr_var
    mean = fq.sum(-1)
    if not db:
        fq, mean = 10.0 ** (fq / 10.0), 10.0 ** (mean / 10.0)
    return fq, mean, F


def fw_sd(s_out, s_in, fs, clipping=1, db=True):
    """Frequency-weighted speech distortion (reference metrics.py:211-279).  Returns (fqwt_sd, fw_sd_mean, F)."""
    F, w, _ = _bank(fs, 4, s_out.device)
    sd_var = band_powers_db(s_in, fs, 4) - band_powers_db(s_out, fs, 4)
    if clipping:
        sd_var = sd_var.clamp(0.0, 25.0)
    fq = w * sd_var
    mean = fq.sum(-1)
    if not db:
        fq, mean = 10.0 ** (fq / 10.0), 10.0 ** (mean / 10.0)
    return fq, mean, F


def _var_nonzero(x):
    x = x.to(torch.float64)
    nz = (x != 0)
    cnt = nz.sum(-1)
    mean = x.sum(-1) / cnt
    return ((x - mean.unsqueeze(-1)) ** 2 * nz).sum(-1) / cnt


def snr(s, n, db=True):
    """metrics.py:9-23: ratio of the variances of the non-zero samples."""
    r = _var_nonzero(s) / _var_nonzero(n)
    return 10.0 * torch.log10(r) if db else r


def delta_snr(s_out, n_out, s_in, n_in, db=True):
    """metrics.py:26-45."""
    d = snr(s_out, n_out, True) - snr(s_in, n_in, True)
    return d if db else 10.0 ** (d / 10.0)


def sd(s_out, s_in, db=True):
    """metrics.py:48-62."""
    r = _var_nonzero(s_in) / _var_nonzero(s_out)
    return 10.0 * torch.log10(r) if db else r
