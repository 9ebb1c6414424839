"""Time-domain metrics with the reference signatures (disco_theque/metrics.py): NumPy 1-D signals in, NumPy / float
out, computed on the device by disco_b200.post (IIR filter-bank kernel + float64 reductions).

    snr, delta_snr, sd        metrics.py:9-62
    fw_snr                    metrics.py:63-128
    fw_sd                     metrics.py:211-279
    si_sdr                    metrics.py:342-391   (also accepts stacked signals, like the reference)
"""
import numpy as np
import torch

from .. import post
from ._util import dev


def _sig(x, name):
    x = np.asarray(x)
    if x.ndim != 1:
        raise NotImplementedError("%s: one 1-D signal per call here; disco_b200.post takes batches (time = last axis)" % name)
    return dev(x.astype(np.float32), torch.float32)


def _scalar(t):
    return float(t.item())


def snr(s, n, db=True):
    return _scalar(post.snr(_sig(s, "s"), _sig(n, "n"), db))


def delta_snr(s_out, n_out, s_in, n_in, db=True):
    return _scalar(post.delta_snr(_sig(s_out, "s_out"), _sig(n_out, "n_out"), _sig(s_in, "s_in"), _sig(n_in, "n_in"), db))


def sd(s_out, s_in, db=True):
    return _scalar(post.sd(_sig(s_out, "s_out"), _sig(s_in, "s_in"), db))


def fw_snr(s, n, fs, vad_tar=None, vad_noi=None, clipping=1, db=True):
    """-> (fqwt_snr (N,), fw_snr_mean, F): per-band weighted SNRs, their sum, the band centre frequencies."""
    vt = None if vad_tar is None else _sig(vad_tar, "vad_tar")
    vn = None if vad_noi is None else _sig(vad_noi, "vad_noi")
    fq, mean, F = post.fw_snr(_sig(s, "s"), _sig(n, "n"), fs, vt, vn, clipping, db)
    return fq.cpu().numpy(), _scalar(mean), F


def fw_sd(s_out, s_in, fs, clipping=1, db=True):
    fq, mean, F = post.fw_sd(_sig(s_out, "s_out"), _sig(s_in, "s_in"), fs, clipping, db)
    return fq.cpu().numpy(), _scalar(mean), F


def si_sdr(reference, estimation):
    """Scale-invariant SDR in dB over the last axis; scalars for 1-D input, arrays for stacked signals."""
    ref = dev(np.asarray(reference, dtype=np.float64), torch.float64)
    est = dev(np.asarray(estimation, dtype=np.float64), torch.float64)
    out = post.si_sdr(ref, est).cpu().numpy()
    return float(out) if out.ndim == 0 else out
