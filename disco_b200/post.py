"""The step right after the beamformer: back to the time domain and scoring, on
the device.  Mirrors the post-processing of disco_theque/speech_enhancement/tango.py:526-539 (six
``lb.core.istft`` calls per node) and disco_theque/metrics.py (``si_sdr`` :342-391, ``snr`` helpers).

`to_time` runs ONE batched iSTFT kernel launch for all outputs, nodes and utterances.  The metrics
are float64 reductions (torch on the device -- they are a few passes over short 1-D signals); the
third-octave filter-bank metrics (`fw_snr`, `fw_sd`) and the third-party `bss_eval` / STOI scores of
the reference stay outside this repository's scope.
"""
import torch

from . import ops


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
