"""tf_mask and vad_oracle_batch with the reference signatures
(disco_theque/sigproc_utils.py:12-86; tf_mask twin in disco_theque/dnn/utils.py:44-71)."""
import math

import numpy as np
import torch

from .. import ops
from ._util import DEVICE, dev


def tf_mask(s, n, type="irm1", bin_thr=0):
    """TF mask from target and noise STFTs ('irmX' Wiener-like, 'ibmX' binary, 'iamX' amplitude).
    Same shape as the inputs; float32 ('ibm' -> bool, like the reference's comparison result)."""
    if np.shape(s) != np.shape(n):
        raise AssertionError("Input spectrograms should have the same shape.")   # sigproc_utils.py:71
    m = ops.tf_mask(dev(s, torch.complex64), dev(n, torch.complex64), type, bin_thr).cpu().numpy()
    return m.astype(bool) if "ibm" in type else m


def vad_oracle_rows_device(x, win_len=512, win_hop=256, thr=0.001, rat=2):
    """Energy VAD of sigproc_utils.py:12-55 for every ROW of x [R, n] at once on the device (float64): a window is
    speech when at least int(N_win / rat) of its samples exceed thr * q99(x^2); returns 0/1 per sample [R, n]."""
    x = x.to(device=DEVICE, dtype=torch.float64)
    x = x - x.mean(dim=1, keepdim=True)
    x2 = (x * x).abs()
    R, n = x2.shape
    thr_ = thr * torch.quantile(x2, 0.99, dim=1, keepdim=True)
    n_win = int(math.ceil((n - win_len) / win_hop + 1))
    above = torch.cat([torch.zeros((R, 1), dtype=torch.int64, device=x.device), (x2 > thr_).to(torch.int64).cumsum(1)], dim=1)
    starts = torch.arange(n_win, device=x.device) * win_hop
    ends = torch.clamp(starts + win_len, max=n)
    cnt = above[:, ends] - above[:, starts]
    need = ((ends - starts).to(torch.float64) / rat).to(torch.int64)          # np.int(N_ / rat) truncates
    active = (cnt >= need[None]).to(torch.int64)                               # [R, n_win]
    diff = torch.zeros((R, n + 1), dtype=torch.int64, device=x.device)         # union of the active windows
    diff.scatter_add_(1, starts[None].expand(R, -1), active)
    diff.scatter_add_(1, ends[None].expand(R, -1), -active)
    return (diff[:, :-1].cumsum(1) > 0).to(torch.float64)


def vad_oracle_batch_device(x_, win_len=512, win_hop=256, thr=0.001, rat=2):
    """One signal: 0/1 per SAMPLE (float64 tensor)."""
    x = x_ if isinstance(x_, torch.Tensor) else torch.from_numpy(np.asarray(x_, dtype=np.float64))
    return vad_oracle_rows_device(x.reshape(1, -1), win_len, win_hop, thr, rat)[0]


def vad_oracle_batch(x_, win_len=512, win_hop=256, thr=0.001, rat=2):
    """Reference signature: np.ndarray of 0/1 per sample (float64)."""
    return vad_oracle_batch_device(x_, win_len, win_hop, thr, rat).cpu().numpy()
