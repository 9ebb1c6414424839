"""intern_filter and spatial_correlation_matrix with the reference signatures
(disco_theque/se_utils/internal_formulas.py:31-103)."""
import sys

import numpy as np
import torch

from .. import ops
from ._util import dev

eps = sys.float_info.epsilon      # internal_formulas.py:6
eta = 1e6                         # internal_formulas.py:7


def get_filter_type(filtre):
    """internal_formulas.py:10-28: 'gevd' / 'r1-gevd' ... -> ('gevd', rank) ; anything else -> (name, None)."""
    if "gevd" in filtre:
        rank = int(filtre.split("-")[0][-1]) if "-" in filtre else "Full"
        return "gevd", rank
    return filtre, None


def intern_filter(Rxx, Rnn, mu=1, type="r1-mwf", rank="Full"):
    """(Wint, (t1, sort_index)) for one pair of covariance matrices, computed by the batched CUDA
    solver (float64 arithmetic).  Wint, t1: complex128 vectors like the reference.
    Differences, documented: rank='Full'/'full' both mean all eigenpairs (the reference crashes on
    its own default 'Full', internal_formulas.py:31 vs :66-67); sort_index is None (the eigenvalue
    ordering is internal to the kernel)."""
    if type not in ops.FILTER_TYPES:
        raise AttributeError("Unknown filter reference")                       # internal_formulas.py:79
    Rs = dev(Rxx, torch.complex64)[None]
    Rn = dev(Rnn, torch.complex64)[None]
    W, t1 = ops.mwf_solve(Rs, Rn, float(mu), type, rank)
    W = W[0].cpu().numpy().astype(np.complex128)
    if type == "gevd":
        return W, (t1[0].cpu().numpy().astype(np.complex128), None)
    e1 = np.zeros(W.shape[0])
    e1[0] = 1.0                                                               # internal_formulas.py:43
    return W, (e1, None)


def spatial_correlation_matrix(Rxx, x, lambda_cor=0.95, M=None):
    """One step of the exponentially smoothed SCM, lambda R + (1 - lambda) [M] x x^H
    (internal_formulas.py:84-103).  Evaluated on the device in complex128."""
    R = dev(Rxx, torch.complex128)
    v = dev(x, torch.complex128)
    upd = (1 - lambda_cor) * torch.outer(v, v.conj())
    if M is not None:
        upd = M * upd
    return (lambda_cor * R + upd).cpu().numpy()
