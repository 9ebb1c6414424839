"""The online (recursive) oracle: per-frame composition of the restated spatial_correlation_matrix / intern_filter
(oracle/tango_np.py) reproduces the same composition of the REFERENCE's functions (tests/golden/online_kat.npz)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden", "online_kat.npz")


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def test_online_oracle_reproduces_reference_composition():
    from oracle import online_np, tango_np
    from oracle.make_golden import online_inputs
    g = np.load(GOLD)
    X, mask = online_inputs()
    for tag, kw in (("p8l1", dict(block=8, lag=1)), ("p5l0", dict(block=5, lag=0)),
                    ("p8l1pow1", dict(block=8, lag=1, power=1, lambda_cor=0.9))):
        z, W, Rs, Rn = online_np.online_mwf(X, mask, tango_np.spatial_correlation_matrix, tango_np.intern_filter, **kw)
        assert rel(z, g[tag + "_z"]) < 2e-6 and rel(W, g[tag + "_W"]) < 2e-6
        assert rel(Rs[-2:], g[tag + "_Rss"]) < 1e-6 and rel(Rn[-2:], g[tag + "_Rnn"]) < 1e-6
    # the recursion really is first order: the last snapshot of power 2 equals the closed form
    lam, T = 0.95, X.shape[2]
    wts = (1 - lam) * lam ** (T - 1 - np.arange(T))
    f = 17
    closed = np.einsum("t,t,it,jt->ij", wts, mask[f].astype(np.float64) ** 2, X[:, f], X[:, f].conj())
    z, W, Rs, Rn = online_np.online_mwf(X[:, f:f + 1], mask[f:f + 1], tango_np.spatial_correlation_matrix,
                                        tango_np.intern_filter, block=8, lag=1)
    assert rel(Rs[-1, 0], closed) < 1e-12
