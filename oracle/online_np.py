"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Frame-by-frame restatement of the recursive MWF step that disco_b200/online.py evaluates in batched form, written
as the composition of the reference's own two functions, called exactly as a Python user would call them:
    spatial_correlation_matrix(Rxx, x, lambda_cor, M)   se_utils/internal_formulas.py:84-103   (every frame)
    intern_filter(Rxx, Rnn, mu, 'gevd', rank)           se_utils/internal_formulas.py:31-81    (every block)
`scm` and `solve` are injected: the unmodified reference functions (oracle/ref_shim.py) when the golden vectors
are generated, the NumPy restatements of oracle/tango_np.py in the tests.
"""
import numpy as np


def online_mwf(X, mask, scm, solve, lambda_cor=0.95, block=8, lag=1, mu=1.0, rank=1, ref=0, power=2, R0=None):
    """X (D, F, T) complex spectra of the concatenated channels, mask (F, T).  Returns z (F, T), W (J, F, D),
    Rss, Rnn (J, F, D, D) -- the smoothed matrices after the last frame of every block."""
    D, F, T = X.shape
    J = (T + block - 1) // block
    Rss = np.zeros((F, D, D), dtype=np.complex128) if R0 is None else np.array(R0[0], dtype=np.complex128)
    Rnn = np.zeros((F, D, D), dtype=np.complex128) if R0 is None else np.array(R0[1], dtype=np.complex128)
    snap_s = np.zeros((J, F, D, D), dtype=np.complex128)
    snap_n = np.zeros_like(snap_s)
    W = np.zeros((J, F, D), dtype=np.complex128)
    z = np.zeros((F, T), dtype=np.complex128)
    for t in range(T):
        j = t // block
        for f in range(F):
            m = float(mask[f, t])
            x = X[:, f, t].astype(np.complex128)
            if power == 2:       # x = estimate of the component, M = None
                Rss[f] = scm(Rss[f], m * x, lambda_cor)
                Rnn[f] = scm(Rnn[f], (1.0 - m) * x, lambda_cor)
            else:                # x = mixture, M = mask
                Rss[f] = scm(Rss[f], x, lambda_cor, m)
                Rnn[f] = scm(Rnn[f], x, lambda_cor, 1.0 - m)
            jw = j - lag
            w = W[jw, f] if jw >= 0 else np.eye(D)[ref]
            z[f, t] = np.inner(np.conj(w), x)
        if t == min(T, (j + 1) * block) - 1:      # block complete: refresh the filter
            snap_s[j], snap_n[j] = Rss, Rnn
            for f in range(F):
                W[j, f] = solve(Rss[f], Rnn[f], mu, "gevd", rank)[0]
            if lag == 0:                          # look-ahead variant: re-filter the block with its own filter
                for tt in range(j * block, t + 1):
                    z[:, tt] = np.einsum("fd,df->f", np.conj(W[j]), X[:, :, tt])
    return z, W, snap_s, snap_n
