"""Tensor-level operators of the MWF path: thin wrappers that hand torch CUDA tensors (device
memory + current stream) to the C ABI of libdisco_b200.so.  PyTorch is plumbing here: all
arithmetic happens in the hand-written kernels.

Layouts: spectra are frame-major ``[..., T, F]`` complex64; ``layout='FT'`` arguments select the
reference's NumPy layout ``[..., F, T]`` for masks / final outputs (SURVEY.md §8b op table).
"""
import ctypes

import torch

from . import _lib

TF, FT = 0, 1
MASK_KINDS = {"irm": 0, "ibm": 1, "iam": 2}
FILTER_TYPES = {"gevd": 0, "r1-mwf": 1, "mwf": 2}


def _layout(layout):
    if layout in (TF, "TF", "tf"):
        return TF
    if layout in (FT, "FT", "ft"):
        return FT
    raise ValueError("layout must be 'TF' or 'FT'")


def _stream():
    """Current stream of the CURRENT device; every public op below runs under the device of its operands
    (see `_on_device`), so this is the stream of the tensors' device."""
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_device(fn):
    """Run an op with the device of its tensor operands current: the library keeps per-device tables
    (cudaGetDevice) and launches on the current stream, so operands on another device than the current one
    would otherwise be dereferenced by kernels running on the wrong GPU.  Mixed devices are rejected."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kw):
        dev = None
        stack = list(args) + list(kw.values())
        while stack:
            a = stack.pop()
            if isinstance(a, (tuple, list)):
                stack.extend(a)
            elif isinstance(a, torch.Tensor) and a.is_cuda:
                if dev is None:
                    dev = a.device
                elif a.device != dev:
                    raise ValueError("%s: operands live on different devices (%s, %s)" % (fn.__name__, dev, a.device))
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kw)
        with torch.cuda.device(dev):
            return fn(*args, **kw)
    return wrapper


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _need(t, dtype, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA tensor (disco_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def n_frames(length, n_fft=512):
    """1 + L // hop (reference tango.py:287)."""
    return 1 + length // (n_fft // 2)


def init(n_fft=512):
    """Create the per-device FFT tables now (required before CUDA-graph capture)."""
    _lib.check(_lib.load().disco_init(int(n_fft)))


@_on_device
def stft(x, n_fft=512):
    """x [..., L] float32 -> Y [..., T, F] complex64 (librosa center/reflect/periodic-Hann semantics)."""
    _need(x, torch.float32, "x")
    L = x.shape[-1]
    n_sig = x.numel() // L
    T, F = n_frames(L, n_fft), n_fft // 2 + 1
    Y = torch.empty(x.shape[:-1] + (T, F), dtype=torch.complex64, device=x.device)
    _lib.check(_lib.load().disco_stft(_ptr(x), _ptr(Y), n_sig, L, n_fft, _stream()))
    return Y


@_on_device
def stft_scm(x, mask, n_fft=512, mask_layout="TF", keep_partials=False):
    """Fused STFT + masked SCM.  x [G, C, L] float32, mask [G, T, F] (or [G, F, T]) float32
    -> Y [G, C, T, F] complex64, Rss, Rnn [G, F, C, C] complex64.
    keep_partials=True returns (Y, workspace) instead: the SCMs stay as per-segment partial sums that
    mwf_solve_workspace() consumes directly (one launch fewer)."""
    _need(x, torch.float32, "x")
    _need(mask, torch.float32, "mask")
    if x.dim() != 3:
        raise ValueError("x must be [groups, channels, samples]")
    G, C, L = x.shape
    T, F = n_frames(L, n_fft), n_fft // 2 + 1
    lib = _lib.load()
    if not lib.disco_stft_scm_supported(n_fft, C, 1):
        raise NotImplementedError("fused STFT+SCM: %d channels at n_fft=%d (use stft + masked_scm)" % (C, n_fft))
    lay = _layout(mask_layout)
    want = (G, T, F) if lay == TF else (G, F, T)
    if tuple(mask.shape) != want:
        raise ValueError("mask shape %s, expected %s" % (tuple(mask.shape), want))
    Y = torch.empty((G, C, T, F), dtype=torch.complex64, device=x.device)
    ws_bytes = lib.disco_stft_scm_workspace(G, C, L, n_fft)
    ws = torch.empty(max(ws_bytes, 16) // 4, dtype=torch.float32, device=x.device)
    if keep_partials:
        _lib.check(lib.disco_stft_scm(_ptr(x), _ptr(mask), lay, _ptr(Y), None, None, G, C, L, n_fft,
                                      _ptr(ws), ws_bytes, _stream()))
        return Y, ws
    Rss = torch.empty((G, F, C, C), dtype=torch.complex64, device=x.device)
    Rnn = torch.empty_like(Rss)
    _lib.check(lib.disco_stft_scm(_ptr(x), _ptr(mask), lay, _ptr(Y), _ptr(Rss), _ptr(Rnn), G, C, L, n_fft,
                                  _ptr(ws), ws_bytes, _stream()))
    return Y, Rss, Rnn


@_on_device
def mwf_solve_workspace(ws, G, C, L, n_fft=512, mu=1.0, type="gevd", rank=1, want_scm=False):
    """mwf_solve on the SCMs a preceding stft_scm(..., keep_partials=True) left in `ws`.
    Returns W, t1 [G, F, C] (and Rss, Rnn [G, F, C, C] when want_scm)."""
    if type not in FILTER_TYPES:
        raise AttributeError("Unknown filter reference")
    F = n_fft // 2 + 1
    r = 0 if rank in ("full", "Full", None) else int(rank)
    W = torch.empty((G, F, C), dtype=torch.complex64, device=ws.device)
    T1 = torch.empty_like(W)
    Rss = Rnn = None
    if want_scm:
        Rss = torch.empty((G, F, C, C), dtype=torch.complex64, device=ws.device)
        Rnn = torch.empty_like(Rss)
    _lib.check(_lib.load().disco_mwf_solve_workspace(_ptr(ws), _ptr(W), _ptr(T1), _ptr(Rss), _ptr(Rnn), G, C, L, n_fft,
                                                     FILTER_TYPES[type], r, float(mu), _stream()))
    return (W, T1, Rss, Rnn) if want_scm else (W, T1)


def set_reserved_sms(n):
    """Leave n SMs free of the persistent fused STFT+SCM kernel (room for a concurrent NCCL collective)."""
    _lib.check(_lib.load().disco_set_reserved_sms(int(n)))


def stft_scm_supported(n_fft, C, n_mask=1):
    """Whether the fused STFT+SCM kernel covers (n_fft, channels per group, number of masks)."""
    return bool(_lib.load().disco_stft_scm_supported(int(n_fft), int(C), int(n_mask)))


@_on_device
def stft_scm2(x, mask_a, mask_b, n_fft=512, mask_layout="TF"):
    """Fused STFT + the masked SCMs under TWO masks in one pass (single-node arrays: step-1 and step-2
    statistics are taken over the same Y, reference tango.py:357-364 and :431-440 with K = 1).
    x [G, C, L], masks [G, T, F] (or [G, F, T]) -> Y [G, C, T, F], workspace (partial sums of both sets,
    consumed by mwf_solve_workspace2 / scm_from_workspace)."""
    _need(x, torch.float32, "x")
    _need(mask_a, torch.float32, "mask_a")
    _need(mask_b, torch.float32, "mask_b")
    if x.dim() != 3:
        raise ValueError("x must be [groups, channels, samples]")
    G, C, L = x.shape
    T, F = n_frames(L, n_fft), n_fft // 2 + 1
    lib = _lib.load()
    if not lib.disco_stft_scm_supported(n_fft, C, 2):
        raise NotImplementedError("two-mask fused STFT+SCM: %d channels at n_fft=%d" % (C, n_fft))
    lay = _layout(mask_layout)
    want = (G, T, F) if lay == TF else (G, F, T)
    if tuple(mask_a.shape) != want or tuple(mask_b.shape) != want:
        raise ValueError("mask shapes %s / %s, expected %s" % (tuple(mask_a.shape), tuple(mask_b.shape), want))
    Y = torch.empty((G, C, T, F), dtype=torch.complex64, device=x.device)
    ws_bytes = lib.disco_stft_scm2_workspace(G, C, L, n_fft)
    ws = torch.empty(max(ws_bytes, 16) // 4, dtype=torch.float32, device=x.device)
    _lib.check(lib.disco_stft_scm2(_ptr(x), _ptr(mask_a), _ptr(mask_b), lay, _ptr(Y), G, C, L, n_fft, _ptr(ws),
                                   ws_bytes, _stream()))
    return Y, ws


@_on_device
def scm_from_workspace(ws, G, C, L, n_fft=512, n_set=1, set=0):
    """Rss, Rnn [G, F, C, C] of mask set `set` from the partial sums a fused STFT+SCM call left in `ws`."""
    F = n_fft // 2 + 1
    Rss = torch.empty((G, F, C, C), dtype=torch.complex64, device=ws.device)
    Rnn = torch.empty_like(Rss)
    _lib.check(_lib.load().disco_scm_from_workspace(_ptr(ws), int(n_set), int(set), _ptr(Rss), _ptr(Rnn), G, C, L,
                                                    n_fft, _stream()))
    return Rss, Rnn


@_on_device
def mwf_solve_workspace2(ws, G, C, L, n_fft=512, mu=1.0, type="gevd", rank=1):
    """Both filter sets of a stft_scm2 workspace in one launch: W, t1 [2, G, F, C] (0: mask_a, 1: mask_b)."""
    if type not in FILTER_TYPES:
        raise AttributeError("Unknown filter reference")
    F = n_fft // 2 + 1
    r = 0 if rank in ("full", "Full", None) else int(rank)
    W = torch.empty((2, G, F, C), dtype=torch.complex64, device=ws.device)
    T1 = torch.empty_like(W)
    _lib.check(_lib.load().disco_mwf_solve_workspace2(_ptr(ws), _ptr(W), _ptr(T1), G, C, L, n_fft,
                                                      FILTER_TYPES[type], r, float(mu), _stream()))
    return W, T1


@_on_device
def filter_dual(W1, W2, Y, ref=0, n_fft=512, out_layout="TF", want_zn=True):
    """Single-node groups: z = w1^H y, zn = y[ref] - z and yf = w2^H y in ONE pass over Y (reference
    tango.py:369-376 and :445-450 with K = 1).  W1, W2 [..., F, C], Y [..., C, T, F] -> z, zn, yf
    [..., T, F] (or [..., F, T])."""
    _need(W1, torch.complex64, "W1")
    _need(W2, torch.complex64, "W2")
    _need(Y, torch.complex64, "Y")
    C, T, F = Y.shape[-3:]
    lead = tuple(Y.shape[:-3])
    G = Y.numel() // (C * T * F)
    if tuple(W1.shape) != lead + (F, C) or tuple(W2.shape) != lead + (F, C):
        raise ValueError("W1 / W2 shape %s / %s, expected %s" % (tuple(W1.shape), tuple(W2.shape), lead + (F, C)))
    lay = _layout(out_layout)
    shape = lead + ((T, F) if lay == TF else (F, T))
    z = torch.empty(shape, dtype=torch.complex64, device=Y.device)
    zn = torch.empty_like(z) if want_zn else None
    yf = torch.empty_like(z)
    _lib.check(_lib.load().disco_filter_dual(_ptr(W1), _ptr(W2), _ptr(Y), _ptr(z), _ptr(zn), _ptr(yf), int(ref), lay,
                                             G, C, T, n_fft, _stream()))
    return z, zn, yf


@_on_device
def tf_mask(S, N, type="irm1", bin_thr=0.0):
    """Oracle mask (reference dnn/utils.py:44-71) on device, float32.  Same shape as S."""
    _need(S, torch.complex64, "S")
    _need(N, torch.complex64, "N")
    if S.shape != N.shape:
        raise AssertionError("Input spectrograms should have the same shape.")   # sigproc_utils.py:71
    kind = type[:-1] if len(type) > 1 else ""
    if kind not in MASK_KINDS or not type[-1].isdigit():
        raise ValueError('Unknown mask type. Should be "irmX", "ibmX" or "iamX"')
    M = torch.empty(S.shape, dtype=torch.float32, device=S.device)
    _lib.check(_lib.load().disco_tf_mask(_ptr(S), _ptr(N), _ptr(M), S.numel(), MASK_KINDS[kind], int(type[-1]),
                                         float(bin_thr), _stream()))
    return M


def _sel(node_sel, K):
    if node_sel is None:
        return None, K, K
    arr = (ctypes.c_int * len(node_sel))(*[int(v) for v in node_sel])
    return arr, len(node_sel), len(node_sel)


def _z_dims(Z, B, T, F, z_layout):
    """K and the layout flag of the exchanged signals: 'BK' = [B, K, T, F], 'KB' = node-major [K, B, T, F]
    (what an all-gather over node-owning ranks delivers, disco_b200/dist.py)."""
    if z_layout not in ("BK", "KB"):
        raise ValueError("z_layout must be 'BK' or 'KB'")
    _need(Z, torch.complex64, "Z")
    K = Z.shape[1] if z_layout == "BK" else Z.shape[0]
    want = (B, K, T, F) if z_layout == "BK" else (K, B, T, F)
    if tuple(Z.shape) != want:
        raise ValueError("Z shape %s, expected %s" % (tuple(Z.shape), want))
    return K, (0 if z_layout == "BK" else 1)


@_on_device
def masked_scm(Y, mask, Z=None, n_fft=512, mask_layout="TF", node_sel=None, z_layout="BK"):
    """Y [B, Ksel, C, T, F], Z [B, K, T, F] (or [K, B, T, F] with z_layout='KB') or None (K = 1),
    mask [B, Ksel, T, F] / [B, Ksel, F, T] or None
    -> Rss, Rnn [B, Ksel, F, D, D], D = C + K - 1 (own mics, then z of the other nodes)."""
    _need(Y, torch.complex64, "Y")
    B, Ks, C, T, F = Y.shape
    K, zl = (1, 0) if Z is None else _z_dims(Z, B, T, F, z_layout)
    n_utt = B
    if Z is None:            # no exchange: every (b, k) is an independent single-node problem
        n_utt, sel, n_sel = B * Ks, None, 1
    else:
        sel, n_sel, _ = _sel(node_sel, K)
        if Ks != n_sel:
            raise ValueError("Y holds %d nodes, selection has %d" % (Ks, n_sel))
    lay = _layout(mask_layout)
    if mask is not None:
        _need(mask, torch.float32, "mask")
        want = (B, Ks, T, F) if lay == TF else (B, Ks, F, T)
        if tuple(mask.shape) != want:
            raise ValueError("mask shape %s, expected %s" % (tuple(mask.shape), want))
    D = C + K - 1
    Rss = torch.empty((B, Ks, F, D, D), dtype=torch.complex64, device=Y.device)
    Rnn = torch.empty_like(Rss)
    _lib.check(_lib.load().disco_masked_scm(_ptr(Y), _ptr(Z), _ptr(mask), lay, _ptr(Rss), _ptr(Rnn), n_utt, K, C, T,
                                            n_fft, sel, n_sel, zl, _stream()))
    return Rss, Rnn


@_on_device
def filter_sum_scm(W1, Y, mask, ref=0, n_fft=512, mask_layout="TF"):
    """Single-node groups (no exchange): z = w1^H y, zn = y[ref] - z AND the masked SCMs of Y under `mask`
    in one pass over Y.  W1 [..., F, C], Y [..., C, T, F], mask [..., T, F] (or [..., F, T]).
    Returns z, zn [..., T, F], Rss, Rnn [..., F, C, C]."""
    _need(W1, torch.complex64, "W1")
    _need(Y, torch.complex64, "Y")
    _need(mask, torch.float32, "mask")
    C, T, F = Y.shape[-3:]
    lead = Y.shape[:-3]
    G = Y.numel() // (C * T * F)
    lay = _layout(mask_layout)
    if tuple(W1.shape) != tuple(lead) + (F, C):
        raise ValueError("W1 shape %s, expected %s" % (tuple(W1.shape), tuple(lead) + (F, C)))
    want = tuple(lead) + ((T, F) if lay == TF else (F, T))
    if tuple(mask.shape) != want:
        raise ValueError("mask shape %s, expected %s" % (tuple(mask.shape), want))
    z = torch.empty(lead + (T, F), dtype=torch.complex64, device=Y.device)
    zn = torch.empty_like(z)
    Rss = torch.empty(lead + (F, C, C), dtype=torch.complex64, device=Y.device)
    Rnn = torch.empty_like(Rss)
    _lib.check(_lib.load().disco_filter_sum_scm(_ptr(W1), _ptr(Y), _ptr(mask), lay, _ptr(z), _ptr(zn), int(ref),
                                                _ptr(Rss), _ptr(Rnn), G, C, T, n_fft, _stream()))
    return z, zn, Rss, Rnn


def tango_mid_supported(C, K):
    return bool(_lib.load().disco_tango_mid_supported(int(C), int(K)))


@_on_device
def tango_mid(W1, Y, mask_w, ref=0, n_fft=512):
    """Multi-node arrays: z, zn of every node AND the step-2 SCMs of every node in one pass over Y.
    W1 [B, K, F, C], Y [B, K, C, T, F], mask_w [B, K, T, F] -> z, zn [B, K, T, F], Rss, Rnn [B, K, F, D, D]."""
    _need(W1, torch.complex64, "W1")
    _need(Y, torch.complex64, "Y")
    _need(mask_w, torch.float32, "mask_w")
    B, K, C, T, F = Y.shape
    D = C + K - 1
    if tuple(W1.shape) != (B, K, F, C) or tuple(mask_w.shape) != (B, K, T, F):
        raise ValueError("shape mismatch")
    z = torch.empty((B, K, T, F), dtype=torch.complex64, device=Y.device)
    zn = torch.empty_like(z)
    Rss = torch.empty((B, K, F, D, D), dtype=torch.complex64, device=Y.device)
    Rnn = torch.empty_like(Rss)
    _lib.check(_lib.load().disco_tango_mid(_ptr(W1), _ptr(Y), _ptr(mask_w), _ptr(z), _ptr(zn), int(ref), _ptr(Rss),
                                           _ptr(Rnn), B, K, C, T, n_fft, _stream()))
    return z, zn, Rss, Rnn


@_on_device
def mwf_solve(Rss, Rnn, mu=1.0, type="gevd", rank=1):
    """Batched intern_filter (reference internal_formulas.py:31-81).  Rss, Rnn [..., D, D] complex64
    -> W [..., D], t1 [..., D] complex64.  rank 'full'/'Full'/None -> all eigenpairs."""
    _need(Rss, torch.complex64, "Rss")
    _need(Rnn, torch.complex64, "Rnn")
    if type not in FILTER_TYPES:
        raise AttributeError("Unknown filter reference")       # internal_formulas.py:79
    D = Rss.shape[-1]
    n_mat = Rss.numel() // (D * D)
    r = 0 if rank in ("full", "Full", None) else int(rank)
    W = torch.empty(Rss.shape[:-1], dtype=torch.complex64, device=Rss.device)
    T1 = torch.empty_like(W)
    _lib.check(_lib.load().disco_mwf_solve(_ptr(Rss), _ptr(Rnn), _ptr(W), _ptr(T1), n_mat, D, FILTER_TYPES[type], r,
                                           float(mu), _stream()))
    return W, T1


@_on_device
def filter_sum(W, Y, Z=None, conj=True, ref=None, n_fft=512, out_layout="TF", node_sel=None, z_layout="BK"):
    """out = w^H x (conj=True) or w^T x over the concatenated channels [Y ; z of other nodes].
    W [B, Ksel, F, D]; Z [B, K, T, F] (or [K, B, T, F] with z_layout='KB');
    returns out (and resid = x[ref] - out when ref is given), [B, Ksel, T, F] or [.., F, T]."""
    _need(W, torch.complex64, "W")
    _need(Y, torch.complex64, "Y")
    B, Ks, C, T, F = Y.shape
    K, zl = (1, 0) if Z is None else _z_dims(Z, B, T, F, z_layout)
    n_utt = B
    if Z is None:
        n_utt, sel, n_sel = B * Ks, None, 1
    else:
        sel, n_sel, _ = _sel(node_sel, K)
        if Ks != n_sel:
            raise ValueError("Y holds %d nodes, selection has %d" % (Ks, n_sel))
    D = C + K - 1
    if tuple(W.shape) != (B, Ks, F, D):
        raise ValueError("W shape %s, expected %s" % (tuple(W.shape), (B, Ks, F, D)))
    lay = _layout(out_layout)
    shape = (B, Ks, T, F) if lay == TF else (B, Ks, F, T)
    out = torch.empty(shape, dtype=torch.complex64, device=Y.device)
    resid = torch.empty_like(out) if ref is not None else None
    _lib.check(_lib.load().disco_filter_sum(_ptr(W), 1 if conj else 0, _ptr(Y), _ptr(Z), _ptr(out), _ptr(resid),
                                            0 if ref is None else int(ref), lay, n_utt, K, C, T, n_fft, sel, n_sel,
                                            zl, _stream()))
    return (out, resid) if ref is not None else out


@_on_device
def istft(Y, length, n_fft=512):
    """Y [..., T, F] complex64 -> x [..., length] float32 (librosa istft semantics, center=True)."""
    _need(Y, torch.complex64, "Y")
    T, F = Y.shape[-2:]
    if F != n_fft // 2 + 1:
        raise ValueError("last dimension must be n_fft/2 + 1 bins")
    n_sig = Y.numel() // (T * F)
    x = torch.empty(Y.shape[:-2] + (int(length),), dtype=torch.float32, device=Y.device)
    _lib.check(_lib.load().disco_istft(_ptr(Y), _ptr(x), n_sig, T, int(length), n_fft, _stream()))
    return x


def _cat_dims(Y, Z, node_sel):
    B, Ks, C, T, F = Y.shape
    K = 1 if Z is None else Z.shape[1]
    if Z is None:
        return B * Ks, None, 1, K, C, T, F
    _need(Z, torch.complex64, "Z")
    sel, n_sel, _ = _sel(node_sel, K)
    if Ks != n_sel:
        raise ValueError("Y holds %d nodes, selection has %d" % (Ks, n_sel))
    return B, sel, n_sel, K, C, T, F


@_on_device
def scm_recursive(Y, mask, Z=None, lambda_cor=0.95, block=8, power=2, R0=None, n_fft=512, node_sel=None):
    """Exponentially smoothed SCM pair, R <- lambda R + (1 - lambda) w x x^H per frame (reference
    spatial_correlation_matrix, internal_formulas.py:84-103), sampled after every block of `block` frames.
    Y [B, Ksel, C, T, F], Z [B, K, T, F] or None, mask [B, Ksel, T, F] or None, R0 = (R0ss, R0nn) [B, Ksel, F, D, D]
    -> Rss, Rnn [B, Ksel, J, F, D, D], J = ceil(T / block)."""
    _need(Y, torch.complex64, "Y")
    if mask is not None:
        _need(mask, torch.float32, "mask")
    n_utt, sel, n_sel, K, C, T, F = _cat_dims(Y, Z, node_sel)
    B, Ks = Y.shape[:2]
    D = C + K - 1
    if mask is not None and tuple(mask.shape) != (B, Ks, T, F):
        raise ValueError("mask shape %s, expected %s" % (tuple(mask.shape), (B, Ks, T, F)))
    J = (T + block - 1) // block
    Rss = torch.empty((B, Ks, J, F, D, D), dtype=torch.complex64, device=Y.device)
    Rnn = torch.empty_like(Rss)
    r0s = r0n = None
    if R0 is not None:
        r0s, r0n = R0
        for r in (r0s, r0n):
            _need(r, torch.complex64, "R0")
            if tuple(r.shape) != (B, Ks, F, D, D):
                raise ValueError("R0 shape %s, expected %s" % (tuple(r.shape), (B, Ks, F, D, D)))
    _lib.check(_lib.load().disco_scm_recursive(_ptr(Y), _ptr(Z), _ptr(mask), _ptr(r0s), _ptr(r0n), _ptr(Rss), _ptr(Rnn),
                                               float(lambda_cor), int(block), int(power), n_utt, K, C, T, n_fft, sel,
                                               n_sel, _stream()))
    return Rss, Rnn


@_on_device
def filter_sum_blocks(W, Y, Z=None, block=8, lag=1, conj=True, ref=0, n_fft=512, node_sel=None):
    """One filter per block of frames: out[t] = W[t // block - lag]^H x[t] (pass-through of channel `ref` while no
    filter exists yet).  W [B, Ksel, J, F, D] -> out, resid = x[ref] - out, [B, Ksel, T, F]."""
    _need(W, torch.complex64, "W")
    _need(Y, torch.complex64, "Y")
    n_utt, sel, n_sel, K, C, T, F = _cat_dims(Y, Z, node_sel)
    B, Ks = Y.shape[:2]
    D, J = C + K - 1, (T + block - 1) // block
    if tuple(W.shape) != (B, Ks, J, F, D):
        raise ValueError("W shape %s, expected %s" % (tuple(W.shape), (B, Ks, J, F, D)))
    out = torch.empty((B, Ks, T, F), dtype=torch.complex64, device=Y.device)
    resid = torch.empty_like(out)
    _lib.check(_lib.load().disco_filter_sum_blocks(_ptr(W), 1 if conj else 0, _ptr(Y), _ptr(Z), _ptr(out), _ptr(resid),
                                                   int(ref), int(block), int(lag), n_utt, K, C, T, n_fft, sel, n_sel,
                                                   _stream()))
    return out, resid


@_on_device
def band_stats(x, ba, sel=None):
    """IIR filter bank + statistics of every band's output (reference metrics.py:96-110: lfilter, then np.var of
    the selected samples).  x [..., L] float32 (a time slice of a contiguous tensor is taken in place),
    ba [n_band, 2, order+1] float64 (b, a), sel optional like x -> stats [..., n_band, 3] float64: count, sum,
    sum of squares of the selected filter outputs."""
    for t, name in ((x, "x"), (sel, "sel")):
        if t is None:
            continue
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise TypeError("%s must be a CUDA tensor (disco_b200 has no CPU path)" % name)
        if t.dtype != torch.float32:
            raise TypeError("%s must be float32" % name)
    L, lead = x.shape[-1], tuple(x.shape[:-1])
    xv = x.reshape(-1, L)                      # a view when the rows are equally spaced
    if xv.stride(1) != 1 or (xv.shape[0] > 1 and xv.stride(0) < L):
        xv = xv.contiguous()
    ld = xv.stride(0) if xv.shape[0] > 1 else L
    sv = None
    if sel is not None:
        if tuple(sel.shape) != tuple(x.shape):
            raise ValueError("sel must have the shape of x")
        sv = torch.empty_strided(xv.shape, (ld, 1), dtype=torch.float32, device=x.device)
        sv.copy_(sel.reshape(-1, L))
    ba = ba.to(device=x.device, dtype=torch.float64).contiguous()
    if ba.dim() != 3 or ba.shape[1] != 2:
        raise ValueError("ba must be [n_band, 2, order + 1]")
    n_band, order = ba.shape[0], ba.shape[2] - 1
    stats = torch.empty(lead + (n_band, 3), dtype=torch.float64, device=x.device)
    _lib.check(_lib.load().disco_band_stats(_ptr(xv), _ptr(sv) if sv is not None else None, _ptr(ba), _ptr(stats),
                                            xv.shape[0], L, int(ld), n_band, order, _stream()))
    return stats


@_on_device
def transpose_last2(a):
    """[..., R, C] -> [..., C, R] (contiguous) for complex64 / float32 device tensors."""
    R, Cc = a.shape[-2:]
    batch = a.numel() // (R * Cc) if a.numel() else 0
    out = torch.empty(a.shape[:-2] + (Cc, R), dtype=a.dtype, device=a.device)
    lib = _lib.load()
    if a.dtype == torch.complex64:
        _need(a, torch.complex64, "a")
        _lib.check(lib.disco_transpose_c64(_ptr(a), _ptr(out), batch, R, Cc, _stream()))
    else:
        _need(a, torch.float32, "a")
        _lib.check(lib.disco_transpose_f32(_ptr(a), _ptr(out), batch, R, Cc, _stream()))
    return out


@_on_device
def apply_mask(X, m, one_minus=False):
    """m * X or (1 - m) * X, complex64 x float32.  Same shapes, or X [..., C, T, F] with one mask plane
    m [..., T, F] shared by the C channels of a group (one launch either way)."""
    _need(X, torch.complex64, "X")
    _need(m, torch.float32, "m")
    out = torch.empty_like(X)
    lib = _lib.load()
    if X.shape == m.shape:
        _lib.check(lib.disco_apply_mask(_ptr(X), _ptr(m), _ptr(out), X.numel(), 1 if one_minus else 0, _stream()))
    elif X.dim() == m.dim() + 1 and tuple(X.shape[:-3]) + tuple(X.shape[-2:]) == tuple(m.shape):
        plane = X.shape[-1] * X.shape[-2]
        _lib.check(lib.disco_apply_mask_channels(_ptr(X), _ptr(m), _ptr(out), m.numel() // plane, X.shape[-3], plane,
                                                 1 if one_minus else 0, _stream()))
    else:
        raise ValueError("shape mismatch")
    return out
