"""NumPy restatement of ``librosa.core.stft`` / ``librosa.core.istft`` (librosa <= 0.9
defaults) — TEST INFRASTRUCTURE, see oracle/__init__.py.

librosa is a third-party dependency of the reference (imported at
disco_theque/speech_enhancement/tango.py:12, math_utils.py:140,152 — the latter
without an import statement) and is neither vendored under /root/reference nor
pinned in requirements.txt:1-11.  The conda env name in exp/ex1/loop_tango.sh:5
dates it to librosa 0.7/0.8, whose published algorithm is restated here:

stft(y, n_fft, hop_length, center=True):                
    window  = scipy.signal.get_window('hann', n_fft, fftbins=True)   (periodic Hann)
    y       = np.pad(y, n_fft // 2, mode='reflect')
    frames  = y[t*hop : t*hop + n_fft] * window,  t = 0 .. (len(y) - n_fft) // hop
    out     = rfft(frames)  -> (1 + n_fft/2, n_frames), complex64 for float32 input

istft(S, hop_length, win_length, center=True, length):  
    per frame irfft -> * window -> overlap-add; divide by the overlap-added
    squared window where it exceeds tiny(float32); drop n_fft//2 samples at the
    start; crop / zero-pad to ``length``.

Parity status: UNPINNED against librosa itself (absent).  Cross-checked in
tests/test_oracle.py against torch.stft / torch.istft (same conventions).
"""
import numpy as np


def hann_periodic(n_fft, dtype=np.float64):
    """scipy.signal.get_window('hann', n_fft, fftbins=True)."""
    n = np.arange(n_fft, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)).astype(dtype)


def n_frames_of(length, n_fft=512, hop=256):
    """1 + L // hop for center=True (== 3 + floor((L - n_fft) / hop), tango.py:287)."""
    return 1 + (length + 2 * (n_fft // 2) - n_fft) // hop


def stft(y, n_fft=512, hop_length=256, center=True, pad_mode="reflect", dtype=None):
    y = np.asarray(y)
    if dtype is None:
        dtype = np.complex64 if y.dtype == np.float32 else np.complex128
    win = hann_periodic(n_fft)
    if center:
        y = np.pad(y, n_fft // 2, mode=pad_mode)
    n_frames = 1 + (len(y) - n_fft) // hop_length
    idx = np.arange(n_fft)[None, :] + hop_length * np.arange(n_frames)[:, None]
    # librosa multiplies the (float32) frames by the float64 scipy window, so the
    # product and the FFT run in float64; only the store into the preallocated
    # ``dtype`` (complex64) matrix rounds.
    frames = y[idx] * win[None, :]
    spec = np.fft.rfft(frames, axis=1)
    return np.ascontiguousarray(spec.T.astype(dtype))


def istft(S, hop_length=256, win_length=512, center=True, length=None, dtype=np.float32):
    S = np.asarray(S)
    n_fft = 2 * (S.shape[0] - 1)
    win = hann_periodic(win_length)
    if win_length != n_fft:  # librosa pads the window to n_fft, centred
        lpad = (n_fft - win_length) // 2
        win = np.pad(win, (lpad, n_fft - win_length - lpad))
    n_frames = S.shape[1]
    if length is not None:
        padded = length + (n_fft if center else 0)
        n_frames = min(n_frames, int(np.ceil(padded / hop_length)))
    expected = n_fft + hop_length * (n_frames - 1)
    y = np.zeros(expected, dtype=np.float64)
    wss = np.zeros(expected, dtype=np.float64)
    seg = np.fft.irfft(S[:, :n_frames].astype(np.complex128), n=n_fft, axis=0)
    w2 = win * win
    for t in range(n_frames):
        a = t * hop_length
        y[a:a + n_fft] += win * seg[:, t]
        wss[a:a + n_fft] += w2
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    if center:
        y = y[n_fft // 2:]
        if length is None:
            y = y[:len(y) - n_fft // 2]
    if length is not None:
        if len(y) > length:
            y = y[:length]
        elif len(y) < length:
            y = np.pad(y, (0, length - len(y)))
    return y.astype(dtype)
