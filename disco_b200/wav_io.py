"""Minimal WAV codec for the dataset post-processing driver (dataset_post.py).

The reference reads and writes audio through the third-party `soundfile` package (libsndfile; pinned
SoundFile==0.10.3.post1 in requirements.txt, absent from this image):
    sf.read(path, dtype='float32')   post_generator.py:112,115; get_z_signals.py:84-89
    sf.write(path, data, fs)         post_generator.py:150-155  (default WAV subtype: PCM_16)
Restated from libsndfile's published conversion rules (parity unpinned for the codec itself):
    PCM16 -> float : x / 32768
    float -> PCM16 : lrint(x * 32767)   (round half to even, no clipping: out-of-range input is the caller's)
Only what the driver needs: mono / multi-channel 16-bit PCM.
"""
import wave

import numpy as np


def read(path, dtype="float32"):
    """-> (data [n] or [n, channels], sample rate)."""
    with wave.open(path, "rb") as w:
        if w.getsampwidth() != 2:
            raise ValueError("only 16-bit PCM WAV files are supported")
        fs, nch, n = w.getframerate(), w.getnchannels(), w.getnframes()
        raw = np.frombuffer(w.readframes(n), dtype="<i2")
    data = (raw.astype(np.float64) / 32768.0).astype(dtype)
    return (data.reshape(-1, nch) if nch > 1 else data), fs


def write(path, data, fs):
    data = np.asarray(data)
    nch = 1 if data.ndim == 1 else data.shape[1]
    pcm = np.rint(data.astype(np.float64) * 32767.0)           # np.rint: half to even, like lrint
    pcm = np.clip(pcm, -32768, 32767).astype("<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(nch)
        w.setsampwidth(2)
        w.setframerate(int(fs))
        w.writeframes(pcm.tobytes())
