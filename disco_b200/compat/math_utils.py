"""my_stft / my_istft with the reference signatures (disco_theque/math_utils.py:134-152)."""
import torch

from .. import ops
from ._util import dev


def my_stft(x):
    """librosa STFT "with the parameters that I always use" (n_fft=512, hop=256, center=True):
    x (L,) -> (257, T) complex64.  Reference math_utils.py:134-140."""
    Y = ops.stft(dev(x, torch.float32).reshape(1, -1), 512)          # [1, T, F]
    return ops.transpose_last2(Y)[0].cpu().numpy()


def my_istft(y, out_len):
    """Inverse of my_stft: y (257, T) complex -> (out_len,) float32.  Reference math_utils.py:143-152."""
    Yt = ops.transpose_last2(dev(y, torch.complex64)[None])        # (F, T) -> [1, T, F]
    return ops.istft(Yt, int(out_len), 512)[0].cpu().numpy()
