"""DNN time-frequency mask estimation on the device: the step right before the beamformer
(SURVEY.md §8 f-1).  Mirrors

    prepare_data            disco_theque/speech_enhancement/utils.py:69-138
    CRNN / build_crnn       disco_theque/dnn/models/crnn.py:9-108, nn_structures.py:39-232
    get_mask ('crnn' path)  disco_theque/speech_enhancement/tango.py:209-215
    reshape_mask            disco_theque/speech_enhancement/tango.py:228-240

The network itself is stock PyTorch layers (cuDNN / cuBLAS -- library code, exactly as in the
reference); what is B200-specific here is that nothing leaves the GPU and nothing is blown up
21x: the reference builds one 21-frame window per STFT frame on the host
(``prepare_data`` -> (T, n_ch, 21, 257) float64 NumPy -> float32 -> .to('cuda')) and pushes every
window through the CNN.  The three Conv2d/BatchNorm2d/MaxPool2d((1,4)) blocks are translation-
invariant along time, so ``CRNN.forward_sequence`` runs the CNN ONCE over the padded spectrogram and
only unfolds the small feature map (64 x 4 per frame) into the 15-frame windows the GRU sees --
the same numbers, ~21x fewer convolution flops, and the mask comes out frame-major (T, F): the
layout the beamforming kernels read.

State-dict keys are identical to the reference's CRNN (``cnn.model.{0,1,3,4,6,7}.*``,
``rnn.model.0.rnn_layer.*``, ``ff.layers.0.*``), so trained checkpoints load unchanged
(tango.py:133-134 ``model.load_state_dict(saved_weights['model_state_dict'])``).
"""
import contextlib
import math

import numpy as np
import torch
from torch import nn

STFT_MIN, STFT_MAX = 1e-6, 1e3        # speech_enhancement/utils.py:7


def get_frames_to_pad(in_len, output_frames, out_len=None):
    """speech_enhancement/utils.py:13-33."""
    out_len = in_len if out_len is None else out_len
    if output_frames == "mid":
        return int(math.floor(in_len / 2)), int(math.floor(in_len / 2))
    if output_frames == "last":
        sel = (in_len + out_len) // 2
        return sel - 1, in_len - sel
    if output_frames == "all":
        return 0, 0
    raise ValueError(":param output_frames: should be 'mid', 'last' or 'all'")


def get_loss_frames(win_len, part):
    """dnn/utils.py:189-211."""
    if part == "all":
        return 0, win_len
    if part == "mid":
        first = int(math.ceil(win_len) / 2)
        return first, first + 1
    if part == "last":
        return win_len - 1, win_len
    if isinstance(part, int):
        return part, part + 1
    raise ValueError("Unknown argument value {}. It should be either 'all', 'mid' or 'last'.".format(part))


def normalization(x, norm_type=None, axis=0):
    """speech_enhancement/utils.py:36-66 on a device tensor (|x| clipped to [1e-6, 1e3])."""
    if norm_type == "pcen":
        raise NotImplementedError("PCEN normalisation needs librosa.pcen (third-party, absent)")
    x = x.abs().clamp(STFT_MIN, STFT_MAX)
    if norm_type == "scale_to_unit_norm":
        return x / torch.linalg.norm(x, dim=axis, keepdim=True)
    if norm_type == "scale_to_1":
        return x / torch.quantile(x, 0.99, dim=axis, keepdim=True)
    if norm_type == "center_and_scale":
        x = x - x.mean(dim=axis, keepdim=True)
        return x / x.std(dim=axis, keepdim=True, unbiased=False)
    return x


def _stack_inputs(y_data, z_data, norm_type, device):
    as_t = lambda a: a.to(device) if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a)).to(device)
    chans = [normalization(as_t(y_data), norm_type, axis=1)]
    if z_data is not None:
        chans += [normalization(as_t(z), norm_type, axis=1) for z in z_data]
    return torch.stack([c.to(torch.float32) for c in chans])          # [n_ch, F, T]


def prepare_data(y_data, three_d_tensor, z_data=None, win_len=21, win_hop=1, frame_to_pred="last",
                 norm_type=None, frames_lost=6, device="cuda"):
    """Reference signature (speech_enhancement/utils.py:69-138): sliding windows for the network,
    (n_samples, n_ch, win_len, n_freq) float32 on the device (three_d_tensor=False stacks the channels
    along frequency: (n_samples, win_len, n_ch * n_freq)).  y_data, z_data[i]: (F, T) spectrograms."""
    x = _stack_inputs(y_data, z_data, norm_type, device)
    pad = get_frames_to_pad(win_len, frame_to_pred, out_len=win_len - frames_lost)
    x = torch.nn.functional.pad(x, pad)                                # zero padding along time
    win = x.unfold(2, win_len, win_hop)                                # [n_ch, F, n_samples, win_len] (view)
    if three_d_tensor:
        return win.permute(2, 0, 3, 1).contiguous()                    # (n_samples, n_ch, win_len, F)
    n_ch, F = x.shape[:2]
    return win.permute(2, 3, 0, 1).reshape(win.shape[2], win_len, n_ch * F).contiguous()


class _Holder(nn.Module):
    """Gives a submodule the attribute path the reference's state dict uses (``<name>.model.*``)."""
    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, x):
        return self.model(x)


class _RNNSingle(nn.Module):
    """nn_structures.py:80-93: a recurrent layer that returns only its output sequence."""
    def __init__(self, cell, **kw):
        super().__init__()
        self.rnn_layer = getattr(nn, cell.upper())(**kw)

    def forward(self, x):
        return self.rnn_layer(x)[0]


class _FF(nn.Module):
    """nn_structures.py:39-77 with one activation name for all layers."""
    def __init__(self, input_size, units, activation):
        super().__init__()
        sizes = [input_size] + list(units)
        self.layers = nn.ModuleList(nn.Linear(sizes[i], sizes[i + 1]) for i in range(len(units)))
        self.activation = activation

    def forward(self, x):
        for layer in self.layers:
            x = getattr(torch, self.activation)(layer(x))
        return x


class CRNN(nn.Module):
    """The reference mask estimator (crnn.py:9-87 as configured at tango.py:127-132):
    3 x [Conv2d 3x3 pad (0,1) -> BatchNorm2d -> MaxPool2d (1,4)] -> GRU -> Linear + sigmoid."""

    def __init__(self, n_ch, win_len=21, n_freq=257, cnn_filters=(32, 64, 64), rnn_units=(256,), rnn_cell="GRU",
                 ff_units=(257,), ff_activation="sigmoid"):
        super().__init__()
        self.input_shape = (n_ch, win_len, n_freq)
        chans = [n_ch] + list(cnn_filters)
        layers, t, f = [], win_len, n_freq
        for i in range(len(cnn_filters)):
            layers += [nn.Conv2d(chans[i], chans[i + 1], 3, stride=1, padding=(0, 1)), nn.BatchNorm2d(chans[i + 1]),
                       nn.MaxPool2d((1, 4))]
            t, f = t - 2, f // 4
        self.cnn = _Holder(nn.Sequential(*layers))
        self.x_out, self.f_out = t, f
        rnn_layers, size = [], chans[-1] * f
        for u in rnn_units:
            rnn_layers.append(_RNNSingle(rnn_cell, input_size=size, hidden_size=u, num_layers=1, batch_first=True))
            size = u
        self.rnn = _Holder(nn.Sequential(*rnn_layers))
        self.ff = _FF(size, ff_units, ff_activation)

    def _head(self, x):
        """x (B, C, x_out, f_out) contiguous -> (B, x_out, n_freq).  The reference RE-VIEWS the CNN
        output as (B, time, C * f) without permuting (crnn.py:59) -- reproduced literally."""
        x = x.reshape(x.size(0), x.size(2), x.size(1) * x.size(-1))
        return self.ff(self.rnn(x))

    def forward(self, inp):
        """Window batch (n_samples, n_ch, win_len, n_freq) -> (n_samples, x_out, n_freq)  (crnn.py:55-63)."""
        if inp.dim() == 3:
            inp = inp.view(inp.size(0), 1, inp.size(1), inp.size(2))
        return self._head(self.cnn(inp).contiguous())

    def forward_sequence(self, x, win_hop=1, chunk=4096):
        """x [n_ch, F, T_padded] (already padded by get_frames_to_pad) -> (n_samples, x_out, n_freq), equal to
        forward() on every win_len-frame window but with the CNN run once over the whole sequence."""
        win_len = self.input_shape[1]
        feats = self.cnn(x.permute(0, 2, 1).unsqueeze(0))[0]            # [C, T_padded - (win_len - x_out), f_out]
        win = feats.unfold(1, self.x_out, win_hop)                      # [C, n_samples, f_out, x_out] (view)
        n = win.shape[1]
        assert n == 1 + (x.shape[2] - win_len) // win_hop
        out = []
        for lo in range(0, n, chunk):                                   # bound the unfolded copy
            w = win[:, lo:lo + chunk].permute(1, 0, 3, 2).contiguous()  # (n, C, x_out, f_out)
            out.append(self._head(w))
        return torch.cat(out, 0)

    @torch.no_grad()
    def forward_batch(self, x, frame_to_pred="mid", chunk=8192):
        """A whole batch of padded spectrograms at once: x [B, n_ch, F, T_padded] -> masks [B, n_samples, n_freq]
        (frame-major, float32), equal to reshape_mask(forward(windows), frame_to_pred) for every utterance.
        The CNN runs once per batch; the recurrent layers only run up to the predicted frame of each window (the
        later steps of a unidirectional GRU cannot influence it) and the output layer only on that frame --
        8 of 15 steps and 1 of 15 frames for the reference's 'mid' (tango.py:34-35)."""
        B = x.shape[0]
        win_len = self.input_shape[1]
        feats = self.cnn(x.permute(0, 1, 3, 2))                          # [B, C, T_padded - (win_len - x_out), f_out]
        Cc, Tf, fo = feats.shape[1:]
        n = Tf - self.x_out + 1
        assert n == 1 + (x.shape[3] - win_len)
        if frame_to_pred == "mid":
            last = int(math.floor(self.x_out / 2))                       # reshape_mask: floor(w/2) : ceil(w/2)
        elif frame_to_pred == "last":
            last = self.x_out - 1
        else:
            raise NotImplementedError("forward_batch predicts the 'mid' or the 'last' frame")
        win = feats.unfold(2, self.x_out, 1)                             # [B, C, n, f_out, x_out] (view)
        out = torch.empty((B, n, self.ff.layers[-1].out_features), dtype=torch.float32, device=feats.device)
        per = max(1, chunk // max(1, n))
        for b0 in range(0, B, per):                                      # bound the unfolded copy
            w = win[b0:b0 + per].permute(0, 2, 1, 4, 3).contiguous()     # (b, n, C, x_out, f_out)
            w = w.view(-1, self.x_out, Cc * fo)[:, :last + 1]            # the reference's literal re-view (crnn.py:59)
            h = self.rnn(w.contiguous())[:, last]                        # hidden state at the predicted frame
            out[b0:b0 + per] = self.ff(h).view(-1, n, out.shape[-1])
        return out

    def get_loss_frames(self, output_frames):
        """crnn.py:65-87."""
        win_in, win_out = self.input_shape[1], self.x_out
        if output_frames == "last":
            ff_in = (win_in + win_out) // 2 - 1
            lf_in = ff_in + 1
        elif output_frames == "mid":
            ff_in = int(math.ceil(win_in) / 2)
            lf_in = ff_in + 1
        elif output_frames == "all":
            ff_in, lf_in = (win_in - win_out) // 2, (win_in + win_out) // 2
        else:
            raise ValueError("Unknown argument value {}. It should be either 'all', 'mid' or 'last'."
                             .format(output_frames))
        return (ff_in, lf_in), get_loss_frames(win_out, output_frames)


def build_crnn(n_ch, **kw):
    """The estimator exactly as tango.py:127-132 configures it; returns the module only."""
    return CRNN(n_ch, **kw)


def reshape_mask_device(m_stack, output_frame="last"):
    """tango.py:228-240 on a device tensor, WITHOUT the final transpose: (n_samples, F) frame-major."""
    if output_frame == "last":
        return m_stack[:, -1, :]
    if output_frame == "mid":
        w = m_stack.shape[1]
        return m_stack[:, int(math.floor(w / 2)):int(math.ceil(w / 2)), :].squeeze(1)
    if output_frame == "all":
        raise NotImplementedError("This case was not implemented yet")
    raise ValueError(":param output_frame: should be either 'last', 'all' or 'mid'")


@contextlib.contextmanager
def fp32_exact(on=True):
    """Inside: cuDNN convolutions and cuBLAS matmuls in IEEE float32 (PyTorch's default lets cuDNN use TF32,
    whose 10-bit mantissa moves the sigmoid outputs by ~1e-4 -- more than the beamformer's parity budget)."""
    if not on:
        yield
        return
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@torch.no_grad()
def estimate_mask(mod, y_spec, z_specs=None, win_len=21, win_hop=1, frame_to_pred="mid", norm_type=None,
                  device="cuda", exact=True):
    """get_mask(..., mask_type='crnn') of tango.py:209-215, on the device.
    y_spec (F, T) mixture STFT (or magnitude) of the reference microphone, z_specs list of (F, T) compressed
    signals of the other nodes (step 2).  Returns the mask FRAME-MAJOR (T, F) float32, ready for the kernels.
    exact=True (default) runs the network in IEEE float32 like the reference's CPU path; exact=False leaves
    PyTorch's TF32 defaults on (faster convolutions, masks within ~3e-4)."""
    with fp32_exact(exact):
        return _estimate_mask(mod, y_spec, z_specs, win_len, win_hop, frame_to_pred, norm_type, device)


@torch.no_grad()
def estimate_masks_batch(mod, Y_ref, win_len=21, frame_to_pred="mid", exact=True):
    """Masks of a whole batch in a few launches: Y_ref [B, T, F] complex64 / float32 device tensor (frame-major
    STFT of the reference microphone of single-node arrays, no compressed signals) -> [B, T, F] float32.
    Same numbers as estimate_mask() per utterance (prepare_data's clipping and zero padding, tango.py:209-215)."""
    mod.eval()
    frames_lost = int(win_len - mod.get_loss_frames("last")[-1][-1])
    x = Y_ref.abs().clamp(STFT_MIN, STFT_MAX).to(torch.float32).transpose(-1, -2).unsqueeze(1)     # [B, 1, F, T]
    x = torch.nn.functional.pad(x, get_frames_to_pad(win_len, frame_to_pred, out_len=win_len - frames_lost))
    with fp32_exact(exact):
        return mod.forward_batch(x, frame_to_pred)


def _estimate_mask(mod, y_spec, z_specs, win_len, win_hop, frame_to_pred, norm_type, device):
    mod.eval()
    frames_lost = int(win_len - mod.get_loss_frames("last")[-1][-1])
    x = _stack_inputs(y_spec, z_specs, norm_type, device)
    x = torch.nn.functional.pad(x, get_frames_to_pad(win_len, frame_to_pred, out_len=win_len - frames_lost))
    if hasattr(mod, "forward_sequence"):
        m_stack = mod.forward_sequence(x, win_hop)
    else:                                     # any module with the reference's window interface
        m_stack = mod(x.unfold(2, win_len, win_hop).permute(2, 0, 3, 1).contiguous())
    return reshape_mask_device(m_stack, frame_to_pred).contiguous()
