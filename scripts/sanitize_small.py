"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from disco_b200 import ops
from disco_b200.synth import make_batch
from disco_b200.tango import tango_batched
dev = torch.device("cuda:0")
for (B, K, C, L, n_fft) in ((3, 2, 4, 9000, 512), (2, 1, 3, 5003, 256), (2, 3, 2, 6000, 1024)):
    y, s, n = make_batch(B, K, C, L, seed0=1)
    out = tango_batched(torch.from_numpy(y).to(dev), torch.from_numpy(s).to(dev), torch.from_numpy(n).to(dev), n_fft=n_fft)
    x = ops.istft(ops.stft(torch.from_numpy(y).to(dev), n_fft), L, n_fft)
    torch.cuda.synchronize()
    print("ok", B, K, C, L, n_fft, float(out["yf"].abs().mean()), float((x.cpu() - torch.from_numpy(y)).abs().max()))

# wide-channel engine (cp.async-staged), Nyquist block, fused z, ragged T; filter bank; online kernels
from disco_b200 import online, post
rng = np.random.default_rng(0)
cplx = lambda *s: torch.from_numpy((rng.standard_normal(s) + 1j * rng.standard_normal(s)).astype(np.complex64)).to(dev)
for (B, K, C, T, n_fft) in ((2, 1, 8, 37, 512), (2, 4, 4, 21, 256), (1, 8, 2, 9, 512), (2, 3, 13, 33, 256), (1, 2, 4, 300, 256)):
    F = n_fft // 2 + 1
    Y, W, Z = cplx(B, K, C, T, F), cplx(B, K, F, C), (cplx(B, K, T, F) if K > 1 else None)
    m = torch.rand(B, K, T, F, device=dev)
    Rs, Rn = ops.masked_scm(Y, m, Z, n_fft=n_fft)
    if K == 1:
        ops.filter_sum_scm(W, Y, m, ref=1, n_fft=n_fft)
    elif ops.tango_mid_supported(C, K):
        ops.tango_mid(W, Y, m, ref=0, n_fft=n_fft)
    if C + K - 1 <= 8:
        o = online.online_mwf(Y, m, Z, block=4, lag=1, n_fft=n_fft)
    torch.cuda.synchronize()
    print("ok wide", B, K, C, T, n_fft, float(Rs.abs().mean()))
x = torch.randn(37, 3000, device=dev)
print("ok bank", float(post.fw_snr(x[:, 100:], 0.5 * torch.randn(37, 2900, device=dev), 16000)[1].mean()))
torch.cuda.synchronize()
