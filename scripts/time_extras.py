"""CUDA-event timings of the steps around the beamformer (rows f-2 .. f-4) at BASELINE-like sizes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from disco_b200 import online, ops, post  # noqa: E402
from scripts.ab_wide import timeit  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    B, K, C, L, T, F = 64, 4, 4, 160000, 626, 257
    cplx = lambda *s: torch.view_as_complex(torch.randn(*s, 2, device=dev, generator=g))
    # f-2: six iSTFTs per node + frequency-weighted SNR / SD of every node
    specs = {n: cplx(B, K, T, F) for n in ("yf", "z_y", "sf", "nf", "z_s", "z_n")}
    us = timeit(lambda: post.to_time(specs, L, layout="TF"), n=10)
    print("to_time: 6 x %d iSTFTs of 10 s: %.1f us" % (B * K, us), flush=True)
    s, n = torch.randn(B, K, L, device=dev, generator=g), torch.randn(B, K, L, device=dev, generator=g)
    us = timeit(lambda: post.fw_snr(s, n, 16000), n=5, warm=1)
    print("fw_snr: 2 x %d signals x 17 bands x %d samples: %.1f us (%.2f G samples*bands/s)"
          % (B * K, L, us, 2 * B * K * 17 * L / us / 1e3), flush=True)
    us = timeit(lambda: post.si_sdr(s, n), n=10)
    print("si_sdr: %d signals: %.1f us" % (B * K, us), flush=True)
    # f-4: recursive two-step Tango on the cfg 2 shape (64 x 4 mics x 10 s), blocks of 8 frames
    y = torch.randn(B, 1, C, L, device=dev, generator=g)
    mz, mw = torch.rand(B, 1, T, F, device=dev, generator=g), torch.rand(B, 1, T, F, device=dev, generator=g)
    us = timeit(lambda: online.online_tango(y, (mz, mw), block=8, lag=1), n=10)
    print("online_tango cfg2 shape, block 8: %.1f us (%.1f M frames/s)" % (us, B * T / us), flush=True)
    Y = ops.stft(y)
    us = timeit(lambda: ops.scm_recursive(Y, mz, None, 0.95, 8), n=10)
    print("  scm_recursive alone: %.1f us" % us, flush=True)


if __name__ == "__main__":
    main()
