"""First-contact check of the round-2 kernels on a GPU box: correctness against torch float64 references
computed on the device (fast), then CUDA-event timings of every kernel at the BASELINE shapes.

    python scripts/quick_check.py [--no-time]

Prints one line per check; exits non-zero on the first failed check (timings are still attempted).
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from disco_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
FAIL = []


def rel(a, b):
    a, b = a.to(torch.complex128), b.to(torch.complex128)
    return (torch.linalg.norm((a - b).reshape(-1)) / torch.linalg.norm(b.reshape(-1))).item()


def check(name, err, tol):
    ok = err < tol
    print("%-58s err %.2e  tol %.0e  %s" % (name, err, tol, "ok" if ok else "FAIL"), flush=True)
    if not ok:
        FAIL.append(name)


def ref_stft(x, n_fft):
    w = torch.hann_window(n_fft, periodic=True, dtype=torch.float64, device=x.device)
    Y = torch.stft(x.double().reshape(-1, x.shape[-1]), n_fft, n_fft // 2, window=w, center=True, pad_mode="reflect",
                   return_complex=True)
    return Y.transpose(-1, -2).reshape(x.shape[:-1] + (Y.shape[-1], Y.shape[-2]))      # [..., T, F]


def ref_scm(Y, m):
    """Y [G, C, T, F] c128, m [G, T, F] -> Rss, Rnn [G, F, C, C] (mean over T of (m y)(m y)^H)."""
    T = Y.shape[2]
    s, n = Y * m[:, None], Y * (1 - m)[:, None]
    return torch.einsum("gitf,gjtf->gfij", s, s.conj()) / T, torch.einsum("gitf,gjtf->gfij", n, n.conj()) / T


def correctness():
    g = torch.Generator(device="cpu").manual_seed(0)
    for n_fft in (512, 256, 1024):
        for (G, C, L) in ((3, 4, 9000), (2, 8, 9000), (70, 6, 6000), (2, 5, 5003), (5, 1, 4000), (2, 3, 7001)):
            if not ops.stft_scm_supported(n_fft, C, 1):
                continue
            x = torch.randn((G, C, L), generator=g).to(dev)
            T, F = 1 + L // (n_fft // 2), n_fft // 2 + 1
            m = torch.rand((G, T, F), generator=g).to(dev)
            Y, Rss, Rnn = ops.stft_scm(x, m, n_fft)
            torch.cuda.synchronize()
            Yr = ref_stft(x, n_fft)
            Rs, Rn = ref_scm(Yr, m.double())
            tag = "stft_scm n_fft=%d G=%d C=%d L=%d" % (n_fft, G, C, L)
            check(tag + " Y", rel(Y, Yr), 2e-6)
            check(tag + " Rss", rel(Rss, Rs), 3e-6)
            check(tag + " Rnn", rel(Rnn, Rn), 3e-6)
            Y0 = ops.stft(x, n_fft)       # groups signals by 4: same two-for-one partner only when C is 4 or 8
            check(tag + " stft vs fused Y", rel(Y0, Y), 1e-12 if C in (4, 8) else 5e-7)
            if ops.stft_scm_supported(n_fft, C, 2):
                m2 = torch.rand((G, T, F), generator=g).to(dev)
                Y2, ws = ops.stft_scm2(x, m, m2, n_fft)
                Ra, _ = ops.scm_from_workspace(ws, G, C, L, n_fft, n_set=2, set=0)
                Rb, Rbn = ops.scm_from_workspace(ws, G, C, L, n_fft, n_set=2, set=1)
                Rs2, Rn2 = ref_scm(Yr, m2.double())
                check(tag + " scm2 set0 == single", rel(Ra, Rss), 1e-12)
                check(tag + " scm2 set1 Rss", rel(Rb, Rs2), 3e-6)
                check(tag + " scm2 set1 Rnn", rel(Rbn, Rn2), 3e-6)
                check(tag + " scm2 Y", rel(Y2, Y), 1e-12)
    # dual filter
    for C in (1, 3, 4):
        B, T, F = 3, 77, 257
        cplx = lambda *s: torch.complex(torch.randn(s, generator=g), torch.randn(s, generator=g)).to(dev)
        Y, W1, W2 = cplx(B, 1, C, T, F), cplx(B, 1, F, C), cplx(B, 1, F, C)
        for lay in ("TF", "FT"):
            z, zn, yf = ops.filter_dual(W1, W2, Y, ref=C - 1, out_layout=lay)
            want = torch.einsum("bkfc,bkctf->bktf", W2.conj().to(torch.complex128), Y.to(torch.complex128))
            wz = torch.einsum("bkfc,bkctf->bktf", W1.conj().to(torch.complex128), Y.to(torch.complex128))
            tr = (lambda a: a) if lay == "TF" else (lambda a: a.transpose(-1, -2))
            check("filter_dual C=%d %s yf" % (C, lay), rel(tr(yf), want), 1e-6)
            check("filter_dual C=%d %s z" % (C, lay), rel(tr(z), wz), 1e-6)
            check("filter_dual C=%d %s zn" % (C, lay), rel(tr(zn), Y[:, :, C - 1].to(torch.complex128) - wz), 1e-6)
    # whole K=1 path: dual route vs the two-pass route
    from disco_b200.tango import tango_batched, tango_step1
    B, K, C, L = 4, 1, 4, 16000
    y = torch.randn((B, K, C, L), generator=g).to(dev) * 0.1
    T, F = 1 + L // 256, 257
    mz, mw = torch.rand((B, K, T, F), generator=g).to(dev), torch.rand((B, K, T, F), generator=g).to(dev)
    out = tango_batched(y, masks=(mz, mw), out_layout="TF", diagnostics=False)
    st1 = tango_step1(y, mz, apply_filter=False)
    z, zn, Rss2, Rnn2 = ops.filter_sum_scm(st1["W1"], st1["Y"], mw)
    W2, _ = ops.mwf_solve(Rss2, Rnn2)
    yf = ops.filter_sum(W2, st1["Y"], None)
    check("tango K=1 dual route vs two-pass route: yf", rel(out["yf"].abs(), yf.abs()), 1e-5)
    check("tango K=1 dual route vs two-pass route: z", rel(out["z_y"], z), 1e-6)
    check("tango K=1 dual route vs two-pass route: zn", rel(out["zn"], zn), 1e-6)


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3      # us


def timings():
    peak = 6571.9
    try:
        import json
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    g = torch.Generator(device="cpu").manual_seed(1)
    L = 160000
    print("\n%-44s %10s %10s %8s" % ("kernel (shape)", "us", "GB/s", "frac"))

    def row(name, us, nbytes):
        gbs = nbytes / us / 1e3
        print("%-44s %10.1f %10.0f %8.3f" % (name, us, gbs, gbs / peak), flush=True)

    for n_fft, G, C in ((512, 64, 4), (256, 64, 4), (1024, 64, 4), (512, 128, 8), (256, 128, 8), (512, 128, 2), (512, 64, 6)):
        T, F = 1 + L // (n_fft // 2), n_fft // 2 + 1
        x = torch.randn((G, C, L), generator=g).to(dev)
        m, m2 = torch.rand((G, T, F), generator=g).to(dev), torch.rand((G, T, F), generator=g).to(dev)
        base = G * (4 * C * L + 8 * C * F * T)
        row("stft n_fft=%d G=%d C=%d" % (n_fft, G, C), timeit(lambda: ops.stft(x, n_fft)), base)
        if ops.stft_scm_supported(n_fft, C, 1):
            row("stft_scm<%d,%d,1> G=%d" % (n_fft, C, G), timeit(lambda: ops.stft_scm(x, m, n_fft, keep_partials=True)),
                base + G * (4 * F * T + 16 * F * C * C))
        if ops.stft_scm_supported(n_fft, C, 2):
            row("stft_scm<%d,%d,2> G=%d" % (n_fft, C, G), timeit(lambda: ops.stft_scm2(x, m, m2, n_fft)),
                base + G * (8 * F * T + 32 * F * C * C))
        Y = ops.stft(x, n_fft).view(G, 1, C, T, F)
        cplx = lambda *s: torch.complex(torch.randn(s, generator=g), torch.randn(s, generator=g)).to(dev)
        W = cplx(G, 1, F, C)
        mm = m.view(G, 1, T, F)
        row("masked_scm D=%d (+z, zn) G=%d n_fft=%d" % (C, G, n_fft), timeit(lambda: ops.filter_sum_scm(W, Y, mm, n_fft=n_fft)),
            G * (8 * C * F * T + 4 * F * T + 16 * F * T + 16 * F * C * C))
        row("masked_scm D=%d G=%d n_fft=%d" % (C, G, n_fft), timeit(lambda: ops.masked_scm(Y, mm, None, n_fft)),
            G * (8 * C * F * T + 4 * F * T + 16 * F * C * C))
        row("filter_sum D=%d G=%d n_fft=%d" % (C, G, n_fft), timeit(lambda: ops.filter_sum(W, Y, None, n_fft=n_fft)),
            G * (8 * C * F * T + 8 * F * T))
        if C <= 4:
            W2 = cplx(G, 1, F, C)
            row("filter_dual C=%d G=%d n_fft=%d" % (C, G, n_fft), timeit(lambda: ops.filter_dual(W, W2, Y, n_fft=n_fft)),
                G * (8 * C * F * T + 24 * F * T))
            _, ws = ops.stft_scm(x, m, n_fft, keep_partials=True)
            row("mwf_solve<%d> partials n=%d" % (C, G * F), timeit(lambda: ops.mwf_solve_workspace(ws, G, C, L, n_fft)), G * F * (16 * C * C + 16 * C))
            if ops.stft_scm_supported(n_fft, C, 2):
                _, ws2 = ops.stft_scm2(x, m, m2, n_fft)
                row("mwf_solve<%d> partials, 2 sets n=%d" % (C, 2 * G * F), timeit(lambda: ops.mwf_solve_workspace2(ws2, G, C, L, n_fft)),
                    2 * G * F * (16 * C * C + 16 * C))
        else:
            _, Rss, Rnn = ops.stft_scm(x, m, n_fft) if ops.stft_scm_supported(n_fft, C, 1) else (None,) + ops.masked_scm(Y, mm, None, n_fft)
            row("mwf_solve<%d> n=%d" % (C, G * F), timeit(lambda: ops.mwf_solve(Rss, Rnn)), G * F * (16 * C * C + 16 * C))
        Yc = Y[:, 0, 0].contiguous()
        row("istft n_fft=%d n_sig=%d" % (n_fft, G), timeit(lambda: ops.istft(Yc, L, n_fft)), G * (8 * F * T + 4 * L))
        del x, Y, m, m2, mm, W, Yc
        torch.cuda.empty_cache()


if __name__ == "__main__":
    ops.init(512)
    correctness()
    print("FAILED: %s" % FAIL if FAIL else "all checks passed", flush=True)
    if "--no-time" not in sys.argv:
        timings()
    sys.exit(1 if FAIL else 0)
