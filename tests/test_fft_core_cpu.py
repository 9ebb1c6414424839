"""The in-register DFT of the STFT / iSTFT kernels (csrc/fft_reg.cuh: bit reversal by register renaming, twiddle
selection, three-packed-instruction butterflies) compiled for the HOST with the packed-FP32 intrinsics restated as
plain float arithmetic (fmaf per half), against numpy's FFT.  Pins the transform's structure without a GPU."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "disco_b200", "csrc")

SHIM = r"""
#pragma once
#include <cmath>
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 v; v.x = x; v.y = y; return v; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#define DISCO_DEV static inline
"""

MAIN = r"""
#include <cstdio>
#include "fft_reg.cuh"
template <int R, bool INV>
static void run(const float* in) {
    float2 v[R];
    for (int i = 0; i < R; ++i) v[i] = make_float2(in[2 * i], in[2 * i + 1]);
    disco::dft_reg<R, INV>(v);
    for (int i = 0; i < R; ++i) printf("%.9g %.9g\n", v[i].x, v[i].y);
}
int main() {
    float in[64];
    for (int i = 0; i < 64; ++i) { if (scanf("%f", &in[i]) != 1) return 1; }
    run<2, false>(in); run<4, false>(in); run<8, false>(in); run<16, false>(in); run<32, false>(in);
    run<8, true>(in); run<16, true>(in); run<32, true>(in);
    return 0;
}
"""


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs a host C++ compiler")
def test_register_dft_matches_numpy(tmp_path):
    for fn in ("fft_reg.cuh", "tw32.cuh"):
        shutil.copy(os.path.join(CSRC, fn), tmp_path / fn)
    (tmp_path / "common.cuh").write_text(SHIM)          # shadows csrc/common.cuh for the copied headers
    (tmp_path / "main.cpp").write_text(MAIN)
    exe = tmp_path / "dft"
    # -ffp-contract=off: only the fmaf calls of the shim fuse, as on the device
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-x", "c++", str(tmp_path / "main.cpp"), "-o", str(exe)],
                   check=True, cwd=tmp_path)
    rng = np.random.default_rng(7)
    x = rng.standard_normal(64).astype(np.float32)
    out = subprocess.run([str(exe)], input=" ".join("%.9g" % v for v in x), capture_output=True, text=True, check=True).stdout
    vals = np.array([[float(a) for a in line.split()] for line in out.strip().splitlines()])
    got = vals[:, 0] + 1j * vals[:, 1]
    z = x[0::2].astype(np.float64) + 1j * x[1::2].astype(np.float64)
    pos = 0
    for R, inv in ((2, False), (4, False), (8, False), (16, False), (32, False), (8, True), (16, True), (32, True)):
        ref = np.fft.ifft(z[:R]) * R if inv else np.fft.fft(z[:R])
        err = np.linalg.norm(got[pos:pos + R] - ref) / np.linalg.norm(ref)
        assert err < 5e-7, (R, inv, err)
        pos += R
    assert pos == len(got)
