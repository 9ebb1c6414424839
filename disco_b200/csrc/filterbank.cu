// IIR filter bank + band statistics for the frequency-weighted metrics that follow the beamformer
// (reference disco_theque/metrics.py:96-110 fw_snr and :264-270 fw_sd: per third-octave band
// `scipy.signal.lfilter(b[i], a[i], x)` on the whole signal, then `np.var` of the selected output samples).
//
// The recurrence (direct form II transposed, float64, the arithmetic AND rounding sequence of scipy's lfilter) is serial in
// time, so the parallelism is (signal, band): lanes <-> 32 different signals, warps <-> bands.  A CTA
// streams chunks of the 32 signals through a double-buffered, padded shared-memory tile (coalesced
// global loads, conflict-free per-lane reads) that all its band-warps share; each thread keeps its
// filter's 2 * NC coefficients and NC - 1 delays in registers and accumulates count / sum / sum of squares
// of its outputs.  Nothing but 3 numbers per (signal, band) is written.
#include "kernels.h"
#include "scm_core.cuh"

namespace disco {

constexpr int kBankChunk = 64;   // samples per shared-memory chunk
constexpr int kBankWarps = 8;    // at most this many bands per CTA

template <int NC>   // coefficients per polynomial (filter order + 1)
__global__ void __launch_bounds__(32 * kBankWarps) band_stats_kernel(BankArgs a) {
    __shared__ float xs[2][32][kBankChunk + 1];
    __shared__ float ss[2][32][kBankChunk + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int sig0 = blockIdx.x * 32;
    const int band = blockIdx.y * nwarp + warp;
    const bool live = band < a.n_band;            // warp-uniform
    const bool has_sel = a.sel != nullptr;

    double b[NC], a_[NC], z[NC - 1];
    {
        const double* q = a.ba + (size_t)(live ? band : 0) * 2 * NC;
        const double a0 = q[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            b[i] = q[i] / a0;                     // lfilter normalises by a[0]
            a_[i] = q[NC + i] / a0;
        }
#pragma unroll
        for (int i = 0; i < NC - 1; ++i) z[i] = 0.0;
    }
    double cnt = 0.0, sum = 0.0, sq = 0.0;

    // cooperative asynchronous chunk load (cp.async, zero fill out of range):
    // element e = threadIdx.x + j * blockDim.x -> (row e / CH, column e % CH)
    auto load_chunk = [&](int c, int buf) {
        for (int e = threadIdx.x; e < 32 * kBankChunk; e += blockDim.x) {
            const int r = e / kBankChunk, col = e % kBankChunk;
            const int n = c * kBankChunk + col;
            const bool ok = sig0 + r < a.n_sig && n < a.L;
            const size_t o = ok ? (size_t)(sig0 + r) * a.ldx + n : 0;
            cp_async4(&xs[buf][r][col], a.x + o, ok);
            if (has_sel) cp_async4(&ss[buf][r][col], a.sel + o, ok);
        }
        cp_async_commit();
    };
    const int nchunk = (a.L + kBankChunk - 1) / kBankChunk;
    load_chunk(0, 0);
    for (int c = 0; c < nchunk; ++c) {
        const int buf = c & 1;
        cp_async_wait<0>();
        __syncthreads();                                  // chunk c visible; everyone is done with chunk c-1
        if (c + 1 < nchunk) load_chunk(c + 1, buf ^ 1);   // in flight while chunk c is filtered
        if (live) {
            const int nval = min(kBankChunk, a.L - c * kBankChunk);
            for (int n = 0; n < nval; ++n) {
                const double x = (double)xs[buf][lane][n];
                // scipy's evaluation order with every product and sum rounded separately (no FMA contraction):
                // the order-8 band-passes of the lowest bands amplify rounding differences by ~1e10, so matching
                // the reference's numbers means matching its arithmetic (x86-64 SciPy builds do not contract)
                const double y = __dadd_rn(z[0], __dmul_rn(b[0], x));
#pragma unroll
                for (int i = 0; i < NC - 2; ++i)
                    z[i] = __dsub_rn(__dadd_rn(z[i + 1], __dmul_rn(x, b[i + 1])), __dmul_rn(y, a_[i + 1]));
                z[NC - 2] = __dsub_rn(__dmul_rn(x, b[NC - 1]), __dmul_rn(y, a_[NC - 1]));
                const bool take = has_sel ? (ss[buf][lane][n] != 0.f) : (y != 0.0);
                if (take) {
                    cnt += 1.0;
                    sum += y;
                    sq = fma(y, y, sq);
                }
            }
        }
    }
    if (live && sig0 + lane < a.n_sig) {
        double* o = a.stats + ((size_t)(sig0 + lane) * a.n_band + band) * 3;
        o[0] = cnt;
        o[1] = sum;
        o[2] = sq;
    }
}

cudaError_t launch_band_stats(const BankArgs& a, int order, cudaStream_t st) {
    if (a.n_sig < 1 || a.n_band < 1 || a.L < 1) return cudaErrorInvalidValue;
    // bands per CTA: as few as needed to give every SM a CTA, at most all of them
    const int sig_blocks = (a.n_sig + 31) / 32;
    int bands_per_cta = a.n_band < kBankWarps ? a.n_band : kBankWarps;
    while (bands_per_cta > 1 && sig_blocks * ((a.n_band + bands_per_cta - 1) / bands_per_cta) < sm_count()) --bands_per_cta;
    dim3 grid(sig_blocks, (a.n_band + bands_per_cta - 1) / bands_per_cta);
    const int threads = 32 * bands_per_cta;
    switch (order) {
        case 2: band_stats_kernel<3><<<grid, threads, 0, st>>>(a); break;
        case 4: band_stats_kernel<5><<<grid, threads, 0, st>>>(a); break;
        case 8: band_stats_kernel<9><<<grid, threads, 0, st>>>(a); break;     // order-4 Butterworth band-pass (fw_snr)
        case 16: band_stats_kernel<17><<<grid, threads, 0, st>>>(a); break;   // order-8 band-pass (the helper's default)
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace disco
