// Fused STFT + mask-weighted spatial-covariance (SCM) accumulation, and the plain STFT.
//
// Replaces, for a whole batch in one launch:
//   lb.core.stft(x, n_fft, hop, center=True)            reference tango.py:335-337
//   s_hat = m * Y, n_hat = (1 - m) * Y                   reference tango.py:347-348
//   np.outer(.., conj(..)) per (f, t) + np.mean over t   reference tango.py:357-364
//
// Persistent, warp-specialised kernel: one CTA per SM walks a contiguous range of TILES
// (tile = TT consecutive frames of one group; group = one array node of one utterance, C mics).
// Three warp roles form a two-stage producer/consumer pipeline through shared memory, linked by
// mbarriers (no __syncthreads in the steady state):
//
//   warp 0        LOADER   stages the (TT+1)*hop samples of every channel with one 1-D bulk TMA
//                          copy per channel (edge tiles: scalar loads with librosa's reflect
//                          padding), one tile ahead.  It also owns the Nyquist bin (lanes <-> frames).
//   warps 1..8    FFT      two real channels are transformed by ONE complex FFT.  A warp computes
//                          32/RA transforms per job (RA = N/32): an RA-point in-register DFT per
//                          lane, a padded transposition through shared memory, then one 32-point
//                          in-register DFT per lane.  Spectra of the channel pairs stay in smem.
//   warps 9..     SCM      thread f owns frequency bin f.  It un-mixes the two-for-one spectra,
//                          writes Y (frame-major rows, coalesced), and accumulates the Hermitian
//                          upper triangles of  sum_t m^2 y y^H  and  sum_t (1-m)^2 y y^H  in
//                          registers (one outer product feeds both).  Mask values are prefetched
//                          one tile ahead.
//
// A CTA's tile range may cross group boundaries; accumulators are flushed per (group, CTA)
// segment into a small workspace and reduced in fixed order by scm_finalize_kernel
// (deterministic, no atomics).
#include "common.cuh"
#include "fft_reg.cuh"
#include "kernels.h"

namespace disco {

template <int N>
struct FftGeom {
    static constexpr int RA = N / 32;              // radix of the per-lane first pass
    static constexpr int NB = 32 / RA;             // transforms per warp job
    static constexpr int HALF = N / 2;             // hop (50 % overlap)
    static constexpr int F = N / 2 + 1;            // bins
    static constexpr int FFT_WARPS = 8;            // one job per FFT warp per tile
    static constexpr int ITEMS = FFT_WARPS * NB;   // (frame, channel-pair) transforms per tile
    static constexpr int ROWP = 1056 / NB;         // spectrum row pitch (complex): a job = 32 x 33 scratch
    static constexpr int SCM_WARPS = N / 64;       // bins 0 .. N/2-1, one per thread
    static constexpr int WARPS = 1 + FFT_WARPS + SCM_WARPS;
    static constexpr int THREADS = 32 * WARPS;
    static constexpr int SPEC = ITEMS * ROWP;      // complex per spectrum stage
};

template <int N>
__host__ __device__ constexpr int tile_frames_n(int C) { return FftGeom<N>::ITEMS / ((C + 1) / 2); }

template <int N>
__host__ __device__ inline size_t smem_bytes(int C) {
    using G = FftGeom<N>;
    const size_t spec = 2 * (size_t)G::SPEC * sizeof(float2);
    const size_t samp = 2 * (size_t)C * (tile_frames_n<N>(C) + 1) * G::HALF * sizeof(float);
    const size_t tw = (size_t)N * sizeof(float2);
    return spec + samp + tw + 128;
}

__device__ __forceinline__ long long range_lo(long long total, int b, int nb) { return total * b / nb; }

// first CTA whose tile range contains tile i
__device__ __forceinline__ int cta_of_tile(long long i, long long total, int nb) {
    int b = (int)((i * nb) / total);
    if (b >= nb) b = nb - 1;
    while (b + 1 < nb && range_lo(total, b + 1, nb) <= i) ++b;
    while (b > 0 && range_lo(total, b, nb) > i) --b;
    return b;
}

template <int N, int C, bool SCM>
__global__ void __launch_bounds__(FftGeom<N>::THREADS, 1) stft_scm_kernel(StftArgs p) {
    using G = FftGeom<N>;
    constexpr int RA = G::RA, NB = G::NB, H = G::HALF, F = G::F, ROWP = G::ROWP;
    constexpr int P = (C + 1) / 2;            // channel pairs per frame
    constexpr int TT = G::ITEMS / P;          // frames per tile
    constexpr int NOFF = C * (C - 1) / 2;
    constexpr int SAMP = C * (TT + 1) * H;    // floats per sample stage
    constexpr int NACC = 2 * C * C;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* spec = reinterpret_cast<float2*>(smem_raw);                  // [2][ITEMS][ROWP]
    float* samp = reinterpret_cast<float*>(spec + 2 * G::SPEC);          // [2][C][(TT+1)*H]
    float2* tw = reinterpret_cast<float2*>(samp + 2 * SAMP);             // [RA][32]
    uint64_t* bars = reinterpret_cast<uint64_t*>(tw + N);
    uint64_t* samp_full = bars;        // [2]  loader -> FFT   (1 arrival + TMA bytes)
    uint64_t* samp_empty = bars + 2;   // [2]  FFT -> loader   (FFT_WARPS arrivals)
    uint64_t* spec_full = bars + 4;    // [2]  FFT -> SCM      (FFT_WARPS arrivals)
    uint64_t* spec_empty = bars + 6;   // [2]  SCM -> FFT      (SCM_WARPS + 1 arrivals)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = p.L, T = p.T;
    const int tiles_per_grp = (T + TT - 1) / TT;
    const long long total = (long long)p.n_grp * tiles_per_grp;
    const long long lo = range_lo(total, blockIdx.x, gridDim.x), hi = range_lo(total, blockIdx.x + 1, gridDim.x);
    const int n_it = (int)(hi - lo);
    if (n_it <= 0) return;

    for (int i = tid; i < N; i += blockDim.x) tw[i] = p.twiddle[i];
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&samp_full[s], 1);
            mbar_init(&samp_empty[s], G::FFT_WARPS);
            mbar_init(&spec_full[s], G::FFT_WARPS);
            mbar_init(&spec_empty[s], G::SCM_WARPS + 1);
        }
        fence_mbar_init();
    }
    __syncthreads();

    // SCM accumulators, flushed to the workspace at the end of every (group, CTA) segment
    float ps_d[C], pn_d[C];
    float2 ps_o[NOFF > 0 ? NOFF : 1], pn_o[NOFF > 0 ? NOFF : 1];
    auto acc_reset = [&]() {
#pragma unroll
        for (int i = 0; i < C; ++i) ps_d[i] = pn_d[i] = 0.f;
#pragma unroll
        for (int i = 0; i < NOFF; ++i) ps_o[i] = pn_o[i] = make_float2(0.f, 0.f);
    };
    auto acc_step = [&](const float2 (&y)[C], float m) {
        const float a = m * m, b = (1.f - m) * (1.f - m);
        int o = 0;
#pragma unroll
        for (int i = 0; i < C; ++i) {
            const float d = fmaf(y[i].x, y[i].x, y[i].y * y[i].y);
            ps_d[i] = fmaf(a, d, ps_d[i]);
            pn_d[i] = fmaf(b, d, pn_d[i]);
#pragma unroll
            for (int j = i + 1; j < C; ++j) {
                const float2 op = cmulc(y[i], y[j]);
                ps_o[o] = cfma_r(a, op, ps_o[o]);
                pn_o[o] = cfma_r(b, op, pn_o[o]);
                ++o;
            }
        }
    };
    // write this thread's accumulators for bin f into the segment's slot
    auto acc_flush = [&](int grp, int f) {
        const int slot = blockIdx.x - cta_of_tile((long long)grp * tiles_per_grp, total, gridDim.x);
        float* out = p.part + ((size_t)grp * p.slots_per_grp + slot) * NACC * F + f;
        int a = 0;
#pragma unroll
        for (int i = 0; i < C; ++i) out[(size_t)(a++) * F] = ps_d[i];
#pragma unroll
        for (int i = 0; i < NOFF; ++i) {
            out[(size_t)(a++) * F] = ps_o[i].x;
            out[(size_t)(a++) * F] = ps_o[i].y;
        }
#pragma unroll
        for (int i = 0; i < C; ++i) out[(size_t)(a++) * F] = pn_d[i];
#pragma unroll
        for (int i = 0; i < NOFF; ++i) {
            out[(size_t)(a++) * F] = pn_o[i].x;
            out[(size_t)(a++) * F] = pn_o[i].y;
        }
    };
    // un-mix the two-for-one spectra of frame tl, bin f (window carries the 1/2):
    //   A = Z[f] + conj(Z[N-f]),  B = -i (Z[f] - conj(Z[N-f]))
    auto unmix = [&](const float2* stage, int tl, int f, float2 (&y)[C]) {
        const int fn = (N - f) & (N - 1);
#pragma unroll
        for (int pr = 0; pr < P; ++pr) {
            const float2* row = stage + (size_t)(tl * P + pr) * ROWP;
            const float2 zf = row[f], zn = row[fn];
            y[2 * pr] = make_float2(zf.x + zn.x, zf.y - zn.y);
            if (2 * pr + 1 < C) y[2 * pr + 1] = make_float2(zf.y + zn.y, zn.x - zf.x);
        }
    };

    if (warp == 0) {
        // =========================================================== LOADER (+ Nyquist bin)
        auto load_tile = [&](int it) {
            const long long i = lo + it;
            const int grp = (int)(i / tiles_per_grp), t0 = (int)(i % tiles_per_grp) * TT;
            const int nfr = min(TT, T - t0), s = it & 1, c_valid = min(C, p.n_sig - grp * C);
            mbar_wait(&samp_empty[s], ((it >> 1) & 1) ^ 1);
            const float* xg = p.x + (size_t)grp * C * L;
            float* dst = samp + s * SAMP;
            const int s0 = t0 * H - H;
            const bool interior = p.use_tma && nfr == TT && s0 >= 0 && s0 + (TT + 1) * H <= L;
            if (interior) {
                if (lane == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(&samp_full[s], (uint32_t)(c_valid * (TT + 1) * H * sizeof(float)));
                    for (int c = 0; c < c_valid; ++c)
                        tma_load_1d(dst + c * (TT + 1) * H, xg + (size_t)c * L + s0,
                                    (uint32_t)((TT + 1) * H * sizeof(float)), &samp_full[s]);
                }
            } else {
                // edge tile (reflect padding / short tail): the in-range part [k_lo, k_hi) still comes by
                // TMA; the FFT warps fill the mirrored samples (at most one hop per side) themselves.
                const int cnt = (nfr + 1) * H;
                const int k_lo = max(0, -s0), k_hi = min(cnt, L - s0);
                if (lane == 0) {
                    if (p.use_tma && k_hi > k_lo) {
                        fence_proxy_async();
                        mbar_expect_tx(&samp_full[s], (uint32_t)(c_valid * (k_hi - k_lo) * sizeof(float)));
                        for (int c = 0; c < c_valid; ++c)
                            tma_load_1d(dst + c * (TT + 1) * H + k_lo, xg + (size_t)c * L + s0 + k_lo,
                                        (uint32_t)((k_hi - k_lo) * sizeof(float)), &samp_full[s]);
                    } else {
                        mbar_arrive(&samp_full[s]);
                    }
                }
            }
        };
        if (SCM) acc_reset();
        load_tile(0);
        for (int it = 0; it < n_it; ++it) {
            if (it + 1 < n_it) load_tile(it + 1);
            const long long i = lo + it;
            const int grp = (int)(i / tiles_per_grp), t0 = (int)(i % tiles_per_grp) * TT;
            const int nfr = min(TT, T - t0), s = it & 1, c_valid = min(C, p.n_sig - grp * C);
            // Nyquist bin: lane tl <-> frame t0 + tl
            float m = 0.f;
            if (SCM && lane < nfr)
                m = p.mask_ft ? p.mask[((size_t)grp * F + (F - 1)) * T + t0 + lane]
                              : p.mask[((size_t)grp * T + t0 + lane) * F + (F - 1)];
            mbar_wait(&spec_full[s], (it >> 1) & 1);
            if (lane < nfr) {
                float2 y[C];
                unmix(spec + s * G::SPEC, lane, N / 2, y);
#pragma unroll
                for (int c = 0; c < C; ++c)
                    if (c < c_valid) p.Y[(((size_t)grp * C + c) * T + t0 + lane) * F + (F - 1)] = y[c];
                if (SCM) acc_step(y, m);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&spec_empty[s]);
            const bool seg_end = (it + 1 == n_it) || ((i + 1) % tiles_per_grp == 0);
            if (SCM && seg_end) {
                // reduce the per-frame lanes (fixed butterfly order), lane 0 writes bin N/2
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
                    for (int q = 0; q < C; ++q) {
                        ps_d[q] += __shfl_xor_sync(0xffffffffu, ps_d[q], off);
                        pn_d[q] += __shfl_xor_sync(0xffffffffu, pn_d[q], off);
                    }
#pragma unroll
                    for (int q = 0; q < NOFF; ++q) {
                        ps_o[q].x += __shfl_xor_sync(0xffffffffu, ps_o[q].x, off);
                        ps_o[q].y += __shfl_xor_sync(0xffffffffu, ps_o[q].y, off);
                        pn_o[q].x += __shfl_xor_sync(0xffffffffu, pn_o[q].x, off);
                        pn_o[q].y += __shfl_xor_sync(0xffffffffu, pn_o[q].y, off);
                    }
                }
                if (lane == 0) acc_flush(grp, F - 1);
                acc_reset();
            }
        }
    } else if (warp <= G::FFT_WARPS) {
        // =========================================================== FFT warps
        const int w = warp - 1;
        float win[RA];   // window for n = lane + 32 j (pre-scaled by 1/2 for the two-for-one split)
#pragma unroll
        for (int j = 0; j < RA; ++j) win[j] = p.window[lane + 32 * j];
        for (int it = 0; it < n_it; ++it) {
            const long long i = lo + it;
            const int grp = (int)(i / tiles_per_grp), t0 = (int)(i % tiles_per_grp) * TT;
            const int nfr = min(TT, T - t0), s = it & 1, c_valid = min(C, p.n_sig - grp * C);
            const uint32_t ph = (it >> 1) & 1;
            const float* sm = samp + s * SAMP;
            float2* job = spec + s * G::SPEC + (size_t)w * NB * ROWP;
            mbar_wait(&samp_full[s], ph);
            {
                const int s0 = t0 * H - H;
                const bool interior = p.use_tma && nfr == TT && s0 >= 0 && s0 + (TT + 1) * H <= L;
                if (!interior) {   // CTA-uniform: cooperative scalar fill by the 8 FFT warps
                    const float* xg = p.x + (size_t)grp * C * L;
                    float* dst = samp + s * SAMP;
                    const int cnt = (nfr + 1) * H;
                    int k_lo = max(0, -s0), k_hi = min(cnt, L - s0);   // [k_lo, k_hi) arrived by TMA
                    if (!p.use_tma || k_hi <= k_lo) k_lo = k_hi = 0;
                    const int n_fill = cnt - (k_hi - k_lo);
                    named_bar_sync(1, 32 * G::FFT_WARPS);     // every FFT warp is done with this stage
                    for (int c = 0; c < c_valid; ++c)
#pragma unroll 4
                        for (int q = w * 32 + lane; q < n_fill; q += 32 * G::FFT_WARPS) {
                            const int k = q < k_lo ? q : q + (k_hi - k_lo);
                            int sidx = s0 + k;               // librosa center=True, pad_mode='reflect'
                            if (sidx < 0) sidx = -sidx;
                            if (sidx >= L) sidx = 2 * (L - 1) - sidx;
                            float v = 0.f;
                            if (sidx >= 0 && sidx < L) v = xg[(size_t)c * L + sidx];
                            dst[c * (TT + 1) * H + k] = v;
                        }
                    named_bar_sync(1, 32 * G::FFT_WARPS);
                }
            }
            mbar_wait(&spec_empty[s], ph ^ 1);                // spectrum stage s free (tile it-2 consumed)
            // inter-pass twiddles W_N^(lane k1): fetched once per job, ahead of the butterflies
            constexpr bool TWREG = (RA <= 16);
            float2 twr[TWREG ? RA : 1];
            if (TWREG) {
#pragma unroll
                for (int k1 = 1; k1 < RA; ++k1) twr[k1] = tw[k1 * 32 + lane];
            }
            const bool full = (nfr == TT && c_valid == C && (C % 2 == 0));
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                const int item = w * NB + q;
                const int tl = item / P, pr = item % P;
                const int ca = 2 * pr, cb = 2 * pr + 1;
                const float* xa = sm + ca * (TT + 1) * H + tl * H + lane;
                const float* xb = sm + cb * (TT + 1) * H + tl * H + lane;
                float2 v[RA];
                if (full || (tl < nfr && cb < c_valid)) {
#pragma unroll
                    for (int j = 0; j < RA; ++j) v[j] = make_float2(xa[32 * j] * win[j], xb[32 * j] * win[j]);
                } else if (tl < nfr && ca < c_valid) {
#pragma unroll
                    for (int j = 0; j < RA; ++j) v[j] = make_float2(xa[32 * j] * win[j], 0.f);
                } else {
#pragma unroll
                    for (int j = 0; j < RA; ++j) v[j] = make_float2(0.f, 0.f);
                }
                if (q == NB - 1) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&samp_empty[s]);   // all samples of this job are in registers
                }
                if (!(p.dbg & 4)) dft_reg<RA, false>(v);
#pragma unroll
                for (int k1 = 0; k1 < RA; ++k1) {
                    const float2 val = (k1 == 0) ? v[0] : cmul(v[k1], TWREG ? twr[k1] : tw[k1 * 32 + lane]);
                    job[(q * RA + k1) * 33 + lane] = val;      // scratch [32 rows][33]: conflict-free both ways
                }
            }
            __syncwarp();
            float2 u[32];
#pragma unroll
            for (int l = 0; l < 32; ++l) u[l] = job[lane * 33 + l];
            __syncwarp();
            if (!(p.dbg & 4)) dft_reg<32, false>(u);
            {
                float2* row = job + (lane / RA) * ROWP + (lane % RA);
#pragma unroll
                for (int k2 = 0; k2 < 32; ++k2) row[RA * k2] = u[k2];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&spec_full[s]);
        }
    } else {
        // =========================================================== SCM warps: thread <-> bin f
        const int f = (warp - 1 - G::FFT_WARPS) * 32 + lane;   // 0 .. N/2 - 1
        float mk[TT];
        auto load_mask = [&](int it) {
            const long long i = lo + it;
            const int grp = (int)(i / tiles_per_grp), t0 = (int)(i % tiles_per_grp) * TT;
            const int nfr = min(TT, T - t0);
#pragma unroll
            for (int tl = 0; tl < TT; ++tl) {
                mk[tl] = 0.f;
                if (tl < nfr)
                    mk[tl] = p.mask_ft ? p.mask[((size_t)grp * F + f) * T + t0 + tl]
                                       : p.mask[((size_t)grp * T + t0 + tl) * F + f];
            }
        };
        if (SCM) {
            acc_reset();
            load_mask(0);
        }
        for (int it = 0; it < n_it; ++it) {
            const long long i = lo + it;
            const int grp = (int)(i / tiles_per_grp), t0 = (int)(i % tiles_per_grp) * TT;
            const int nfr = min(TT, T - t0), s = it & 1, c_valid = min(C, p.n_sig - grp * C);
            float mcur[TT];
#pragma unroll
            for (int tl = 0; tl < TT; ++tl) mcur[tl] = SCM ? mk[tl] : 0.f;
            if (SCM && it + 1 < n_it) load_mask(it + 1);          // in flight while this tile is processed
            mbar_wait(&spec_full[s], (it >> 1) & 1);
            const float2* stage = spec + s * G::SPEC;
            float2* yc[C];   // per-channel output rows: frame offsets below are compile-time immediates
#pragma unroll
            for (int c = 0; c < C; ++c) yc[c] = p.Y + (((size_t)grp * C + c) * T + t0) * F + f;
            if (nfr == TT && c_valid == C) {                     // full tile: straight-line code
#pragma unroll
                for (int tl = 0; tl < TT; ++tl) {
                    float2 y[C];
                    unmix(stage, tl, f, y);
                    if (!(p.dbg & 1)) {
#pragma unroll
                        for (int c = 0; c < C; ++c) __stcs(&yc[c][tl * F], y[c]);
                    }
                    if (SCM && !(p.dbg & 2)) acc_step(y, mcur[tl]);
                }
            } else {
#pragma unroll
                for (int tl = 0; tl < TT; ++tl) {
                    if (tl < nfr) {
                        float2 y[C];
                        unmix(stage, tl, f, y);
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            if (c < c_valid) yc[c][tl * F] = y[c];
                        if (SCM) acc_step(y, mcur[tl]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&spec_empty[s]);
            const bool seg_end = (it + 1 == n_it) || ((i + 1) % tiles_per_grp == 0);
            if (SCM && seg_end) {
                acc_flush(grp, f);
                acc_reset();
            }
        }
    }
}

// Reduce the (group, CTA) segment partials in fixed slot order, scale by 1/T, expand to full Hermitian
// matrices Rss, Rnn [n_grp][F][C][C] complex64 (R[i][j] = mean_t a_i conj(a_j), np.outer convention).
// One block per group, one thread per bin: all 2 C^2 * n_slot loads of a thread are independent
// (coalesced over bins), so the kernel costs about one memory round trip.
template <int C>
__global__ void __launch_bounds__(288) scm_finalize_kernel(const float* __restrict__ part, float2* __restrict__ Rss,
                                                           float2* __restrict__ Rnn, int n_grp, int slots_per_grp,
                                                           int tiles_per_grp, int n_cta, int F, float inv_T) {
    constexpr int NACC = 2 * C * C;
    const int g = blockIdx.x;
    const long long total = (long long)n_grp * tiles_per_grp;
    const int b_first = cta_of_tile((long long)g * tiles_per_grp, total, n_cta);
    const int n_slot = cta_of_tile((long long)(g + 1) * tiles_per_grp - 1, total, n_cta) - b_first + 1;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        const float* base = part + (size_t)g * slots_per_grp * NACC * F + f;
        float acc[NACC];
#pragma unroll
        for (int a = 0; a < NACC; ++a) acc[a] = 0.f;
        for (int sl = 0; sl < n_slot; ++sl) {
#pragma unroll
            for (int a = 0; a < NACC; ++a) acc[a] += __ldg(base + ((size_t)sl * NACC + a) * F);
        }
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            float2* R = (which == 0 ? Rss : Rnn) + ((size_t)g * F + f) * C * C;
            const float* q = acc + which * C * C;
            int o = 0;
#pragma unroll
            for (int i = 0; i < C; ++i) {
                R[i * C + i] = make_float2(q[i] * inv_T, 0.f);
#pragma unroll
                for (int j = i + 1; j < C; ++j) {
                    const float re = q[C + 2 * o] * inv_T, im = q[C + 2 * o + 1] * inv_T;
                    R[i * C + j] = make_float2(re, im);
                    R[j * C + i] = make_float2(re, -im);
                    ++o;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------ host side
template <int N>
static int tiles_per_grp_n(int C, int T) {
    const int tt = tile_frames_n<N>(C);
    return (T + tt - 1) / tt;
}

int stft_tiles_per_grp(int n_fft, int C, int T) {
    switch (n_fft) {
        case 256: return tiles_per_grp_n<256>(C, T);
        case 512: return tiles_per_grp_n<512>(C, T);
        default: return tiles_per_grp_n<1024>(C, T);
    }
}

// Upper bound on the number of CTAs whose tile range intersects one group.
int stft_slots_per_grp(int n_grp, int tiles_per_grp, int n_cta) {
    const long long total = (long long)n_grp * tiles_per_grp;
    const long long min_range = total / n_cta;   // every CTA owns floor or ceil(total / n_cta) tiles
    if (min_range == 0) return tiles_per_grp + 1;
    return (int)(tiles_per_grp / min_range) + 2;
}

template <int N, int C, bool SCM>
static cudaError_t launch_one(const StftArgs& a, int n_cta, cudaStream_t st) {
    using G = FftGeom<N>;
    auto kern = stft_scm_kernel<N, C, SCM>;
    const size_t smem = smem_bytes<N>(C);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<n_cta, G::THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

template <int N, bool SCM>
static cudaError_t launch_c(const StftArgs& a, int C, int n_cta, cudaStream_t st) {
    switch (C) {
        case 1: return launch_one<N, 1, SCM>(a, n_cta, st);
        case 2: return launch_one<N, 2, SCM>(a, n_cta, st);
        case 3: return launch_one<N, 3, SCM>(a, n_cta, st);
        case 4: return launch_one<N, 4, SCM>(a, n_cta, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_stft_scm(const StftArgs& a, int n_fft, int C, int n_cta, bool scm, cudaStream_t st) {
    switch (n_fft) {
        case 256: return scm ? launch_c<256, true>(a, C, n_cta, st) : launch_c<256, false>(a, C, n_cta, st);
        case 512: return scm ? launch_c<512, true>(a, C, n_cta, st) : launch_c<512, false>(a, C, n_cta, st);
        case 1024: return scm ? launch_c<1024, true>(a, C, n_cta, st) : launch_c<1024, false>(a, C, n_cta, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_scm_finalize(const float* part, float2* Rss, float2* Rnn, int n_grp, int slots_per_grp,
                                int tiles_per_grp, int n_cta, int C, int F, int T, cudaStream_t st) {
    const float inv_T = 1.0f / (float)T;
    switch (C) {
        case 1: scm_finalize_kernel<1><<<n_grp, 288, 0, st>>>(part, Rss, Rnn, n_grp, slots_per_grp, tiles_per_grp, n_cta, F, inv_T); break;
        case 2: scm_finalize_kernel<2><<<n_grp, 288, 0, st>>>(part, Rss, Rnn, n_grp, slots_per_grp, tiles_per_grp, n_cta, F, inv_T); break;
        case 3: scm_finalize_kernel<3><<<n_grp, 288, 0, st>>>(part, Rss, Rnn, n_grp, slots_per_grp, tiles_per_grp, n_cta, F, inv_T); break;
        case 4: scm_finalize_kernel<4><<<n_grp, 288, 0, st>>>(part, Rss, Rnn, n_grp, slots_per_grp, tiles_per_grp, n_cta, F, inv_T); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace disco
