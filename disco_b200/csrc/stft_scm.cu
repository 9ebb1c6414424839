// Fused STFT + mask-weighted spatial-covariance (SCM) accumulation, and the plain STFT.
//
// Replaces, for a whole batch in one launch:
//   lb.core.stft(x, n_fft, hop, center=True)            reference tango.py:335-337
//   s_hat = m * Y, n_hat = (1 - m) * Y                   reference tango.py:347-348, :413-414
//   np.outer(.., conj(..)) per (f, t) + np.mean over t   reference tango.py:357-364, :433-440
// for up to 8 microphones per array node and for ONE or TWO masks at once (NM = 2: the step-1 mask
// AND the step-2 mask of a single-node array, whose step-2 input is the same Y -- tango.py:431-440
// with K = 1 -- so Y never has to be read back for the second set of statistics).
//
// Persistent, warp-specialised kernel: one CTA per SM walks a contiguous range of TILES
// (tile = TT consecutive frames of one group; group = one array node of one utterance, C mics).
// Three warp roles form a two-stage producer/consumer pipeline through shared memory, linked by
// mbarriers (no __syncthreads in the steady state):
//
//   warp 0        LOADER   stages the (TT+1)*hop samples of every channel with one 1-D bulk TMA
//                          copy per channel (edge tiles: the FFT warps fill in librosa's reflect
//                          padding), one tile ahead.  It also owns the Nyquist bin: lanes <-> channel
//                          PAIRS (the Nyquist spectrum is real, an SCM entry a plain product).
//   FFT warps     two real channels are transformed by ONE complex FFT.  A job is 32/RA transforms
//                          (RA = N/32): an RA-point in-register DFT per lane, a padded transposition
//                          through shared memory, then one 32-point in-register DFT per lane, all on
//                          the packed FP32 pipe (fft_reg.cuh).  Spectra of the pairs stay in smem.
//   SCM warps     thread f owns frequency bin f.  It un-mixes the two-for-one spectra, writes Y
//                          (frame-major rows, coalesced), and accumulates the Hermitian upper
//                          triangles of  sum_t m^2 y y^H  and  sum_t (1-m)^2 y y^H  (per mask) in
//                          registers: one packed outer product (FMUL2 + FFMA2) feeds all of them.
//                          Mask values are prefetched a chunk of frames ahead.
// For 5..8 microphones the 128 accumulators of a bin need ~184 registers, so the roles are laid out
// on warpgroup boundaries and the register file is redistributed with setmaxnreg (loader 40, FFT 96,
// SCM 184) -- the same mechanism the TMA/MMA kernels of this architecture use.
//
// A CTA's tile range may cross group boundaries; accumulators are flushed per (group, CTA)
// segment into a small workspace and reduced in fixed order by scm_finalize_kernel or directly by
// the solver (deterministic, no atomics).
#include "common.cuh"
#include "fft_reg.cuh"
#include "kernels.h"

// tuning knobs (scripts/build_variants.py builds alternatives for A/B runs; the defaults are the measured best)
#ifndef DISCO_SS_PF
#define DISCO_SS_PF 0          // tiles ahead of the TMA load that the loader prefetches into L2 (0 = off)
#endif
#ifndef DISCO_SS_FW
#define DISCO_SS_FW 8          // FFT warps for up to 4 microphones (8 jobs per tile: 8 or 4)
#endif
#ifndef DISCO_SS_FG
#define DISCO_SS_FG 1          // FFT warp groups working on alternate (smaller) tiles; 2 * FG pipeline stages
#endif
#ifndef DISCO_SS_FFTHI
#define DISCO_SS_FFTHI 0       // 1: FFT warps get the highest warp indices (issue priority), SCM warps the lower ones
#endif

namespace disco {

template <int N, int C, int NM = 1>
struct StftCfg {
    static constexpr int RA = N / 32;              // radix of the per-lane first pass
    static constexpr int NB = 32 / RA;             // transforms per warp job
    static constexpr int HALF = N / 2;             // hop (50 % overlap)
    static constexpr int F = N / 2 + 1;            // bins
    static constexpr bool WIDE = C > 4;            // 128 accumulators per bin
    // FFT warp groups: group g transforms tiles it = g (mod FG); the tile shrinks with FG so that the shared-memory
    // budget is unchanged while the pipeline gets 2 * FG stages (deeper input prefetch, finer hand-over to the SCM warps)
    static constexpr int FG = (!WIDE && N == 512) ? DISCO_SS_FG : 1;
    static constexpr int NSTG = 2 * FG;            // pipeline stages (samples and spectra)
    static constexpr int JOBS = 8 / FG;            // jobs per tile
    static constexpr int ITEMS = JOBS * NB;        // (frame, channel-pair) transforms per tile
    static constexpr int P = (C + 1) / 2;          // channel pairs per frame
    static constexpr int TT = ITEMS / P;           // frames per tile
    // Roles on warpgroup boundaries + setmaxnreg: the SCM warps of wide arrays (128 accumulators) and of
    // two-mask runs (64) need more registers than an even split of the register file gives them.
    static constexpr bool REALLOC = (N == 512) || (WIDE && N == 256);
    // measured (profiles/ab_r2.md): with two masks the SCM warps carry as much work as the FFT warps and four FFT
    // warps running two jobs each per tile beat eight (143 vs 158 us at 64 x 4 mics x 10 s): fewer warps contend
    // for the issue slots and the SCM warps get 152 registers
    static constexpr int FFT_WARPS = (WIDE && N == 512) ? 4 : (WIDE ? 8 : ((NM == 2 && N == 512) ? 4 : DISCO_SS_FW));
    static constexpr int FWG = FFT_WARPS / FG;     // FFT warps per group
    static constexpr int JPW = JOBS / FWG;         // jobs per FFT warp and tile
    static constexpr int ROWP = 1056 / NB;         // spectrum row pitch (complex): a job = 32 x 33 scratch
    static constexpr int SCM_WARPS = N / 64;       // bins 0 .. N/2-1, one per thread
    static constexpr int LEAD_WARPS = REALLOC ? 4 : 1;   // warp 0 = loader; 1..3 idle (warpgroup padding)
    static constexpr int WARPS = LEAD_WARPS + FFT_WARPS + SCM_WARPS;
    static constexpr int THREADS = 32 * WARPS;
    static constexpr int SPEC = ITEMS * ROWP;      // complex per spectrum stage
    static constexpr int SAMP = C * (TT + 1) * HALF;   // floats per sample stage
    // registers per thread at launch: each of the 4 SM sub-partitions holds 16384 registers and ceil(WARPS/4) warps
    static constexpr int REG_LAUNCH = 16384 / ((WARPS + 3) / 4) / 32 / 8 * 8;
    static constexpr int REG_LEAD = 40, REG_FFT = 96, REG_SCM = WIDE ? 184 : (FFT_WARPS == 4 ? 152 : 120);   // REALLOC only
    static constexpr int REG_SUM = 128 * REG_LEAD + 32 * FFT_WARPS * REG_FFT + 32 * SCM_WARPS * REG_SCM;
    static_assert(!REALLOC || REG_SUM <= (REG_LAUNCH > 255 ? 255 : REG_LAUNCH) * THREADS,
                  "register budgets exceed the CTA's allocation");
};

int stft_tile_frames(int n_fft, int C) {
    const int fg = (C <= 4 && n_fft == 512) ? DISCO_SS_FG : 1;
    return ((8 / fg) * (32 / (n_fft / 32))) / ((C + 1) / 2);
}

template <int N, int C>
__host__ __device__ inline size_t smem_bytes() {
    using G = StftCfg<N, C>;
    return G::NSTG * ((size_t)G::SPEC * sizeof(float2) + (size_t)G::SAMP * sizeof(float)) + (size_t)N * sizeof(float2) + 256 +
           64 * sizeof(float);
}

__device__ __forceinline__ long long range_lo(long long total, int b, int nb) { return total * b / nb; }

// first CTA whose tile range contains tile i
__host__ __device__ __forceinline__ int cta_of_tile(long long i, long long total, int nb) {
    int b = (int)((i * nb) / total);
    if (b >= nb) b = nb - 1;
    while (b + 1 < nb && total * (b + 1) / nb <= i) ++b;
    while (b > 0 && total * b / nb > i) --b;
    return b;
}

// Warpgroup register reallocation: grow (inc) or shrink (dec) relative to the launch allocation; the PTX
// rules make the wrong direction undefined behaviour (an illegal-instruction fault on sm_100a).
template <int R, int LAUNCH>
DISCO_DEV void set_maxnreg() {
    if constexpr (R > LAUNCH) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(R));
    } else if constexpr (R < LAUNCH) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(R));
    }
}

// Accumulators of one bin: per mask, C diagonal pairs (s-weighted, n-weighted) and C(C-1)/2
// complex off-diagonal sums for each of the two weights.
template <int C, int NM>
struct ScmAcc {
    static constexpr int NOFF = C * (C - 1) / 2;
    static constexpr int NO = NOFF > 0 ? NOFF : 1;
    float2 d[NM > 0 ? NM : 1][C];     // (sum m^2 |y_i|^2, sum (1-m)^2 |y_i|^2)
    float2 os[NM > 0 ? NM : 1][NO];   // sum m^2 y_i conj(y_j), i < j row-major
    float2 on[NM > 0 ? NM : 1][NO];   // sum (1-m)^2 y_i conj(y_j)
    DISCO_DEV void reset() {
#pragma unroll
        for (int q = 0; q < NM; ++q) {
#pragma unroll
            for (int i = 0; i < C; ++i) d[q][i] = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < NOFF; ++i) os[q][i] = on[q][i] = make_float2(0.f, 0.f);
        }
    }
    // one (frame, bin) point: masks m[q]
    DISCO_DEV void step(const float2 (&y)[C], const float (&m)[NM > 0 ? NM : 1]) {
        float2 ab[NM > 0 ? NM : 1];
#pragma unroll
        for (int q = 0; q < NM; ++q) {
            const float om = 1.f - m[q];
            ab[q] = make_float2(m[q] * m[q], om * om);
        }
        int o = 0;
#pragma unroll
        for (int i = 0; i < C; ++i) {
            const float dd = fmaf(y[i].x, y[i].x, y[i].y * y[i].y);
#pragma unroll
            for (int q = 0; q < NM; ++q) d[q][i] = __ffma2_rn(make_float2(dd, dd), ab[q], d[q][i]);
#pragma unroll
            for (int j = i + 1; j < C; ++j) {
                const float2 op = cmulc(y[i], y[j]);
#pragma unroll
                for (int q = 0; q < NM; ++q) {
                    os[q][o] = cfma_r(ab[q].x, op, os[q][o]);
                    on[q][o] = cfma_r(ab[q].y, op, on[q][o]);
                }
                ++o;
            }
        }
    }
    // rows of the workspace for bin f: per mask [s: C diag, NOFF x (re, im)][n: the same]
    DISCO_DEV void flush(float* out, int F) const {
        int a = 0;
#pragma unroll
        for (int q = 0; q < NM; ++q) {
#pragma unroll
            for (int i = 0; i < C; ++i) out[(size_t)(a++) * F] = d[q][i].x;
#pragma unroll
            for (int i = 0; i < NOFF; ++i) {
                out[(size_t)(a++) * F] = os[q][i].x;
                out[(size_t)(a++) * F] = os[q][i].y;
            }
#pragma unroll
            for (int i = 0; i < C; ++i) out[(size_t)(a++) * F] = d[q][i].y;
#pragma unroll
            for (int i = 0; i < NOFF; ++i) {
                out[(size_t)(a++) * F] = on[q][i].x;
                out[(size_t)(a++) * F] = on[q][i].y;
            }
        }
    }
};

template <int N, int C, int NM>
__global__ void __launch_bounds__(StftCfg<N, C, NM>::THREADS, 1) stft_scm_kernel(StftArgs p) {
    using G = StftCfg<N, C, NM>;
    constexpr int RA = G::RA, NB = G::NB, H = G::HALF, F = G::F, ROWP = G::ROWP, P = G::P, TT = G::TT;
    constexpr int SAMP = G::SAMP;
    constexpr int NACC = NM * 2 * C * C;
    constexpr int NMX = NM > 0 ? NM : 1;
    constexpr bool SCM = NM > 0;
    constexpr int MCAP = (G::REALLOC && G::REG_SCM >= 152 && !G::WIDE) ? 16 : 8;     // mask values in flight per thread
    constexpr int MC = (TT * NM <= MCAP) ? TT : ((MCAP / NMX) < TT ? (MCAP / NMX) : TT);   // frames per mask chunk
    constexpr int NCH = (TT + MC - 1) / MC;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NSTG = G::NSTG;
    float2* spec = reinterpret_cast<float2*>(smem_raw);                  // [NSTG][ITEMS][ROWP]
    float* samp = reinterpret_cast<float*>(spec + NSTG * G::SPEC);       // [NSTG][C][(TT+1)*H]
    float2* tw = reinterpret_cast<float2*>(samp + NSTG * SAMP);          // [RA][32]
    uint64_t* bars = reinterpret_cast<uint64_t*>(tw + N);
    uint64_t* samp_full = bars;               // [NSTG]  loader -> FFT   (1 arrival + TMA bytes)
    uint64_t* samp_empty = bars + NSTG;       // [NSTG]  FFT -> loader   (FWG arrivals)
    uint64_t* spec_full = bars + 2 * NSTG;    // [NSTG]  FFT -> SCM      (FWG arrivals)
    uint64_t* spec_empty = bars + 3 * NSTG;   // [NSTG]  SCM -> FFT      (SCM_WARPS + 1 arrivals)
    float* nyq = reinterpret_cast<float*>(bars + 32);   // [TT * C <= 64] Nyquist-bin values of the current tile
    static_assert(4 * NSTG <= 32, "barrier area");
    static_assert(TT * C <= 64 && TT <= 32, "Nyquist staging");

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = p.L, T = p.T;
    const int tiles_per_grp = (T + TT - 1) / TT;
    const long long total = (long long)p.n_grp * tiles_per_grp;
    const long long lo = range_lo(total, blockIdx.x, gridDim.x), hi = range_lo(total, blockIdx.x + 1, gridDim.x);
    const int n_it = (int)(hi - lo);
    if (n_it <= 0) return;

    for (int i = tid; i < N; i += blockDim.x) tw[i] = p.twiddle[i];
    if (tid == 0) {
        for (int s = 0; s < NSTG; ++s) {
            mbar_init(&samp_full[s], 1);
            mbar_init(&samp_empty[s], G::FWG);
            mbar_init(&spec_full[s], G::FWG);
            mbar_init(&spec_empty[s], G::SCM_WARPS + 1);
        }
        fence_mbar_init();
    }
    __syncthreads();

    auto tile_of = [&](int it, int& grp, int& t0) {
        const long long i = lo + it;
        grp = (int)(i / tiles_per_grp);
        t0 = (int)(i % tiles_per_grp) * TT;
    };
    auto seg_slot = [&](int grp) {
        return blockIdx.x - cta_of_tile((long long)grp * tiles_per_grp, total, gridDim.x);
    };
    auto mask_at = [&](int q, int grp, int t, int f) {
        const float* m = q == 0 ? p.mask : p.mask2;
        return p.mask_ft ? m[((size_t)grp * F + f) * T + t] : m[((size_t)grp * T + t) * F + f];
    };

    // role layout: lead warp(s) first; then FFT and SCM warps in the order DISCO_SS_FFTHI selects (the warp
    // scheduler favours higher warp indices among ready warps, and the FFT warps are the critical path)
    constexpr int FFT_WARP0 = G::LEAD_WARPS + (DISCO_SS_FFTHI ? G::SCM_WARPS : 0);
    constexpr int SCM_WARP0 = G::LEAD_WARPS + (DISCO_SS_FFTHI ? 0 : G::FFT_WARPS);
    const bool is_fft = warp >= FFT_WARP0 && warp < FFT_WARP0 + G::FFT_WARPS;
    if (warp < G::LEAD_WARPS) {
        if (G::REALLOC) set_maxnreg<G::REG_LEAD, G::REG_LAUNCH>();
        if (warp != 0) return;
        // =========================================================== LOADER (+ Nyquist bin)
        auto load_tile = [&](int it) {
            int grp, t0;
            tile_of(it, grp, t0);
            const int nfr = min(TT, T - t0), s = it % NSTG, c_valid = min(C, p.n_sig - grp * C);
            mbar_wait(&samp_empty[s], ((it / NSTG) & 1) ^ 1);
            const float* xg = p.x + (size_t)grp * C * L;
            float* dst = samp + s * SAMP;
            const int s0 = t0 * H - H;
            // the in-range part [k_lo, k_hi) of the tile's samples comes by TMA; the FFT warps fill the
            // mirrored samples of edge tiles (at most one hop per side) themselves
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
        };
        // Nyquist bin: the spectrum is real there, so an SCM entry is a plain product y_i y_j.
        // lane l <-> (frame l / C, channel l % C) for the un-mixing and the Y store,
        // lane p <-> channel pair p (and p + 32 when C = 8) for the accumulation.
        constexpr int NP = C * (C + 1) / 2;
        constexpr int NSLOT = NP > 32 ? 2 : 1;
        int pi[NSLOT], pj[NSLOT], row[NSLOT];
#pragma unroll
        for (int u = 0; u < NSLOT; ++u) {
            const int pp = min(lane + 32 * u, NP - 1);
            if (pp < C) {
                pi[u] = pj[u] = pp;
                row[u] = pp;
            } else {
                int o = pp - C, i = 0, n = C - 1;
                while (o >= n) {
                    o -= n;
                    --n;
                    ++i;
                }
                pi[u] = i;
                pj[u] = i + 1 + o;
                row[u] = C + 2 * (pp - C);
            }
        }
        float as[NMX][NSLOT], an[NMX][NSLOT];
        auto nyq_reset = [&]() {
#pragma unroll
            for (int q = 0; q < NMX; ++q)
#pragma unroll
                for (int u = 0; u < NSLOT; ++u) as[q][u] = an[q][u] = 0.f;
        };
        // L2 prefetch of the in-range samples of tile `it` (DISCO_SS_PF tiles ahead of its TMA load)
        auto prefetch_tile = [&](int it) {
            if (!p.use_tma || it >= n_it) return;
            int grp, t0;
            tile_of(it, grp, t0);
            const int nfr = min(TT, T - t0), c_valid = min(C, p.n_sig - grp * C);
            const int s0 = t0 * H - H, cnt = (nfr + 1) * H;
            const int k_lo = max(0, -s0), k_hi = min(cnt, L - s0);
            if (k_hi > k_lo && lane < c_valid)
                tma_prefetch_l2(p.x + ((size_t)grp * C + lane) * L + s0 + k_lo, (uint32_t)((k_hi - k_lo) * sizeof(float)));
        };
        nyq_reset();
        if (DISCO_SS_PF > 0) {
            for (int i = NSTG - 1; i < NSTG - 1 + DISCO_SS_PF; ++i) prefetch_tile(i);
        }
        for (int i = 0; i < NSTG - 1; ++i)
            if (i < n_it) load_tile(i);
        for (int it = 0; it < n_it; ++it) {
            if (DISCO_SS_PF > 0) prefetch_tile(it + NSTG - 1 + DISCO_SS_PF);
            if (it + NSTG - 1 < n_it) load_tile(it + NSTG - 1);
            int grp, t0;
            tile_of(it, grp, t0);
            const int nfr = min(TT, T - t0), s = it % NSTG, c_valid = min(C, p.n_sig - grp * C);
            float mq[NMX];
#pragma unroll
            for (int q = 0; q < NMX; ++q) mq[q] = (SCM && lane < nfr) ? mask_at(q, grp, t0 + lane, F - 1) : 0.f;
            mbar_wait(&spec_full[s], (it / NSTG) & 1);
#pragma unroll
            for (int r = lane; r < TT * C; r += 32) {
                const int tl_l = r / C, c_l = r % C;
                float yv = 0.f;
                if (tl_l < nfr && c_l < c_valid) {
                    const float2 z = spec[s * G::SPEC + (size_t)(tl_l * P + c_l / 2) * ROWP + N / 2];
                    yv = (c_l & 1) ? z.y + z.y : z.x + z.x;
                    p.Y[(((size_t)grp * C + c_l) * T + t0 + tl_l) * F + (F - 1)] = make_float2(yv, 0.f);
                }
                nyq[r] = yv;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&spec_empty[s]);
            if (SCM) {
#pragma unroll
                for (int tl = 0; tl < TT; ++tl) {
                    if (tl < nfr) {   // warp-uniform
                        float pr[NSLOT];
#pragma unroll
                        for (int u = 0; u < NSLOT; ++u)
                            pr[u] = nyq[tl * C + pi[u]] * nyq[tl * C + pj[u]];
#pragma unroll
                        for (int q = 0; q < NM; ++q) {
                            const float m = __shfl_sync(0xffffffffu, mq[q], tl), om = 1.f - m;
                            const float a = m * m, b = om * om;
#pragma unroll
                            for (int u = 0; u < NSLOT; ++u) {
                                as[q][u] = fmaf(a, pr[u], as[q][u]);
                                an[q][u] = fmaf(b, pr[u], an[q][u]);
                            }
                        }
                    }
                }
                const bool seg_end = (it + 1 == n_it) || ((lo + it + 1) % tiles_per_grp == 0);
                if (seg_end) {
                    float* out = p.part + ((size_t)grp * p.slots_per_grp + seg_slot(grp)) * NACC * F + (F - 1);
#pragma unroll
                    for (int u = 0; u < NSLOT; ++u) {
                        const int pp = lane + 32 * u;
                        if (pp < NP) {
#pragma unroll
                            for (int q = 0; q < NM; ++q) {
                                float* os = out + (size_t)(q * 2 * C * C) * F;
                                float* on = os + (size_t)(C * C) * F;
                                os[(size_t)row[u] * F] = as[q][u];
                                on[(size_t)row[u] * F] = an[q][u];
                                if (pp >= C) {
                                    os[(size_t)(row[u] + 1) * F] = 0.f;
                                    on[(size_t)(row[u] + 1) * F] = 0.f;
                                }
                            }
                        }
                    }
                    nyq_reset();
                }
            }
            __syncwarp();   // nyq[] is rewritten by the next tile
        }
    } else if (is_fft) {
        if (G::REALLOC) set_maxnreg<G::REG_FFT, G::REG_LAUNCH>();
        // =========================================================== FFT warps
        const int wf = warp - FFT_WARP0, fgrp = wf / G::FWG, w = wf % G::FWG;   // group, index within the group
        constexpr bool WINREG = (RA <= 16);
        float win[WINREG ? RA : 1];   // window for n = lane + 32 j (pre-scaled by 1/2 for the two-for-one split)
        if (WINREG) {
#pragma unroll
            for (int j = 0; j < RA; ++j) win[j] = p.window[lane + 32 * j];
        }
        for (int it = fgrp; it < n_it; it += G::FG) {
            int grp, t0;
            tile_of(it, grp, t0);
            const int nfr = min(TT, T - t0), s = it % NSTG, c_valid = min(C, p.n_sig - grp * C);
            const uint32_t ph = (it / NSTG) & 1;
            const float* sm = samp + s * SAMP;
            mbar_wait(&samp_full[s], ph);
            {
                const int s0 = t0 * H - H;
                const int cnt = (nfr + 1) * H;
                int k_lo = max(0, -s0), k_hi = min(cnt, L - s0);   // [k_lo, k_hi) arrived by TMA
                if (!p.use_tma || k_hi <= k_lo) k_lo = k_hi = 0;
                const int n_fill = cnt - (k_hi - k_lo);
                if (n_fill > 0) {   // CTA-uniform: cooperative scalar fill (reflect padding) by the FFT warps
                    const float* xg = p.x + (size_t)grp * C * L;
                    float* dst = samp + s * SAMP;
                    named_bar_sync(1 + fgrp, 32 * G::FWG);        // every FFT warp of the group is done with this stage
                    for (int c = 0; c < c_valid; ++c)
#pragma unroll 4
                        for (int q = w * 32 + lane; q < n_fill; q += 32 * G::FWG) {
                            const int k = q < k_lo ? q : q + (k_hi - k_lo);
                            int sidx = s0 + k;               // librosa center=True, pad_mode='reflect'
                            if (sidx < 0) sidx = -sidx;
                            if (sidx >= L) sidx = 2 * (L - 1) - sidx;
                            float v = 0.f;
                            if (sidx >= 0 && sidx < L) v = xg[(size_t)c * L + sidx];
                            dst[c * (TT + 1) * H + k] = v;
                        }
                    named_bar_sync(1 + fgrp, 32 * G::FWG);
                }
            }
            mbar_wait(&spec_empty[s], ph ^ 1);                // spectrum stage s free (tile it-2 consumed)
            const bool full = (nfr == TT && c_valid == C && (C % 2 == 0) && (G::ITEMS % P == 0));
#pragma unroll 1
            for (int jj = 0; jj < G::JPW; ++jj) {
                const int jb = w * G::JPW + jj;
                float2* job = spec + s * G::SPEC + (size_t)jb * NB * ROWP;
                // inter-pass twiddles W_N^(lane k1): fetched once per job, shared by its transforms
                constexpr bool TWREG = (RA <= 16);
                float2 twr[TWREG ? RA : 1];
                if (TWREG) {
#pragma unroll
                    for (int k1 = 1; k1 < RA; ++k1) twr[k1] = tw[k1 * 32 + lane];
                }
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    const int item = jb * NB + q;
                    const int tl = item / P, pr = item % P;
                    const int ca = 2 * pr, cb = 2 * pr + 1;
                    const float* xa = sm + ca * (TT + 1) * H + tl * H + lane;
                    const float* xb = sm + cb * (TT + 1) * H + tl * H + lane;
                    float2 v[RA];
                    if (full || (tl < nfr && cb < c_valid)) {
#pragma unroll
                        for (int j = 0; j < RA; ++j) {
                            const float wj = WINREG ? win[j] : p.window[lane + 32 * j];
                            v[j] = __fmul2_rn(make_float2(xa[32 * j], xb[32 * j]), make_float2(wj, wj));
                        }
                    } else if (tl < nfr && ca < c_valid) {
#pragma unroll
                        for (int j = 0; j < RA; ++j) {
                            const float wj = WINREG ? win[j] : p.window[lane + 32 * j];
                            v[j] = make_float2(xa[32 * j] * wj, 0.f);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < RA; ++j) v[j] = make_float2(0.f, 0.f);
                    }
                    if (jj == G::JPW - 1 && q == NB - 1) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&samp_empty[s]);   // all samples of this warp are in registers
                    }
                    dft_reg<RA, false>(v);
#pragma unroll
                    for (int k1 = 1; k1 < RA; ++k1) v[k1] = cmul(v[k1], TWREG ? twr[k1] : tw[k1 * 32 + lane]);
#pragma unroll
                    for (int k1 = 0; k1 < RA; ++k1) job[(q * RA + k1) * 33 + lane] = v[k1];   // scratch [32 rows][33]
                }
                __syncwarp();
                float2 u[32];
#pragma unroll
                for (int l = 0; l < 32; ++l) u[l] = job[lane * 33 + l];
                __syncwarp();
                dft_reg<32, false>(u);
                {
                    float2* row = job + (lane / RA) * ROWP + (lane % RA);
#pragma unroll
                    for (int k2 = 0; k2 < 32; ++k2) row[RA * k2] = u[k2];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&spec_full[s]);
        }
    } else {
        if (G::REALLOC) set_maxnreg<G::REG_SCM, G::REG_LAUNCH>();
        // =========================================================== SCM warps: thread <-> bin f
        const int f = (warp - SCM_WARP0) * 32 + lane;   // 0 .. N/2 - 1
        const int fn = (N - f) & (N - 1);
        ScmAcc<C, NM> acc;
        float mk[NMX][MC];
        // masks of chunk `ch` of tile `it` (a chunk past the CTA's last tile loads nothing): one base pointer per
        // mask, frame stride hoisted (frame-major: F floats, (F, T) layout: 1)
        const int m_st = p.mask_ft ? 1 : F;
        const size_t m_f = p.mask_ft ? (size_t)f * T : (size_t)f;
        auto load_mask = [&](int it, int ch) {
            if (it >= n_it) return;
            int grp, t0;
            tile_of(it, grp, t0);
            const size_t base = (size_t)grp * T * F + m_f + (size_t)(t0 + ch * MC) * m_st;
#pragma unroll
            for (int q = 0; q < NM; ++q) {
                const float* mb = (q == 0 ? p.mask : p.mask2) + base;
#pragma unroll
                for (int i = 0; i < MC; ++i)
                    mk[q][i] = (ch * MC + i < TT && t0 + ch * MC + i < T) ? mb[i * m_st] : 0.f;
            }
        };
        if (SCM) {
            acc.reset();
            load_mask(0, 0);
        }
        for (int it = 0; it < n_it; ++it) {
            int grp, t0;
            tile_of(it, grp, t0);
            const int nfr = min(TT, T - t0), s = it % NSTG, c_valid = min(C, p.n_sig - grp * C);
            const float2* stage = spec + s * G::SPEC;
            float2* ybase = p.Y + ((size_t)grp * C * T + t0) * F + f;
            const size_t cstride = (size_t)T * F;
            const bool full = (nfr == TT && c_valid == C);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                float mcur[NMX][MC];
#pragma unroll
                for (int q = 0; q < NMX; ++q)
#pragma unroll
                    for (int i = 0; i < MC; ++i) mcur[q][i] = SCM ? mk[q][i] : 0.f;
                if (SCM) {   // next chunk's masks: in flight while this chunk is processed
                    if (ch + 1 < NCH)
                        load_mask(it, ch + 1);
                    else
                        load_mask(it + 1, 0);
                }
                if (ch == 0) mbar_wait(&spec_full[s], (it / NSTG) & 1);
#pragma unroll
                for (int i = 0; i < MC; ++i) {
                    const int tl = ch * MC + i;
                    if (tl < TT && (full || tl < nfr)) {
                        // un-mix the two-for-one spectra (window carries the 1/2):
                        //   A = Z[f] + conj(Z[N-f]),  B = -i (Z[f] - conj(Z[N-f]))
                        float2 y[C];
#pragma unroll
                        for (int pr = 0; pr < P; ++pr) {
                            const float2* row = stage + (size_t)(tl * P + pr) * ROWP;
                            const float2 zf = row[f], zn = row[fn];
                            y[2 * pr] = __fadd2_rn(zf, make_float2(zn.x, -zn.y));
                            if (2 * pr + 1 < C) y[2 * pr + 1] = __fadd2_rn(make_float2(zf.y, -zf.x), make_float2(zn.y, zn.x));
                        }
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            if (full || c < c_valid) __stcs(ybase + c * cstride + tl * F, y[c]);
                        if (SCM) {
                            float m[NMX];
#pragma unroll
                            for (int q = 0; q < NMX; ++q) m[q] = mcur[q][i];
                            acc.step(y, m);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&spec_empty[s]);
            const bool seg_end = (it + 1 == n_it) || ((lo + it + 1) % tiles_per_grp == 0);
            if (SCM && seg_end) {
                acc.flush(p.part + ((size_t)grp * p.slots_per_grp + seg_slot(grp)) * NACC * F + f, F);
                acc.reset();
            }
        }
    }
}

// Reduce the (group, CTA) segment partials of mask set `set` in fixed slot order, scale by 1/T, expand
// to full Hermitian matrices Rss, Rnn [n_grp][F][C][C] complex64 (R[i][j] = mean_t a_i conj(a_j),
// np.outer convention).  One block per group, one thread per bin: the loads of a thread are
// independent (coalesced over bins), so the kernel costs about one memory round trip.
template <int C>
__global__ void __launch_bounds__(288) scm_finalize_kernel(const float* __restrict__ part, float2* __restrict__ Rss,
                                                           float2* __restrict__ Rnn, int n_grp, int slots_per_grp,
                                                           int tiles_per_grp, int n_cta, int F, float inv_T, int n_set,
                                                           int set) {
    constexpr int NA = 2 * C * C;
    const int g = blockIdx.x;
    const long long total = (long long)n_grp * tiles_per_grp;
    const int b_first = cta_of_tile((long long)g * tiles_per_grp, total, n_cta);
    const int n_slot = cta_of_tile((long long)(g + 1) * tiles_per_grp - 1, total, n_cta) - b_first + 1;
    const size_t slot_stride = (size_t)n_set * NA * F;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        const float* base = part + (size_t)g * slots_per_grp * slot_stride + (size_t)set * NA * F + f;
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            float acc[C * C];
#pragma unroll
            for (int a = 0; a < C * C; ++a) acc[a] = 0.f;
            for (int sl = 0; sl < n_slot; ++sl) {
#pragma unroll
                for (int a = 0; a < C * C; ++a) acc[a] += __ldg(base + sl * slot_stride + (size_t)(which * C * C + a) * F);
            }
            float2* R = (which == 0 ? Rss : Rnn) + ((size_t)g * F + f) * C * C;
            int o = 0;
#pragma unroll
            for (int i = 0; i < C; ++i) {
                R[i * C + i] = make_float2(acc[i] * inv_T, 0.f);
#pragma unroll
                for (int j = i + 1; j < C; ++j) {
                    const float re = acc[C + 2 * o] * inv_T, im = acc[C + 2 * o + 1] * inv_T;
                    R[i * C + j] = make_float2(re, im);
                    R[j * C + i] = make_float2(re, -im);
                    ++o;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------ host side
int stft_tiles_per_grp(int n_fft, int C, int T) {
    const int tt = stft_tile_frames(n_fft, C);
    return (T + tt - 1) / tt;
}

// Upper bound on the number of CTAs whose tile range intersects one group.
int stft_slots_per_grp(int n_grp, int tiles_per_grp, int n_cta) {
    const long long total = (long long)n_grp * tiles_per_grp;
    const long long min_range = total / n_cta;   // every CTA owns floor or ceil(total / n_cta) tiles
    if (min_range == 0) return tiles_per_grp + 1;
    return (int)(tiles_per_grp / min_range) + 2;
}

bool stft_scm_supported(int n_fft, int C, int n_mask) {
    if (C < 1 || C > 8 || n_mask < 0 || n_mask > 2) return false;
    if (n_fft != 256 && n_fft != 512 && n_fft != 1024) return false;
    if (C > 4 && (n_mask == 2 || n_fft == 1024)) return false;   // 128 accumulators per bin are the register limit
    if (n_mask == 2 && n_fft == 1024) return false;
    return true;
}

template <int N, int C, int NM>
static cudaError_t launch_one(const StftArgs& a, int n_cta, cudaStream_t st) {
    using G = StftCfg<N, C, NM>;
    auto kern = stft_scm_kernel<N, C, NM>;
    const size_t smem = smem_bytes<N, C>();
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (G::REALLOC) {   // setmaxnreg.inc would wait forever if the launch allocation were smaller than the budgets
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, kern);
        if (e != cudaSuccess) return e;
        if ((long long)fa.numRegs * G::THREADS < G::REG_SUM) return cudaErrorLaunchOutOfResources;
    }
    kern<<<n_cta, G::THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

template <int N, int C>
static cudaError_t launch_nm(const StftArgs& a, int nm, int n_cta, cudaStream_t st) {
    if (nm == 0) return launch_one<N, C, 0>(a, n_cta, st);
    if (nm == 1) return launch_one<N, C, 1>(a, n_cta, st);
    if constexpr (C <= 4 && N <= 512) {
        if (nm == 2) return launch_one<N, C, 2>(a, n_cta, st);
    }
    return cudaErrorInvalidValue;
}

template <int N>
static cudaError_t launch_c(const StftArgs& a, int C, int nm, int n_cta, cudaStream_t st) {
    switch (C) {
        case 1: return launch_nm<N, 1>(a, nm, n_cta, st);
        case 2: return launch_nm<N, 2>(a, nm, n_cta, st);
        case 3: return launch_nm<N, 3>(a, nm, n_cta, st);
        case 4: return launch_nm<N, 4>(a, nm, n_cta, st);
        default: break;
    }
    if constexpr (N <= 512) {
        switch (C) {
            case 5: return launch_nm<N, 5>(a, nm, n_cta, st);
            case 6: return launch_nm<N, 6>(a, nm, n_cta, st);
            case 7: return launch_nm<N, 7>(a, nm, n_cta, st);
            case 8: return launch_nm<N, 8>(a, nm, n_cta, st);
            default: break;
        }
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_stft_scm(const StftArgs& a, int n_fft, int C, int n_cta, int n_mask, cudaStream_t st) {
    if (!stft_scm_supported(n_fft, C, n_mask)) return cudaErrorInvalidValue;
    switch (n_fft) {
        case 256: return launch_c<256>(a, C, n_mask, n_cta, st);
        case 512: return launch_c<512>(a, C, n_mask, n_cta, st);
        case 1024: return launch_c<1024>(a, C, n_mask, n_cta, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_scm_finalize(const float* part, float2* Rss, float2* Rnn, int n_grp, int slots_per_grp,
                                int tiles_per_grp, int n_cta, int C, int F, int T, int n_set, int set, cudaStream_t st) {
    const float inv_T = 1.0f / (float)T;
#define DISCO_FIN(CC)                                                                                              \
    case CC:                                                                                                       \
        scm_finalize_kernel<CC><<<n_grp, 288, 0, st>>>(part, Rss, Rnn, n_grp, slots_per_grp, tiles_per_grp, n_cta, \
                                                       F, inv_T, n_set, set);                                      \
        break;
    switch (C) {
        DISCO_FIN(1)
        DISCO_FIN(2)
        DISCO_FIN(3)
        DISCO_FIN(4)
        DISCO_FIN(5)
        DISCO_FIN(6)
        DISCO_FIN(7)
        DISCO_FIN(8)
        default: return cudaErrorInvalidValue;
    }
#undef DISCO_FIN
    return cudaGetLastError();
}

}  // namespace disco
