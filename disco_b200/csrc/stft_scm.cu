// Fused STFT + mask-weighted spatial-covariance (SCM) accumulation, and the plain STFT.
//
// Replaces, for a whole batch in one launch:
//   lb.core.stft(x, n_fft, hop, center=True)            reference tango.py:335-337
//   s_hat = m * Y, n_hat = (1 - m) * Y                   reference tango.py:347-348
//   np.outer(.., conj(..)) per (f, t) + np.mean over t   reference tango.py:357-364
//
// Work decomposition: one CTA owns one group (= one array node of one utterance, C
// microphones) and one chunk of consecutive frames.  It walks its chunk in tiles of
// 16 / ceil(C/2) frames:
//   1. the (TT+1)*hop samples of every channel are staged in shared memory by one 1-D
//      bulk TMA copy per channel (edge tiles: scalar loads with librosa's reflect padding);
//   2. FFT phase: two real channels are transformed by ONE complex FFT.  A warp computes
//      32/RA transforms at a time (RA = N/32): an RA-point in-register DFT per lane,
//      a swizzled transposition through shared memory, then one 32-point in-register
//      DFT per lane.  Spectra of the channel pairs stay in shared memory;
//   3. SCM phase: thread f owns frequency bin f for the CTA's lifetime.  It un-mixes the
//      two-for-one spectra, writes Y (frame-major rows, coalesced), and accumulates
//      the Hermitian upper triangles of  sum_t m^2 y y^H  and  sum_t (1-m)^2 y y^H
//      in registers (the outer product is shared between the two).
// Partial sums per chunk go to a small workspace and are reduced in a fixed order by
// scm_finalize_kernel (deterministic; no atomics).
#include "common.cuh"
#include "fft_reg.cuh"
#include "kernels.h"

namespace disco {

constexpr int kItems = 16;  // (frame, channel-pair) transforms per tile

template <int N>
struct FftGeom {
    static constexpr int RA = N / 32;            // radix of the per-lane first pass
    static constexpr int NB = 32 / RA;           // transforms per warp job
    static constexpr int HALF = N / 2;           // hop (50 % overlap)
    static constexpr int F = N / 2 + 1;          // bins
    static constexpr int ROW = N + (RA == 8 ? 8 : 0);  // spectrum row pitch (complex), bank padding
    static constexpr int FFT_WARPS = N / 64;     // one job per warp per tile
    static constexpr int THREADS = N / 2 + 32;   // thread f <-> bin f, f = 0 .. N/2
};

__host__ __device__ inline int tile_frames(int C) { return kItems / ((C + 1) / 2); }

// shared memory carve-up (bytes)
template <int N>
__host__ __device__ inline size_t smem_bytes(int C) {
    using G = FftGeom<N>;
    size_t spec = (size_t)kItems * G::ROW * sizeof(float2);
    size_t samp = (size_t)C * (tile_frames(C) + 1) * G::HALF * sizeof(float);
    size_t tw = (size_t)N * sizeof(float2);
    return spec + samp + tw + 64;
}

template <int N, int C, bool SCM>
__global__ void __launch_bounds__(FftGeom<N>::THREADS)
stft_scm_kernel(StftArgs p) {
    using G = FftGeom<N>;
    constexpr int RA = G::RA, NB = G::NB, H = G::HALF, F = G::F, ROW = G::ROW;
    constexpr int P = (C + 1) / 2;          // channel pairs
    constexpr int TT = kItems / P;          // frames per tile
    constexpr int NOFF = C * (C - 1) / 2;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* spec = reinterpret_cast<float2*>(smem_raw);                 // [kItems][ROW]
    float* samp = reinterpret_cast<float*>(spec + kItems * ROW);        // [C][(TT+1)*H]
    float2* tw = reinterpret_cast<float2*>(samp + C * (TT + 1) * H);    // [RA][32]
    uint64_t* bar = reinterpret_cast<uint64_t*>(tw + N);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = blockIdx.y, chunk = blockIdx.x;
    const int c_valid = min(C, p.n_sig - grp * C);       // channels present in this group
    const int L = p.L, T = p.T;
    const float* xg = p.x + (size_t)grp * C * L;
    const int t_begin = chunk * p.frames_per_chunk;
    const int t_end = min(T, t_begin + p.frames_per_chunk);

    for (int i = tid; i < N; i += blockDim.x) tw[i] = p.twiddle[i];
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    // per-lane window values for n = lane + 32 j (pre-scaled by 1/2 for the two-for-one split)
    float win[RA];
#pragma unroll
    for (int j = 0; j < RA; ++j) win[j] = p.window[lane + 32 * j];

    // SCM accumulators (thread <-> bin)
    float ps_d[C], pn_d[C];
    float2 ps_o[NOFF > 0 ? NOFF : 1], pn_o[NOFF > 0 ? NOFF : 1];
    if (SCM) {
#pragma unroll
        for (int i = 0; i < C; ++i) ps_d[i] = pn_d[i] = 0.f;
#pragma unroll
        for (int i = 0; i < NOFF; ++i) ps_o[i] = pn_o[i] = make_float2(0.f, 0.f);
    }
    __syncthreads();

    uint32_t phase = 0;
    for (int t0 = t_begin; t0 < t_end; t0 += TT) {
        const int nfr = min(TT, t_end - t0);
        // ---------------- 1. stage samples [t0*H - H, t0*H - H + (TT+1)*H) of each channel
        const int s0 = t0 * H - H;
        const bool interior = p.use_tma && nfr == TT && s0 >= 0 && s0 + (TT + 1) * H <= L;
        if (interior) {
            if (tid == 0) {
                fence_proxy_async();
                mbar_expect_tx(bar, (uint32_t)(c_valid * (TT + 1) * H * sizeof(float)));
                for (int c = 0; c < c_valid; ++c)
                    tma_load_1d(samp + c * (TT + 1) * H, xg + (size_t)c * L + s0,
                                (uint32_t)((TT + 1) * H * sizeof(float)), bar);
            }
        } else {
            const int cnt = (nfr + 1) * H;
            for (int c = 0; c < c_valid; ++c)
                for (int i = tid; i < cnt; i += blockDim.x) {
                    int s = s0 + i;                     // librosa center=True, pad_mode='reflect'
                    if (s < 0) s = -s;
                    if (s >= L) s = 2 * (L - 1) - s;
                    float v = 0.f;
                    if (s >= 0 && s < L) v = xg[(size_t)c * L + s];
                    samp[c * (TT + 1) * H + i] = v;
                }
        }
        // mask values of this tile for my bin (issued early; consumed in phase 3)
        float mk[TT];
        if (SCM && tid < F) {
#pragma unroll
            for (int tl = 0; tl < TT; ++tl) {
                mk[tl] = 0.f;
                if (tl < nfr) {
                    const int t = t0 + tl;
                    mk[tl] = p.mask_ft ? p.mask[((size_t)grp * F + tid) * T + t]
                                       : p.mask[((size_t)grp * T + t) * F + tid];
                }
            }
        }
        if (interior) {
            mbar_wait(bar, phase);
            phase ^= 1;
        } else {
            __syncthreads();
        }

        // ---------------- 2. FFT phase: warp `warp` transforms items [warp*NB, warp*NB + NB)
        if (warp < G::FFT_WARPS) {
            float2* job = spec + (size_t)warp * NB * ROW;
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                const int it = warp * NB + q;
                const int tl = it / P, pr = it % P;
                if (tl < nfr) {
                    const int ca = 2 * pr, cb = 2 * pr + 1;
                    const float* xa = samp + ca * (TT + 1) * H + tl * H + lane;
                    const float* xb = samp + cb * (TT + 1) * H + tl * H + lane;
                    const bool has_b = cb < c_valid;
                    const bool has_a = ca < c_valid;
                    float2 v[RA];
#pragma unroll
                    for (int j = 0; j < RA; ++j) {
                        float a = has_a ? xa[32 * j] : 0.f;
                        float b = has_b ? xb[32 * j] : 0.f;
                        v[j] = make_float2(a * win[j], b * win[j]);
                    }
                    dft_reg<RA, false>(v);
#pragma unroll
                    for (int k1 = 0; k1 < RA; ++k1) {
                        float2 val = (k1 == 0) ? v[0] : cmul(v[k1], tw[k1 * 32 + lane]);
                        const int m = q * RA + k1;
                        job[m * 32 + ((lane + m) & 31)] = val;
                    }
                }
            }
            __syncwarp();
            {
                const int m = lane, qq = m / RA, k1 = m % RA;
                const int it = warp * NB + qq;
                const bool live = (it / P) < nfr;
                float2 u[32];
#pragma unroll
                for (int l = 0; l < 32; ++l) u[l] = job[m * 32 + ((l + m) & 31)];
                __syncwarp();
                if (live) {
                    dft_reg<32, false>(u);
                    float2* row = job + qq * ROW + k1;
#pragma unroll
                    for (int k2 = 0; k2 < 32; ++k2) row[RA * k2] = u[k2];
                }
            }
        }
        __syncthreads();

        // ---------------- 3. un-mix, write Y, accumulate SCMs (thread <-> bin)
        if (tid < F) {
            const int f = tid, fn = (N - f) & (N - 1);
            for (int tl = 0; tl < nfr; ++tl) {
                const int t = t0 + tl;
                float2 y[C];
#pragma unroll
                for (int pr = 0; pr < P; ++pr) {
                    const float2* row = spec + (size_t)(tl * P + pr) * ROW;
                    const float2 zf = row[f], zn = row[fn];
                    // window carries the 1/2:  A = Z[f] + conj(Z[N-f]),  B = -i (Z[f] - conj(Z[N-f]))
                    y[2 * pr] = make_float2(zf.x + zn.x, zf.y - zn.y);
                    if (2 * pr + 1 < C) y[2 * pr + 1] = make_float2(zf.y + zn.y, zn.x - zf.x);
                }
#pragma unroll
                for (int c = 0; c < C; ++c)
                    if (c < c_valid) p.Y[(((size_t)grp * C + c) * T + t) * F + f] = y[c];
                if (SCM) {
                    const float m = mk[tl];
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
                }
            }
        }
        __syncthreads();
    }

    // ---------------- partial sums of this chunk -> workspace [grp][chunk][acc][F]
    if (SCM && tid < F) {
        float* out = p.part + ((size_t)grp * p.n_chunk + chunk) * (2 * C * C) * F + tid;
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
    }
}

// Reduce chunk partials in fixed order, scale by 1/T, expand to full Hermitian matrices
// Rss, Rnn [n_grp][F][C][C] complex64 (R[i][j] = mean_t a_i conj(a_j), np.outer convention).
__global__ void scm_finalize_kernel(const float* __restrict__ part, float2* __restrict__ Rss,
                                    float2* __restrict__ Rnn, int n_grp, int n_chunk, int C, int F,
                                    float inv_T) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_grp * F) return;
    const int g = idx / F, f = idx % F;
    const int nacc = 2 * C * C, noff = C * (C - 1) / 2;
    for (int which = 0; which < 2; ++which) {
        float2* R = (which == 0 ? Rss : Rnn) + ((size_t)g * F + f) * C * C;
        const int base = which * C * C;
        auto sum = [&](int a) {
            float s = 0.f;
            for (int ch = 0; ch < n_chunk; ++ch)
                s += part[(((size_t)g * n_chunk + ch) * nacc + base + a) * F + f];
            return s * inv_T;
        };
        for (int i = 0; i < C; ++i) R[i * C + i] = make_float2(sum(i), 0.f);
        int o = 0;
        for (int i = 0; i < C; ++i)
            for (int j = i + 1; j < C; ++j) {
                float re = sum(C + 2 * o), im = sum(C + 2 * o + 1);
                R[i * C + j] = make_float2(re, im);
                R[j * C + i] = make_float2(re, -im);
                ++o;
            }
        (void)noff;
    }
}

// ------------------------------------------------------------------------------ host side
template <int N, int C, bool SCM>
static cudaError_t launch_one(const StftArgs& a, int n_grp, cudaStream_t st) {
    using G = FftGeom<N>;
    auto kern = stft_scm_kernel<N, C, SCM>;
    const size_t smem = smem_bytes<N>(C);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(a.n_chunk, n_grp);
    kern<<<grid, G::THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

template <int N, bool SCM>
static cudaError_t launch_c(const StftArgs& a, int C, int n_grp, cudaStream_t st) {
    switch (C) {
        case 1: return launch_one<N, 1, SCM>(a, n_grp, st);
        case 2: return launch_one<N, 2, SCM>(a, n_grp, st);
        case 3: return launch_one<N, 3, SCM>(a, n_grp, st);
        case 4: return launch_one<N, 4, SCM>(a, n_grp, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_stft_scm(const StftArgs& a, int n_fft, int C, int n_grp, bool scm, cudaStream_t st) {
    switch (n_fft) {
        case 256: return scm ? launch_c<256, true>(a, C, n_grp, st) : launch_c<256, false>(a, C, n_grp, st);
        case 512: return scm ? launch_c<512, true>(a, C, n_grp, st) : launch_c<512, false>(a, C, n_grp, st);
        case 1024: return scm ? launch_c<1024, true>(a, C, n_grp, st) : launch_c<1024, false>(a, C, n_grp, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_scm_finalize(const float* part, float2* Rss, float2* Rnn, int n_grp, int n_chunk, int C,
                                int F, int T, cudaStream_t st) {
    const int total = n_grp * F;
    scm_finalize_kernel<<<(total + 127) / 128, 128, 0, st>>>(part, Rss, Rnn, n_grp, n_chunk, C, F, 1.0f / (float)T);
    return cudaGetLastError();
}

int stft_tile_frames(int C) { return tile_frames(C); }

}  // namespace disco
