// Inverse STFT with overlap-add and window-sum-square normalisation, batched.
//
// Replaces lb.core.istft(S, hop_length=N/2, win_length=N, center=True, length=L)
// (reference tango.py:528-539, math_utils.py:143-152; librosa <= 0.9 semantics):
//   per frame irfft -> * periodic Hann -> overlap-add -> divide by overlap-added window^2 where
//   it exceeds tiny(float32) -> drop N/2 leading samples -> crop / zero-pad to L.
//
// One CTA owns a PAIR of signals and a chunk of frames.  The two real inverse transforms are done
// by one complex inverse FFT: Z[k] = A[k] + i B[k] (k <= N/2), Z[N-k] = conj(A[k]) + i conj(B[k]),
// so Re z = a, Im z = b.  The FFT itself is the same two-pass in-register scheme as the forward
// kernel (stft_scm.cu) with conjugated twiddles.  With 50 % overlap every output hop block j
// (samples [(j-1) hop, j hop)) is  w[n] frame_j[n] + w[n+hop] frame_{j-1}[n+hop]; the second half of
// the last frame of a tile is carried in shared memory to the next tile, and a chunk recomputes
// the one frame before its first block, so there are no atomics and the result is deterministic.
#include "common.cuh"
#include "fft_reg.cuh"
#include "kernels.h"

namespace disco {

template <int N>
struct IGeom {
    static constexpr int RA = N / 32, NB = 32 / RA, H = N / 2, F = N / 2 + 1;
    static constexpr int ROW = N + (RA == 8 ? 8 : 0);
    static constexpr int FFT_WARPS = N / 64;
    static constexpr int THREADS = N / 2 + 32;
    static constexpr int ITEMS = 16;
};

template <int N>
__global__ void __launch_bounds__(IGeom<N>::THREADS) istft_kernel(IstftArgs p, int frames_per_chunk, int T_eff) {
    using G = IGeom<N>;
    constexpr int RA = G::RA, NB = G::NB, H = G::H, F = G::F, ROW = G::ROW, TT = G::ITEMS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* rows = reinterpret_cast<float2*>(smem_raw);    // [TT][ROW]
    float2* carry = rows + TT * ROW;                       // [H] second half of the previous frame (windowed)
    float2* tw = carry + H;                                // [RA][32]
    float* win = reinterpret_cast<float*>(tw + N);         // [N]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pair = blockIdx.y, chunk = blockIdx.x;
    const int sa = 2 * pair, sb = 2 * pair + 1;
    const bool has_b = sb < p.n_sig;
    const int T = p.T, L = p.L;
    const float2* Ya = p.Y + (size_t)sa * T * F;
    const float2* Yb = p.Y + (size_t)(has_b ? sb : sa) * T * F;
    float* xa = p.x + (size_t)sa * L;
    float* xb = p.x + (size_t)(has_b ? sb : sa) * L;

    const int j_begin = chunk * frames_per_chunk;                 // first hop block of this chunk
    const int j_end = min(T_eff, j_begin + frames_per_chunk);     // blocks [j_begin, j_end) (+ T_eff if last)
    const bool last_chunk = (j_end == T_eff);
    const int fs = max(j_begin - 1, 0);                           // first frame to transform

    for (int i = tid; i < N; i += blockDim.x) {
        tw[i] = p.twiddle[i];
        win[i] = p.window[i];
    }
    if (tid < H) carry[tid] = make_float2(0.f, 0.f);
    __syncthreads();
    const float inv_n = 1.0f / (float)N;
    const float tiny = 1.17549435e-38f;

    for (int t0 = fs; t0 < j_end; t0 += TT) {
        const int nfr = min(TT, j_end - t0);
        // ---- 1. gather the two half spectra into one full complex spectrum per frame
        if (tid < F) {
            const int f = tid;
            for (int tl = 0; tl < nfr; ++tl) {
                const size_t off = (size_t)(t0 + tl) * F + f;
                float2 A = Ya[off];
                float2 B = has_b ? Yb[off] : make_float2(0.f, 0.f);
                float2* row = rows + tl * ROW;
                if (f == 0 || f == N / 2) {
                    row[f] = make_float2(A.x, B.x);  // irfft ignores the imaginary part of DC / Nyquist
                } else {
                    row[f] = make_float2(A.x - B.y, A.y + B.x);
                    row[N - f] = make_float2(A.x + B.y, B.x - A.y);
                }
            }
        }
        __syncthreads();
        // ---- 2. inverse FFT, in place in the rows
        if (warp < G::FFT_WARPS) {
            float2* job = rows + (size_t)warp * NB * ROW;
            float2 v[NB][RA];
#pragma unroll
            for (int q = 0; q < NB; ++q)
#pragma unroll
                for (int j = 0; j < RA; ++j) v[q][j] = job[q * ROW + lane + 32 * j];
            __syncwarp();
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                dft_reg<RA, true>(v[q]);
#pragma unroll
                for (int k1 = 0; k1 < RA; ++k1) {
                    float2 val = (k1 == 0) ? v[q][0] : cmul(v[q][k1], cconj(tw[k1 * 32 + lane]));
                    const int m = q * RA + k1;
                    job[m * 32 + ((lane + m) & 31)] = val;
                }
            }
            __syncwarp();
            const int m = lane, qq = m / RA, k1 = m % RA;
            float2 u[32];
#pragma unroll
            for (int l = 0; l < 32; ++l) u[l] = job[m * 32 + ((l + m) & 31)];
            __syncwarp();
            dft_reg<32, true>(u);
            float2* row = job + qq * ROW + k1;
#pragma unroll
            for (int k2 = 0; k2 < 32; ++k2) row[RA * k2] = u[k2];
        }
        __syncthreads();
        // ---- 3. overlap-add: thread n <-> sample offset n inside a hop block
        if (tid < H) {
            const int n = tid;
            const float w0 = win[n], w1 = win[n + H];
            float2 prev = carry[n];
            for (int tl = 0; tl < nfr; ++tl) {
                const int j = t0 + tl;               // frame j, hop block j
                const float2 cur = rows[tl * ROW + n];
                const float2 nxt = rows[tl * ROW + n + H];
                if (j >= j_begin) {
                    float2 val = cadd(cscale(cur, w0 * inv_n), prev);
                    const float wss = w0 * w0 + (j >= 1 ? w1 * w1 : 0.f);
                    if (wss > tiny) val = cscale(val, 1.0f / wss);
                    const int s = (j - 1) * H + n;
                    if (s >= 0 && s < L) {
                        xa[s] = val.x;
                        if (has_b) xb[s] = val.y;
                    }
                }
                prev = cscale(nxt, w1 * inv_n);
            }
            carry[n] = prev;
        }
        __syncthreads();
    }
    // ---- tail: block T_eff has only the second half of the last frame; then zero-fill up to L
    if (last_chunk) {
        if (tid < H) {
            const int n = tid;
            const float w1 = win[n + H];
            float2 val = carry[n];
            const float wss = w1 * w1;
            if (wss > tiny) val = cscale(val, 1.0f / wss);
            const int s = (T_eff - 1) * H + n;
            if (s >= 0 && s < L) {
                xa[s] = val.x;
                if (has_b) xb[s] = val.y;
            }
        }
        for (int s = T_eff * H + tid; s < L; s += blockDim.x) {
            xa[s] = 0.f;
            if (has_b) xb[s] = 0.f;
        }
    }
}

template <int N>
static cudaError_t launch_n(const IstftArgs& a, cudaStream_t st) {
    using G = IGeom<N>;
    const int H = G::H;
    int T_eff = min(a.T, (a.L + N + H - 1) / H);
    if (T_eff < 1) return cudaErrorInvalidValue;
    const int pairs = (a.n_sig + 1) / 2;
    int chunks = 1;
    while (pairs * chunks < sm_count() * 2 && (T_eff + chunks - 1) / chunks > 4 * G::ITEMS) chunks *= 2;
    int fpc = ((T_eff + chunks - 1) / chunks + G::ITEMS - 1) / G::ITEMS * G::ITEMS;
    chunks = (T_eff + fpc - 1) / fpc;
    const size_t smem = (size_t)G::ITEMS * G::ROW * sizeof(float2) + H * sizeof(float2) + N * sizeof(float2) +
                        N * sizeof(float);
    auto kern = istft_kernel<N>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(chunks, pairs);
    kern<<<grid, G::THREADS, smem, st>>>(a, fpc, T_eff);
    return cudaGetLastError();
}

cudaError_t launch_istft(const IstftArgs& a, int n_fft, cudaStream_t st) {
    if (a.n_sig <= 0) return cudaSuccess;
    switch (n_fft) {
        case 256: return launch_n<256>(a, st);
        case 512: return launch_n<512>(a, st);
        case 1024: return launch_n<1024>(a, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace disco
