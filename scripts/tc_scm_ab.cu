// A/B micro-benchmark: the mask-weighted covariance of a D = 8 complex channel stack (two sets: m and 1 - m)
// accumulated by (A) the packed-FP32 engine the library ships (thread <-> bin, Hermitian upper triangle, FFMA2) and
// (B) tensor cores: the real 16 x 16 Gram matrix of [re; im] per bin with mma.sync.m16n8k8 TF32, frames as the K
// dimension, 3xTF32 split for FP32-grade accuracy (and 1xTF32 for reference).
//
// Both engines read the SAME tile from shared memory (each in the layout that is conflict-free for it; the cost of
// producing the transposed layout the MMA fragments need is NOT charged to B) and loop over it REP times, so the
// figure is the pure accumulation rate of one SM: bin-frames / s, and the HBM rate that accumulation rate could keep
// up with (D x 8 B of Y + 4 B of mask per bin-frame).  Accuracy is checked against float64 on the host.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tc_scm_ab scripts/tc_scm_ab.cu && /tmp/tc_scm_ab
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <complex>

constexpr int D = 8, BINS = 256, TT = 8, THREADS = 256;
constexpr int PITCH = D * TT + 1;                 // float2 per bin in the MMA layout (odd: conflict-free fill)

__host__ __device__ inline unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__host__ __device__ inline float urand(unsigned k) { return (hash32(k) >> 8) * (1.0f / 16777216.0f); }
__host__ __device__ inline float2 sample_x(int cta, int ch, int fr, int bin) {
    const unsigned k = (((unsigned)cta * D + ch) * TT + fr) * BINS + bin;
    return make_float2(urand(2 * k + 1) - 0.5f, urand(2 * k + 2) - 0.5f);
}
__host__ __device__ inline float sample_m(int cta, int fr, int bin) {
    return urand(0x9e3779b9U + ((unsigned)cta * TT + fr) * BINS + bin);
}

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// volatile shared loads: the tile is re-read on every pass (ptxas otherwise hoists the loads AND the TF32 splits out
// of the pass loop, leaving only HMMAs in it -- a real kernel sees new frames on every k-step)
__device__ __forceinline__ float2 lds_v2(const float2* p) {
    float2 v;
    asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}
__device__ __forceinline__ float lds_v1(const float* p) {
    float v;
    asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}

// ------------------------------------------------------------------------------------------ (A) packed FP32
// out[cta][set][bin][i][j] (i <= j), complex
__global__ void __launch_bounds__(THREADS, 1) scm_ffma2(float2* __restrict__ out, int rep) {
    extern __shared__ float2 sm[];
    float2* X = sm;                                            // [D][TT][BINS]
    float* M = reinterpret_cast<float*>(sm + D * TT * BINS);   // [TT][BINS]
    const int bin = threadIdx.x, cta = blockIdx.x;
    for (int ch = 0; ch < D; ++ch)
        for (int fr = 0; fr < TT; ++fr) X[(ch * TT + fr) * BINS + bin] = sample_x(cta, ch, fr, bin);
    for (int fr = 0; fr < TT; ++fr) M[fr * BINS + bin] = sample_m(cta, fr, bin);
    __syncthreads();
    float2 dg[D];                                              // diagonal, (s, n) packed
    float2 os[D * (D - 1) / 2], on[D * (D - 1) / 2];
#pragma unroll
    for (int i = 0; i < D; ++i) dg[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < D * (D - 1) / 2; ++i) os[i] = on[i] = make_float2(0.f, 0.f);
    for (int r = 0; r < rep; ++r) {
#pragma unroll 1
        for (int fr = 0; fr < TT; ++fr) {
            const float m = lds_v1(M + fr * BINS + bin);
            const float2 mm = make_float2(m, 1.f - m);
            float2 x[D], xs[D], xn[D];
#pragma unroll
            for (int c = 0; c < D; ++c) {
                x[c] = lds_v2(X + (c * TT + fr) * BINS + bin);
                xs[c] = __fmul2_rn(x[c], make_float2(m, m));
                xn[c] = __fadd2_rn(x[c], make_float2(-xs[c].x, -xs[c].y));
                const float p = fmaf(x[c].y, x[c].y, x[c].x * x[c].x);
                dg[c] = __ffma2_rn(mm, make_float2(p, p), dg[c]);
            }
            int k = 0;
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = i + 1; j < D; ++j, ++k) {         // acc += a conj(b), a = masked x_i, b = x_j
                    float2 t = __ffma2_rn(xs[i], make_float2(x[j].x, x[j].x), os[k]);
                    os[k] = __ffma2_rn(make_float2(xs[i].y, -xs[i].x), make_float2(x[j].y, x[j].y), t);
                    t = __ffma2_rn(xn[i], make_float2(x[j].x, x[j].x), on[k]);
                    on[k] = __ffma2_rn(make_float2(xn[i].y, -xn[i].x), make_float2(x[j].y, x[j].y), t);
                }
        }
    }
    float2* o = out + (size_t)cta * 2 * BINS * D * D;
    int k = 0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        o[((0 * BINS + bin) * D + i) * D + i] = make_float2(dg[i].x, 0.f);
        o[((1 * BINS + bin) * D + i) * D + i] = make_float2(dg[i].y, 0.f);
#pragma unroll
        for (int j = i + 1; j < D; ++j, ++k) {
            o[((0 * BINS + bin) * D + i) * D + j] = os[k];
            o[((1 * BINS + bin) * D + i) * D + j] = on[k];
        }
    }
}

// ------------------------------------------------------------------------------------------ (B) tensor cores
__device__ __forceinline__ unsigned tf32_hi(float v) { unsigned r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v)); return r; }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int NSPLIT, int NB>
__global__ void __launch_bounds__(THREADS, 1) scm_mma(float2* __restrict__ out, int rep) {
    extern __shared__ float2 sm[];
    float2* X = sm;                                            // [BINS][PITCH]: (ch, frame) contiguous per bin
    float* M = reinterpret_cast<float*>(sm + BINS * PITCH);    // [BINS][TT]
    const int cta = blockIdx.x;
    {
        const int bin = threadIdx.x;
        for (int ch = 0; ch < D; ++ch)
            for (int fr = 0; fr < TT; ++fr) X[bin * PITCH + ch * TT + fr] = sample_x(cta, ch, fr, bin);
        for (int fr = 0; fr < TT; ++fr) M[bin * TT + fr] = sample_m(cta, fr, bin);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    constexpr int BPW = BINS / (THREADS / 32);                 // bins per warp
    float2* o = out + (size_t)cta * 2 * BINS * D * D;
    for (int b0 = 0; b0 < BPW; b0 += NB) {
        float acc[NB][2][2][4];                                // [bin][set][column block: re_j | im_j][fragment]
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[q][s][h][e] = 0.f;
        for (int r = 0; r < rep; ++r) {
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                const int bin = warp * BPW + b0 + q;
                const float2 x0 = lds_v2(X + bin * PITCH + g * TT + t), x1 = lds_v2(X + bin * PITCH + g * TT + t + 4);
                const float m0 = lds_v1(M + bin * TT + t), m1 = lds_v1(M + bin * TT + t + 4);
                // A = masked [re; im] (rows) x frames; B = frames x [re | im] (the same samples, unmasked)
                float av[2][4];
                av[0][0] = m0 * x0.x; av[0][1] = m0 * x0.y; av[0][2] = m1 * x1.x; av[0][3] = m1 * x1.y;
                av[1][0] = x0.x - av[0][0]; av[1][1] = x0.y - av[0][1]; av[1][2] = x1.x - av[0][2]; av[1][3] = x1.y - av[0][3];
                const float bv[2][2] = {{x0.x, x1.x}, {x0.y, x1.y}};
                unsigned bh[2][2], bl[2][2];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        bh[h][e] = tf32_hi(bv[h][e]);
                        bl[h][e] = tf32_hi(bv[h][e] - __uint_as_float(bh[h][e]));
                    }
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    unsigned ah[4], al[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        ah[e] = tf32_hi(av[s][e]);
                        al[e] = tf32_hi(av[s][e] - __uint_as_float(ah[e]));
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (NSPLIT == 3) {
                            mma_tf32(acc[q][s][h], al, bh[h][0], bh[h][1]);
                            mma_tf32(acc[q][s][h], ah, bl[h][0], bl[h][1]);
                        }
                        mma_tf32(acc[q][s][h], ah, bh[h][0], bh[h][1]);
                    }
                }
            }
        }
        // fragment (row g | g + 8, columns 2t, 2t + 1): Re R_ij = G[re_i][re_j] + G[im_i][im_j],
        //                                               Im R_ij = G[im_i][re_j] - G[re_i][im_j]
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            const int bin = warp * BPW + b0 + q;
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int i = g, j = 2 * t + e;
                    const float re = acc[q][s][0][e] + acc[q][s][1][2 + e];
                    const float im = acc[q][s][0][2 + e] - acc[q][s][1][e];
                    if (i <= j) o[((s * BINS + bin) * D + i) * D + j] = make_float2(re, i == j ? 0.f : im);
                }
        }
    }
}

// ------------------------------------------------------------------------------------------ host
static double check(const std::vector<float2>& got, int rep, int cta, int nbins_checked) {
    double worst = 0.0;
    for (int bin = 0; bin < nbins_checked; ++bin)
        for (int s = 0; s < 2; ++s) {
            double num = 0.0, den = 0.0;
            for (int i = 0; i < D; ++i)
                for (int j = i; j < D; ++j) {
                    std::complex<double> ref(0.0, 0.0);
                    for (int fr = 0; fr < TT; ++fr) {
                        const float2 a = sample_x(cta, i, fr, bin), b = sample_x(cta, j, fr, bin);
                        const double m = sample_m(cta, fr, bin), w = s ? 1.0 - m : m;
                        ref += w * std::complex<double>(a.x, a.y) * std::conj(std::complex<double>(b.x, b.y));
                    }
                    ref *= (double)rep;
                    const float2 v = got[((size_t)cta * 2 * BINS + s * BINS + bin) * D * D + i * D + j];
                    num += std::norm(std::complex<double>(v.x, v.y) - ref);
                    den += std::norm(ref);
                }
            worst = std::max(worst, std::sqrt(num / den));
        }
    return worst;
}

template <typename K>
static void run(const char* name, K kern, size_t smem, int grid, int rep, float2* d_out, std::vector<float2>& h_out, double bytes_per_bf) {
    CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) kern<<<grid, THREADS, smem>>>(d_out, rep);
    CHECK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int it = 0; it < 5; ++it) {
        CHECK(cudaEventRecord(e0));
        kern<<<grid, THREADS, smem>>>(d_out, rep);
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    CHECK(cudaMemcpy(h_out.data(), d_out, h_out.size() * sizeof(float2), cudaMemcpyDeviceToHost));
    // accuracy at a short accumulation (rep = 4: 32 frames), so that the float32 accumulation itself is not the error
    kern<<<grid, THREADS, smem>>>(d_out, 4);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaMemcpy(h_out.data(), d_out, h_out.size() * sizeof(float2), cudaMemcpyDeviceToHost));
    const double err = check(h_out, 4, 0, 16);
    const double bf = (double)grid * BINS * TT * rep;
    printf("%-28s %8.3f ms  %8.2f G bin-frames/s  keeps up with %7.0f GB/s of Y+mask  rel.err vs float64 %.2e\n",
           name, best, bf / best / 1e6, bf * bytes_per_bf / best / 1e6, err);
}

int main(int argc, char** argv) {
    int rep = argc > 1 ? atoi(argv[1]) : 400;
    cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
    const int grid = p.multiProcessorCount;
    printf("device %s, %d SMs; D = %d channels, two mask sets, tile %d bins x %d frames per SM, %d passes\n",
           p.name, grid, D, BINS, TT, rep);
    const size_t n_out = (size_t)grid * 2 * BINS * D * D;
    float2* d_out; CHECK(cudaMalloc(&d_out, n_out * sizeof(float2)));
    CHECK(cudaMemset(d_out, 0, n_out * sizeof(float2)));
    std::vector<float2> h(n_out);
    const double bpf = D * 8.0 + 4.0;
    run("FFMA2 (shipped engine)", scm_ffma2, (size_t)D * TT * BINS * 8 + TT * BINS * 4, grid, rep, d_out, h, bpf);
    const size_t smem_b = (size_t)BINS * PITCH * 8 + BINS * TT * 4;
    run("mma.sync TF32 x3, 2 bins/warp", scm_mma<3, 2>, smem_b, grid, rep, d_out, h, bpf);
    run("mma.sync TF32 x3, 4 bins/warp", scm_mma<3, 4>, smem_b, grid, rep, d_out, h, bpf);
    run("mma.sync TF32 x1, 4 bins/warp", scm_mma<1, 4>, smem_b, grid, rep, d_out, h, bpf);
    cudaFree(d_out);
    return 0;
}
