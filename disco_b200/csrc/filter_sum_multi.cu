// Final filter-and-sum of ALL K nodes of an utterance in one pass (multi-node arrays, frame-major output):
//   yf_k[f, t] = w2_k[f]^H [y_k ; z_j, j != k][:, f, t]        (reference tango.py:445-450, per node)
// The per-node kernel (filter_sum.cu) reads every z K-1 times and is launched over B*K groups; here a CTA owns
// (utterance, 32-bin block), streams tiles of the K*C microphone spectra and the K compressed signals through
// the cp.async ring of scm_core.cuh ONCE, and its warps deal the (node, frame) outputs round-robin.  HBM traffic
// = Y + Z + outputs (cfg 5: 2.6 GB instead of 6.6 GB of loads); bound: HBM.
#include "kernels.h"
#include "scm_core.cuh"

namespace disco {

template <int C, int K, int NS>
struct FsmCfg {
    static constexpr int D = C + K - 1;
    static constexpr int TS = 4, NW = 8;
    static constexpr int NCH = K * C + K;              // staged channels: all microphones, then all z
    static constexpr int ROWS = NCH * TS;
    static constexpr size_t OFF_W = (size_t)NS * ROWS * 32 * sizeof(float2);
    static constexpr size_t SMEM = OFF_W + (size_t)K * D * 32 * sizeof(float2);
    static constexpr int ITEMS = K * TS;
};

template <int C, int K, int NS>
__global__ void __launch_bounds__(256) filter_sum_multi_kernel(FilterArgs a) {
    using G = FsmCfg<C, K, NS>;
    constexpr int D = G::D, TS = G::TS, NW = G::NW;
    extern __shared__ __align__(16) unsigned char fsm_smem[];
    float2* const stage = reinterpret_cast<float2*>(fsm_smem);
    float2* const ws = reinterpret_cast<float2*>(fsm_smem + G::OFF_W);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, T = a.in.T, F = a.in.F;
    const LaneGeom lg = lane_geom(blockIdx.x, lane, F);
    const int tspan = TS * lg.tmul;
    const int ntile_all = (T + tspan - 1) / tspan;
    // the outputs of different frames are independent: gridDim.z CTAs split the tiles of a (utterance, bin block)
    const int tile0 = (int)((long long)ntile_all * blockIdx.z / gridDim.z);
    const int ntile = (int)((long long)ntile_all * (blockIdx.z + 1) / gridDim.z) - tile0;
    if (ntile <= 0) return;

    for (int i = warp; i < K * D; i += NW) {           // filters of every node for this lane's bin
        const float2 v = a.W[((size_t)(b * K + i / D) * F + lg.fcol) * D + i % D];
        ws[i * 32 + lane] = a.conj_w ? cconj(v) : v;
    }
    // NW % K == 0: a warp always serves the same node (item % K == warp % K), so its D taps live in registers
    // for the whole pass instead of being re-read from shared memory for every frame (half of the LDS traffic)
    constexpr bool WREG = (NW % K == 0);
    float2 wr[WREG ? D : 1];
    if (WREG) {
        __syncthreads();
#pragma unroll
        for (int d = 0; d < D; ++d) wr[d] = ws[((warp % K) * D + d) * 32 + lane];
    }
    const float2* const Yb = a.in.Y + (size_t)b * K * C * T * F + lg.fcol;
    const float2* const Zb = a.in.Z + (size_t)b * K * T * F + lg.fcol;
    const size_t plane = (size_t)T * F;
    const int lslot = warp % TS;                       // NW % TS == 0: a thread always loads the same slot
    auto issue = [&](int i) {
        if (i < ntile) {
            const int t = (tile0 + i) * tspan + lg.tl + lslot * lg.tmul;
            const bool v = lg.ok && t < T;
            const size_t toff = (size_t)(v ? t : 0) * F;
            float2* dst = stage + ((i % NS) * G::ROWS + warp) * 32 + lane;
#pragma unroll
            for (int q = 0; q < (G::ROWS + NW - 1) / NW; ++q) {
                const int ch = warp / TS + q * (NW / TS);
                if (G::ROWS % NW == 0 || ch < G::NCH) {
                    const float2* src = ch < K * C ? Yb + (size_t)ch * plane : Zb + (size_t)(ch - K * C) * plane;
                    cp_async8(dst + q * NW * 32, src + toff, v);
                }
            }
        }
        cp_async_commit();
    };

#pragma unroll
    for (int i = 0; i < NS - 1; ++i) issue(i);
    for (int i = 0; i < ntile; ++i) {
        cp_async_wait<NS - 2>();                       // tile i has landed (this thread's copies)
        __syncthreads();                               // ... everyone's (and ws); stage (i-1) % NS is free
        issue(i + NS - 1);
        const float2* xs = stage + (i % NS) * G::ROWS * 32 + lane;
#pragma unroll
        for (int q = 0; q < (G::ITEMS + NW - 1) / NW; ++q) {
            const int item = warp + q * NW;
            if (item < G::ITEMS) {
                const int k = item % K, ts = item / K;
                const float2* wk = ws + k * D * 32 + lane;
                const float2* yk = xs + (k * C * TS + ts) * 32;
                float2 acc = cfma(WREG ? wr[0] : wk[0], yk[0], make_float2(0.f, 0.f));
                float2 xr = yk[0];
#pragma unroll
                for (int c = 1; c < C; ++c) {
                    const float2 x = yk[c * TS * 32];
                    acc = cfma(WREG ? wr[c] : wk[c * 32], x, acc);
                    if (c == a.ref) xr = x;
                }
#pragma unroll
                for (int r = 0; r < K - 1; ++r) {      // reference order: nodes < k, then nodes > k
                    const int j = r + (r >= k ? 1 : 0);
                    const float2 x = xs[((K * C + j) * TS + ts) * 32];
                    acc = cfma(WREG ? wr[C + r] : wk[(C + r) * 32], x, acc);
                    if (C + r == a.ref) xr = x;
                }
                const int t = (tile0 + i) * tspan + lg.tl + ts * lg.tmul;
                if (lg.ok && t < T) {
                    const size_t o = ((size_t)(b * K + k) * T + t) * F + lg.fcol;
                    a.out[o] = acc;
                    if (a.resid) a.resid[o] = csub(xr, acc);
                }
            }
        }
    }
}

template <int C, int K>
static cudaError_t launch_fsm(const FilterArgs& a, cudaStream_t st) {
    constexpr int NS = 3;
    using G = FsmCfg<C, K, NS>;
    auto kern = filter_sum_multi_kernel<C, K, NS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
    if (e != cudaSuccess) return e;
    // enough CTAs for ~8 waves of the resident slots (equal-sized CTAs: few waves quantise badly), at least 8 tiles each
    const int nblk = (a.in.F + 31) / 32, B = a.in.n_grp / K;
    const int by_smem = (int)((227 * 1024) / (G::SMEM + 1024));
    const int slots = sm_count() * (by_smem < 1 ? 1 : (by_smem > 8 ? 8 : by_smem));
    int nseg = (8 * slots + nblk * B - 1) / (nblk * B);
    const int max_seg = (a.in.T / G::TS + 7) / 8;
    nseg = nseg < 1 ? 1 : (nseg > max_seg ? (max_seg < 1 ? 1 : max_seg) : nseg);
    dim3 grid(nblk, B, nseg);
    kern<<<grid, 256, G::SMEM, st>>>(a);
    return cudaGetLastError();
}

// All K nodes, frame-major output, an instantiated (C, K): otherwise cudaErrorNotSupported and the caller
// keeps the per-node kernel.
cudaError_t launch_filter_sum_multi(const FilterArgs& a, cudaStream_t st) {
    if (a.in.K < 2 || a.in.n_sel != a.in.K || a.out_ft || !a.in.Z || a.in.z_sk != 1) return cudaErrorNotSupported;
#define FSM_CASE(c, k) \
    if (a.in.C == c && a.in.K == k) return launch_fsm<c, k>(a, st);
    FSM_CASE(1, 2) FSM_CASE(2, 2) FSM_CASE(3, 2) FSM_CASE(4, 2)
    FSM_CASE(1, 3) FSM_CASE(2, 3) FSM_CASE(3, 3) FSM_CASE(4, 3)
    FSM_CASE(1, 4) FSM_CASE(2, 4) FSM_CASE(3, 4) FSM_CASE(4, 4)
    FSM_CASE(2, 8) FSM_CASE(4, 8) FSM_CASE(2, 6)
#undef FSM_CASE
    return cudaErrorNotSupported;
}

}  // namespace disco
