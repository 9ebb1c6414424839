// Fused middle pass of Tango for MULTI-NODE arrays (K > 1): step-1 filter-and-sum of every node and
// the step-2 mask-weighted SCMs of every node in ONE pass over Y.
//
// Replaces, per utterance (reference tango.py:369-376, 379-386, 431-440 with mask_for_z = 'local'):
//   z_k[f, t]  = w1_k[f]^H y_k[:, f, t],  zn_k = y_k[ref] - z_k            for every node k
//   x_k        = concat(y_k, z_j for j != k)                               ("exchange" of the z's)
//   R_ss_k[f]  = mean_t (m_k x_k)(m_k x_k)^H,  R_nn_k[f] = mean_t ((1-m_k) x_k)(...)^H
// The two-kernel route (filter_sum, then masked_scm) reads Y twice and every z K-1 more times.  Here a
// CTA owns (utterance, 32-bin block[, subset of KS nodes]) and walks time in tiles of TS frames that
// stream through a ring of shared-memory stages (cp.async; scm_core.cuh):
//   loads    all K*C microphone spectra and the KS masks of tile i+NS-1 -> stage ring (asynchronous)
//   phase A  z_j of ALL K nodes for tile i+1 from shared memory -> double-buffered z tile (+ z, zn
//            written out once); items (node, frame) dealt round-robin to the warps
//   phase B  warp (node k, pair-partition p) accumulates its share of the D(D+1)/2 Hermitian
//            pairs of tile i in registers: own C spectra and the other nodes' z from shared memory
// with ONE barrier per tile (phase A runs one tile ahead of phase B).  The SCMs are written directly.
// Channels are accumulated in ROTATED node order (own mics, then z_{k+1}, z_{k+2}, ... mod K) so that all
// register indices are compile-time; the final store maps them back to the reference's channel order
// (own mics, then nodes < k, then nodes > k; concatenate_signals, tango.py:153-155).
// For K > 4 the K nodes' SCMs are split over K/KS CTAs (adjacent in launch order, so the spectra they
// share are L2 hits; each recomputes all z: C complex multiplies per value).
#include "kernels.h"
#include "scm_core.cuh"

namespace disco {

template <int C, int K, int KS, int NPART, int NS>
struct MidCfg {
    static constexpr int D = C + K - 1;
    static constexpr int TS = 4;                     // frames per tile (x32 in the Nyquist block)
    static constexpr int NPP = PairGeom<D, NPART>::NPP;
    static constexpr int NW = KS * NPART;
    static constexpr int NCH = K * C;
    static constexpr int YROWS = NCH * TS;           // rows of 32 float2 per stage
    static constexpr int MROWS = KS * TS;            // rows of 32 float per stage
    static constexpr int ZT = TS * K * 32;           // float2 per z tile
    static constexpr size_t OFF_M = (size_t)NS * YROWS * 32 * sizeof(float2);
    static constexpr size_t OFF_Z = OFF_M + (size_t)NS * MROWS * 32 * sizeof(float);
    static constexpr size_t OFF_W = OFF_Z + (size_t)2 * ZT * sizeof(float2);
    static constexpr size_t SMEM = OFF_W + (size_t)NCH * 32 * sizeof(float2);
    static constexpr int LOAD_ROUNDS = (YROWS + MROWS + NW - 1) / NW;
    static constexpr int A_ROUNDS = (K * TS + NW - 1) / NW;
};

template <int C, int K, int KS, int NPART, int NS, int PART>
DISCO_DEV void mid_phase_b(const float2* yb, const float2* zb, const float* mb, const int (&zoff)[K > 1 ? K - 1 : 1],
                           float2 (&ps)[MidCfg<C, K, KS, NPART, NS>::NPP],
                           float2 (&pn)[MidCfg<C, K, KS, NPART, NS>::NPP]) {
    using G = MidCfg<C, K, KS, NPART, NS>;
    constexpr int D = G::D, TS = G::TS;
#pragma unroll
    for (int ts = 0; ts < TS; ++ts) {
        float2 x[D];
#pragma unroll
        for (int c = 0; c < C; ++c) x[c] = yb[(c * TS + ts) * 32];
#pragma unroll
        for (int r = 0; r < K - 1; ++r) x[C + r] = zb[ts * K * 32 + zoff[r]];
        wide_point<D, NPART, PART>(x, mb[ts * 32], true, ps, pn);
    }
}

template <int C, int K, int KS, int NPART, int NS, int MINB>
__global__ void __launch_bounds__(32 * KS * NPART, MINB) tango_mid_kernel(MidArgs a) {
    using G = MidCfg<C, K, KS, NPART, NS>;
    constexpr int D = G::D, TS = G::TS, NW = G::NW;
    extern __shared__ __align__(16) unsigned char mid_smem[];
    float2* const ystage = reinterpret_cast<float2*>(mid_smem);
    float* const mstage = reinterpret_cast<float*>(mid_smem + G::OFF_M);
    float2* const zbuf = reinterpret_cast<float2*>(mid_smem + G::OFF_Z);
    float2* const w1s = reinterpret_cast<float2*>(mid_smem + G::OFF_W);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z, T = a.T, F = a.F;
    const LaneGeom lg = lane_geom(blockIdx.y, lane, F);
    const int k0 = blockIdx.x * KS;                  // first node whose SCMs this CTA accumulates
    const int kl = warp / NPART, part = warp % NPART;
    const int k = k0 + kl;
    const bool writer = (blockIdx.x == 0);           // z, zn are written once per utterance
    const int tspan = TS * lg.tmul;                  // frames covered by one tile
    const int ntile = (T + tspan - 1) / tspan;

    for (int i = warp; i < G::NCH; i += NW)          // conj(w1) of every node for this lane's bin
        w1s[i * 32 + lane] = cconj(a.W1[((size_t)(b * K + i / C) * F + lg.fcol) * C + i % C]);

    const float2* const Yb = a.Y + (size_t)b * G::NCH * T * F + lg.fcol;
    const float* const Mb = a.mask + (size_t)(b * K + k0) * T * F + lg.fcol;
    auto issue = [&](int i) {
        if (i < ntile) {
            const int st = i % NS;
            const int tbase = i * tspan + lg.tl;
            if constexpr (NW % TS == 0) {
                // row r = warp + q * NW keeps the slot ts = warp % TS for every q: one frame index and one
                // validity per thread and tile, and the channel advances by NW / TS per round
                const int t = tbase + (warp % TS) * lg.tmul;
                const bool v = lg.ok && t < T;
                const size_t step = (size_t)(NW / TS) * T * F;
                const float2* src = Yb + ((size_t)(warp / TS) * T + (v ? t : 0)) * F;
                float2* dst = ystage + (st * G::YROWS + warp) * 32 + lane;
#pragma unroll
                for (int q = 0; q < (G::YROWS + NW - 1) / NW; ++q)
                    if (G::YROWS % NW == 0 || warp + q * NW < G::YROWS) cp_async8(dst + q * NW * 32, src + q * step, v);
                const float* msrc = Mb + ((size_t)(warp / TS) * T + (v ? t : 0)) * F;
                float* mdst = mstage + (st * G::MROWS + warp) * 32 + lane;
#pragma unroll
                for (int q = 0; q < (G::MROWS + NW - 1) / NW; ++q)
                    if (warp + q * NW < G::MROWS) cp_async4(mdst + q * NW * 32, msrc + q * step, v);
            } else {
#pragma unroll
                for (int q = 0; q < G::LOAD_ROUNDS; ++q) {
                    const int r = warp + q * NW;
                    if (r < G::YROWS) {
                        const int t = tbase + (r % TS) * lg.tmul;
                        const bool v = lg.ok && t < T;
                        cp_async8(ystage + (st * G::YROWS + r) * 32 + lane,
                                  Yb + ((size_t)(r / TS) * T + (v ? t : 0)) * F, v);
                    } else if (r < G::YROWS + G::MROWS) {
                        const int rr = r - G::YROWS;
                        const int t = tbase + (rr % TS) * lg.tmul;
                        const bool v = lg.ok && t < T;
                        cp_async4(mstage + (st * G::MROWS + rr) * 32 + lane,
                                  Mb + ((size_t)(rr / TS) * T + (v ? t : 0)) * F, v);
                    }
                }
            }
        }
        cp_async_commit();   // always: keeps the group count per tile uniform
    };
    // z of every node for tile i (spectra already in stage i % NS) -> z tile i & 1, and out to HBM
    auto phase_a = [&](int i) {
        const float2* ys = ystage + (i % NS) * G::YROWS * 32 + lane;
        float2* zt = zbuf + (i & 1) * G::ZT + lane;
        const int tbase = i * tspan + lg.tl;
#pragma unroll
        for (int q = 0; q < G::A_ROUNDS; ++q) {
            const int item = warp + q * NW;
            if (item < K * TS) {
                const int j = item % K, ts = item / K;
                const float2* yj = ys + (j * C * TS + ts) * 32;
                const float2* wj = w1s + j * C * 32 + lane;
                float2 y0 = yj[0];
                float2 z = cfma(wj[0], y0, make_float2(0.f, 0.f)), yr = y0;
#pragma unroll
                for (int c = 1; c < C; ++c) {
                    const float2 yc = yj[c * TS * 32];
                    z = cfma(wj[c * 32], yc, z);
                    if (c == a.ref) yr = yc;
                }
                zt[(ts * K + j) * 32] = z;
                const int t = tbase + ts * lg.tmul;
                if (writer && lg.ok && t < T) {
                    const size_t o = ((size_t)(b * K + j) * T + t) * F + lg.fcol;
                    a.Z[o] = z;
                    if (a.ZN) a.ZN[o] = csub(yr, z);
                }
            }
        }
    };

    float2 ps[G::NPP], pn[G::NPP];
#pragma unroll
    for (int q = 0; q < G::NPP; ++q) ps[q] = pn[q] = make_float2(0.f, 0.f);
    int zoff[K > 1 ? K - 1 : 1];                     // rotated order: z_{k+1}, z_{k+2}, ...
#pragma unroll
    for (int r = 0; r < K - 1; ++r) zoff[r] = ((k + 1 + r) % K) * 32;

#pragma unroll
    for (int i = 0; i < NS - 1; ++i) issue(i);
    cp_async_wait<NS - 2>();                         // tile 0 has landed (this thread's copies)
    __syncthreads();                                 // ... everyone's, and w1s
    phase_a(0);
    for (int i = 0; i < ntile; ++i) {
        cp_async_wait<NS - 3>();                     // tile i+1 has landed
        __syncthreads();                             // z tile i visible; stage (i-1) % NS is free
        issue(i + NS - 1);
        if (i + 1 < ntile) phase_a(i + 1);
        const float2* yb = ystage + ((i % NS) * G::YROWS + k * C * TS) * 32 + lane;
        const float2* zb = zbuf + (i & 1) * G::ZT + lane;
        const float* mb = mstage + ((i % NS) * G::MROWS + kl * TS) * 32 + lane;
        switch (part) {   // warp-uniform: keeps the (i, j) of every accumulator compile-time
            case 0: mid_phase_b<C, K, KS, NPART, NS, 0>(yb, zb, mb, zoff, ps, pn); break;
            case 1: if (NPART > 1) mid_phase_b<C, K, KS, NPART, NS, (NPART > 1 ? 1 : 0)>(yb, zb, mb, zoff, ps, pn); break;
            case 2: if (NPART > 2) mid_phase_b<C, K, KS, NPART, NS, (NPART > 2 ? 2 : 0)>(yb, zb, mb, zoff, ps, pn); break;
            default: if (NPART > 3) mid_phase_b<C, K, KS, NPART, NS, (NPART > 3 ? 3 : 0)>(yb, zb, mb, zoff, ps, pn); break;
        }
    }

    if (lg.nyq) lane_butterfly<G::NPP>(ps, pn);      // CTA-uniform
    if (lg.nyq ? lane == 0 : lg.ok) {
        const size_t m = (size_t)(b * K + k) * F + lg.fcol;
        auto ref_index = [&](int r) {                // rotated channel r -> reference channel
            if (r < C) return r;
            int j = k + 1 + (r - C);
            if (j >= K) j -= K;
            return C + (j < k ? j : j - 1);
        };
        store_pairs<D, NPART>(ps, pn, part, 1.0f / (float)T, a.Rss + m * D * D, a.Rnn + m * D * D, ref_index);
    }
}

template <int C, int K, int KS, int NPART>
static cudaError_t launch_cfg(const MidArgs& a, cudaStream_t st) {
    constexpr int D = C + K - 1;
    constexpr int NS = MidCfg<C, K, KS, NPART, 4>::SMEM <= 76 * 1024 ? 4 : 3;
    using G = MidCfg<C, K, KS, NPART, NS>;
    constexpr int THREADS = 32 * G::NW;
    constexpr int BY_SMEM = (int)((227 * 1024) / (G::SMEM + 1024));
    constexpr int BY_REGS = 65536 / (THREADS * (4 * G::NPP + 2 * D + 40));   // accumulators + operands + addressing
    constexpr int MINB = BY_SMEM < BY_REGS ? (BY_SMEM < 1 ? 1 : BY_SMEM) : (BY_REGS < 1 ? 1 : BY_REGS);
    auto kern = tango_mid_kernel<C, K, KS, NPART, NS, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
    if (e != cudaSuccess) return e;
    dim3 grid(K / KS, (a.F + 31) / 32, a.B);
    kern<<<grid, THREADS, G::SMEM, st>>>(a);
    return cudaGetLastError();
}

template <int C, int K>
static cudaError_t launch_ck(const MidArgs& a, cudaStream_t st) {
    constexpr int D = C + K - 1;
    constexpr int NPAIR = D * (D + 1) / 2;
    constexpr int NPART = NPAIR <= 16 ? 1 : (NPAIR <= 32 ? 2 : (NPAIR <= 48 ? 3 : 4));
    // nodes per CTA: all of them while the CTA stays at <= 12 warps, else a divisor of K (measured at C = 2, K = 8:
    // KS 4 / NPART 3 1.81 ms; KS 2 / NPART 3 2.41; KS 2 / NPART 4 2.00; KS 4 / NPART 4 2.01)
    constexpr int KS = K * NPART <= 12 ? K : (K % 4 == 0 && 4 * NPART <= 12 ? 4 : (K % 2 == 0 ? 2 : 1));
    return launch_cfg<C, K, KS, NPART>(a, st);
}

// Supported (C, K) combinations; anything else reports cudaErrorNotSupported and the caller uses the
// two-kernel route.
cudaError_t launch_tango_mid(const MidArgs& a, cudaStream_t st) {
#define MID_CASE(c, k) \
    if (a.C == c && a.K == k) return launch_ck<c, k>(a, st);
    MID_CASE(1, 2) MID_CASE(2, 2) MID_CASE(3, 2) MID_CASE(4, 2)
    MID_CASE(1, 3) MID_CASE(2, 3) MID_CASE(3, 3) MID_CASE(4, 3)
    MID_CASE(1, 4) MID_CASE(2, 4) MID_CASE(3, 4) MID_CASE(4, 4)
    MID_CASE(2, 8) MID_CASE(4, 8) MID_CASE(2, 6)
#undef MID_CASE
    return cudaErrorNotSupported;
}

bool tango_mid_supported(int C, int K) {
    if (K == 2 || K == 3 || K == 4) return C >= 1 && C <= 4;
    if (K == 8) return C == 2 || C == 4;
    if (K == 6) return C == 2;
    return false;
}

}  // namespace disco
