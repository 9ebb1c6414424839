// Fused middle pass of Tango for MULTI-NODE arrays (K > 1): step-1 filter-and-sum of every node and
// the step-2 mask-weighted SCMs of every node in ONE pass over Y.
//
// Replaces, per utterance (reference tango.py:369-376, 379-386, 431-440 with mask_for_z = 'local'):
//   z_k[f, t]  = w1_k[f]^H y_k[:, f, t],  zn_k = y_k[ref] - z_k            for every node k
//   x_k        = concat(y_k, z_j for j != k)                               ("exchange" of the z's)
//   R_ss_k[f]  = mean_t (m_k x_k)(m_k x_k)^H,  R_nn_k[f] = mean_t ((1-m_k) x_k)(...)^H
// The two-kernel route (filter_sum, then masked_scm) reads Y twice and every z K-1 more times; here a
// CTA owns (utterance, 32-bin block) and walks time in tiles of TS frames:
//   phase A  its warps compute z_j for ALL K nodes of the tile (items (node, frame) dealt round-robin)
//            into a double-buffered shared-memory tile (and write z, zn out once),
//   barrier  one __syncthreads per tile,
//   phase B  warp (node k, pair-partition p) re-reads its own C spectra (L1-resident), takes the other
//            nodes' z from shared memory and accumulates its share of the D(D+1)/2 Hermitian pairs in
//            registers over all T frames -- so the SCMs are written directly, no partial sums.
// Channels are accumulated in ROTATED node order (own mics, then z_{k+1}, z_{k+2}, ... mod K) so that all
// register indices are compile-time; the final store maps them back to the reference's channel order
// (own mics, then nodes < k, then nodes > k; concatenate_signals, tango.py:153-155).
// For K > 4 the K nodes' SCMs are split over K/KS CTAs (each recomputes all z: C cmuls per value).
#include "common.cuh"
#include "kernels.h"

namespace disco {
namespace midv1 {

template <int D, int NPART>
struct MidGeom {
    static constexpr int NPAIR = D * (D + 1) / 2;
    static constexpr int NPP = (NPAIR + NPART - 1) / NPART;
};

template <int D>
__host__ __device__ constexpr int mp_i(int p) {
    int i = 0, n = D;
    while (p >= n) {
        p -= n;
        --n;
        ++i;
    }
    return i;
}
template <int D>
__host__ __device__ constexpr int mp_j(int p) {
    int i = 0, n = D;
    while (p >= n) {
        p -= n;
        --n;
        ++i;
    }
    return i + p;
}

template <int D, int NPART, int PART, int Q>
struct MidPairAcc {
    using G = MidGeom<D, NPART>;
    static DISCO_DEV void run(const float2 (&x)[D], float wa, float wb, float2 (&ps)[G::NPP], float2 (&pn)[G::NPP]) {
        if constexpr (Q < G::NPP) {
            constexpr int pidx = Q * NPART + PART;
            if constexpr (pidx < G::NPAIR) {
                constexpr int i = mp_i<D>(pidx), j = mp_j<D>(pidx);
                const float2 op = cmulc(x[i], x[j]);
                ps[Q] = cfma_r(wa, op, ps[Q]);
                pn[Q] = cfma_r(wb, op, pn[Q]);
            }
            MidPairAcc<D, NPART, PART, Q + 1>::run(x, wa, wb, ps, pn);
        }
    }
};

constexpr int kMidTS = 4;   // frames per tile

template <int C, int K, int KS, int NPART, int PART>
DISCO_DEV void mid_phase_b(const MidArgs& a, const float2* zt, int b, int k, int f, bool active, int lane, int t0,
                           int nfr, const float (&mcur)[kMidTS], float2 (&ps)[MidGeom<C + K - 1, NPART>::NPP],
                           float2 (&pn)[MidGeom<C + K - 1, NPART>::NPP]) {
    constexpr int D = C + K - 1;
    const int T = a.T, F = a.F;
    const float2* yk = a.Y + ((size_t)(b * K + k) * C) * T * F + f;
#pragma unroll
    for (int ts = 0; ts < kMidTS; ++ts) {
        if (ts < nfr) {
            const int t = t0 + ts;
            float2 x[D];
            const float m = mcur[ts];
            if (active) {
#pragma unroll
                for (int c = 0; c < C; ++c) x[c] = yk[((size_t)c * T + t) * F];
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c) x[c] = make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < K - 1; ++i) {                 // rotated order: z_{k+1}, z_{k+2}, ...
                int j = k + 1 + i;
                if (j >= K) j -= K;
                x[C + i] = zt[(ts * K + j) * 32 + lane];
            }
            const float wa = m * m, wb = (1.f - m) * (1.f - m);
            MidPairAcc<D, NPART, PART, 0>::run(x, wa, wb, ps, pn);
        }
    }
}

template <int C, int K, int KS, int NPART>
__global__ void __launch_bounds__(32 * KS * NPART) tango_mid_kernel(MidArgs a) {
    constexpr int D = C + K - 1;
    using G = MidGeom<D, NPART>;
    constexpr int NW = KS * NPART;
    __shared__ float2 zbuf[2][kMidTS * K * 32];
    __shared__ float2 w1s[K * C * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, T = a.T, F = a.F;
    const int f = blockIdx.x * 32 + lane;
    const bool active = f < F;
    const int fc = active ? f : F - 1;
    const int k = blockIdx.z * KS + warp / NPART;     // node whose SCMs this warp accumulates
    const int part = warp % NPART;
    const bool writer = (blockIdx.z == 0);            // z, zn are written once per utterance

    for (int i = warp; i < K * C; i += NW) {          // conj(w1) of every node for this bin block
        const int j = i / C, c = i % C;
        w1s[i * 32 + lane] = cconj(a.W1[((size_t)(b * K + j) * F + fc) * C + c]);
    }
    float2 ps[G::NPP], pn[G::NPP];
#pragma unroll
    for (int q = 0; q < G::NPP; ++q) ps[q] = pn[q] = make_float2(0.f, 0.f);
    __syncthreads();

    // Software pipeline: the phase-A spectra and the masks of the NEXT tile are loaded (into registers)
    // while the current tile's pairs are accumulated, so no HBM latency sits between two barriers.
    constexpr int MAXI = (K * kMidTS + NW - 1) / NW;       // phase-A items per warp and tile
    float2 yv[MAXI][C];
    float mnext[kMidTS];
    const float* mk = a.mask + (size_t)(b * K + k) * T * F + fc;
    auto prefetch = [&](int t0) {
        const int nfr = min(kMidTS, T - t0);
#pragma unroll
        for (int q = 0; q < MAXI; ++q) {
            const int item = warp + q * NW;
            if (item < K * nfr) {
                const int j = item % K, ts = item / K;
                const float2* yj = a.Y + ((size_t)(b * K + j) * C) * T * F + fc + (size_t)(t0 + ts) * F;
#pragma unroll
                for (int c = 0; c < C; ++c) yv[q][c] = yj[(size_t)c * T * F];
            }
        }
#pragma unroll
        for (int ts = 0; ts < kMidTS; ++ts) mnext[ts] = (ts < nfr && active) ? mk[(size_t)(t0 + ts) * F] : 0.f;
    };
    prefetch(0);
    for (int t0 = 0, it = 0; t0 < T; t0 += kMidTS, ++it) {
        const int nfr = min(kMidTS, T - t0);
        float2* zt = zbuf[it & 1];
        float mcur[kMidTS];
#pragma unroll
        for (int ts = 0; ts < kMidTS; ++ts) mcur[ts] = mnext[ts];
        // ---- phase A: z of every node for the frames of this tile (operands already in registers)
#pragma unroll
        for (int q = 0; q < MAXI; ++q) {
            const int item = warp + q * NW;
            if (item < K * nfr) {
                const int j = item % K, ts = item / K, t = t0 + ts;
                float2 z = cmul(w1s[(j * C) * 32 + lane], yv[q][0]);
                float2 yr = yv[q][0];
#pragma unroll
                for (int c = 1; c < C; ++c) {
                    z = cadd(z, cmul(w1s[(j * C + c) * 32 + lane], yv[q][c]));
                    if (c == a.ref) yr = yv[q][c];
                }
                zt[(ts * K + j) * 32 + lane] = z;
                if (writer && active) {
                    const size_t o = ((size_t)(b * K + j) * T + t) * F + f;
                    a.Z[o] = z;
                    if (a.ZN) a.ZN[o] = csub(yr, z);
                }
            }
        }
        if (t0 + kMidTS < T) prefetch(t0 + kMidTS);       // in flight during the barrier and phase B
        __syncthreads();
        // ---- phase B: this warp's share of the pairs of node k
        switch (part) {
            case 0: mid_phase_b<C, K, KS, NPART, 0>(a, zt, b, k, fc, active, lane, t0, nfr, mcur, ps, pn); break;
            case 1: if (NPART > 1) mid_phase_b<C, K, KS, NPART, (NPART > 1 ? 1 : 0)>(a, zt, b, k, fc, active, lane, t0, nfr, mcur, ps, pn); break;
            case 2: if (NPART > 2) mid_phase_b<C, K, KS, NPART, (NPART > 2 ? 2 : 0)>(a, zt, b, k, fc, active, lane, t0, nfr, mcur, ps, pn); break;
            default: if (NPART > 3) mid_phase_b<C, K, KS, NPART, (NPART > 3 ? 3 : 0)>(a, zt, b, k, fc, active, lane, t0, nfr, mcur, ps, pn); break;
        }
    }
    // ---- store: rotated channel index -> reference order (own mics, nodes < k, nodes > k)
    if (active) {
        const float inv_T = 1.0f / (float)T;
        float2* Rs = a.Rss + ((size_t)(b * K + k) * F + f) * D * D;
        float2* Rn = a.Rnn + ((size_t)(b * K + k) * F + f) * D * D;
        auto ref_index = [&](int r) {          // rotated channel r -> reference channel
            if (r < C) return r;
            int j = k + 1 + (r - C);
            if (j >= K) j -= K;
            return C + (j < k ? j : j - 1);
        };
#pragma unroll
        for (int q = 0; q < G::NPP; ++q) {
            const int pidx = q * NPART + part;
            if (pidx < G::NPAIR) {
                int i = 0, n = D, pp = pidx;
                while (pp >= n) {
                    pp -= n;
                    --n;
                    ++i;
                }
                const int ri = ref_index(i), rj = ref_index(i + pp);
                float2 s = cscale(ps[q], inv_T), nn = cscale(pn[q], inv_T);
                if (ri == rj) s.y = 0.f, nn.y = 0.f;
                Rs[ri * D + rj] = s;
                Rn[ri * D + rj] = nn;
                if (ri != rj) {
                    Rs[rj * D + ri] = cconj(s);
                    Rn[rj * D + ri] = cconj(nn);
                }
            }
        }
    }
}

template <int C, int K>
static cudaError_t launch_ck(const MidArgs& a, cudaStream_t st) {
    constexpr int D = C + K - 1;
    constexpr int KS = K <= 4 ? K : (K % 4 == 0 ? 4 : (K % 3 == 0 ? 3 : (K % 2 == 0 ? 2 : 1)));
    constexpr int NPAIR = D * (D + 1) / 2;
    constexpr int NPART = NPAIR <= 16 ? 1 : (NPAIR <= 32 ? 2 : (NPAIR <= 48 ? 3 : 4));
    dim3 grid((a.F + 31) / 32, a.B, K / KS);
    tango_mid_kernel<C, K, KS, NPART><<<grid, 32 * KS * NPART, 0, st>>>(a);
    return cudaGetLastError();
}

// Supported (C, K) combinations; anything else reports cudaErrorNotSupported and the caller uses the
// two-kernel route.
cudaError_t launch(const MidArgs& a, cudaStream_t st) {
#define MID_CASE(c, k) \
    if (a.C == c && a.K == k) return launch_ck<c, k>(a, st);
    MID_CASE(1, 2) MID_CASE(2, 2) MID_CASE(3, 2) MID_CASE(4, 2)
    MID_CASE(1, 3) MID_CASE(2, 3) MID_CASE(3, 3) MID_CASE(4, 3)
    MID_CASE(1, 4) MID_CASE(2, 4) MID_CASE(3, 4) MID_CASE(4, 4)
    MID_CASE(2, 8) MID_CASE(4, 8) MID_CASE(2, 6)
#undef MID_CASE
    return cudaErrorNotSupported;
}

bool supported(int C, int K) {
    if (K == 2 || K == 3 || K == 4) return C >= 1 && C <= 4;
    if (K == 8) return C == 2 || C == 4;
    if (K == 6) return C == 2;
    return false;
}

}  // namespace midv1
// previous generation of the fused middle pass, kept for A/B timing (DISCO_MID_V1=1)
cudaError_t launch_tango_mid_v1(const MidArgs& a, cudaStream_t st) { return midv1::launch(a, st); }
}  // namespace disco
