// In-register radix-2 decimation-in-frequency DFTs of length R in {2,4,8,16,32} on float2
// arrays.  Everything is fully unrolled: array indices, twiddle selection and the final
// bit-reversal are compile-time, so the arrays live in registers and the permutation is a
// register renaming.
#pragma once
#include "common.cuh"
#include "tw32.cuh"

namespace disco {

template <int R>
DISCO_DEV constexpr int bitrev(int i) {
    int r = 0;
    for (int b = 1; b < R; b <<= 1) {
        r = (r << 1) | (i & 1);
        i >>= 1;
    }
    return r;
}

// d * W_32^k (forward) or d * conj(W_32^k) (inverse); k is a compile-time constant after unrolling
template <bool INV>
DISCO_DEV float2 mul_tw32(float2 d, int k) {
    if (k == 0) return d;
    if (k == 8) return INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);  // -i / +i
    float2 w = tw32(k);
    if (INV) w.y = -w.y;
    if (k == 4 || k == 12) {  // |re| == |im| == 1/sqrt(2): 2 mul + 2 add instead of 4 fma-class
        const float h = 0.707106781186547524f;
        float a = d.x * h, b = d.y * h;
        // (a + ib) * (sx + i sy), sx, sy in {+1,-1}
        float sx = (w.x > 0.f) ? 1.f : -1.f, sy = (w.y > 0.f) ? 1.f : -1.f;
        return make_float2(sx * a - sy * b, sy * a + sx * b);
    }
    return cmul(d, w);
}

// In-place DFT, natural-order output.  forward: X[k] = sum_n v[n] exp(-2 pi i n k / R)
template <int R, bool INV>
DISCO_DEV void dft_reg(float2 (&v)[R]) {
#pragma unroll
    for (int span = R / 2; span >= 1; span >>= 1) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if ((i & span) == 0) {
                const int j = i + span;
                float2 a = v[i], b = v[j];
                v[i] = cadd(a, b);
                float2 d = csub(a, b);
                const int k = (i & (span - 1)) * (16 / span);  // W_{2 span}^{i mod span} as a power of W_32
                v[j] = mul_tw32<INV>(d, k);
            }
        }
    }
    float2 t[R];
#pragma unroll
    for (int i = 0; i < R; ++i) t[i] = v[bitrev<R>(i)];
#pragma unroll
    for (int i = 0; i < R; ++i) v[i] = t[i];
}

}  // namespace disco
