// In-register radix-2 decimation-in-time DFTs of length R in {2,4,8,16,32} on float2 arrays,
// written for Blackwell's packed FP32 pipe (FADD2 / FMUL2 / FFMA2 with per-half swap / negate
// operand modifiers): a butterfly with a general twiddle is THREE packed instructions
//     x = a + w b :  t = fma2(b, (c, c), a);  x = fma2((-b.y, b.x), (s, s), t)
//     y = a - w b :  y = fma2(a, (2, 2), -x)
// and a butterfly with a trivial twiddle (1, -i) is two packed additions.  Everything is fully
// unrolled: array indices, twiddle selection and the bit-reversal are compile-time, so the arrays
// live in registers and the permutation is a register renaming.
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

// (a, b) <- (a + w b, a - w b),  w = W_32^k (forward) or conj(W_32^k) (inverse); k is a compile-time
// constant after unrolling
template <bool INV>
DISCO_DEV void bfly_dit(float2& a, float2& b, int k) {
    if (k == 0) {
        const float2 x = __fadd2_rn(a, b);
        b = __fadd2_rn(a, make_float2(-b.x, -b.y));
        a = x;
        return;
    }
    if (k == 8) {   // w = -i (forward): w b = (b.y, -b.x);  +i (inverse): w b = (-b.y, b.x)
        const float2 wb = INV ? make_float2(-b.y, b.x) : make_float2(b.y, -b.x);
        const float2 x = __fadd2_rn(a, wb);
        b = __fadd2_rn(a, make_float2(-wb.x, -wb.y));
        a = x;
        return;
    }
    float2 w = tw32(k);
    if (INV) w.y = -w.y;
    float2 x = __ffma2_rn(b, make_float2(w.x, w.x), a);
    x = __ffma2_rn(make_float2(-b.y, b.x), make_float2(w.y, w.y), x);
    b = __ffma2_rn(a, make_float2(2.f, 2.f), make_float2(-x.x, -x.y));
    a = x;
}

// In-place DFT, natural-order input and output.  forward: X[k] = sum_n v[n] exp(-2 pi i n k / R)
template <int R, bool INV>
DISCO_DEV void dft_reg(float2 (&v)[R]) {
    float2 t[R];
#pragma unroll
    for (int i = 0; i < R; ++i) t[i] = v[bitrev<R>(i)];
#pragma unroll
    for (int span = 1; span < R; span <<= 1) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if ((i & span) == 0) {
                const int k = (i & (span - 1)) * (16 / span);   // W_{2 span}^{i mod span} as a power of W_32
                bfly_dit<INV>(t[i], t[i + span], k);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < R; ++i) v[i] = t[i];
}

}  // namespace disco
