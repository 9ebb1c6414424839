// Shared device helpers for the disco_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DISCO_DEV __device__ __forceinline__

namespace disco {

// ---------------------------------------------------------------- complex helpers (float2)
DISCO_DEV float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
DISCO_DEV float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// Complex products on the packed FP32 pipe: one FMUL2 + one FFMA2 each.  ptxas folds the half swaps
// and sign flips into operand modifiers when the modified VECTOR is the first operand and the
// broadcast scalar the second (the other order costs two extra instructions).
// Rounding: re = fma(a.x, b.x, -(a.y b.y)), im = fma(a.x, b.y, a.y b.x).
DISCO_DEV float2 cmul(float2 a, float2 b) {
    const float2 t = __fmul2_rn(make_float2(-b.y, b.x), make_float2(a.y, a.y));
    return __ffma2_rn(b, make_float2(a.x, a.x), t);
}
// a * conj(b):  re = fma(a.x, b.x, a.y b.y), im = fma(a.y, b.x, -(a.x b.y))
DISCO_DEV float2 cmulc(float2 a, float2 b) {
    const float2 t = __fmul2_rn(make_float2(a.y, -a.x), make_float2(b.y, b.y));
    return __ffma2_rn(a, make_float2(b.x, b.x), t);
}
// acc + a * b  (two FFMA2)
DISCO_DEV float2 cfma(float2 a, float2 b, float2 acc) {
    const float2 t = __ffma2_rn(b, make_float2(a.x, a.x), acc);
    return __ffma2_rn(make_float2(-b.y, b.x), make_float2(a.y, a.y), t);
}
// acc + conj(a) * b  (two FFMA2): the filter-and-sum term  conj(w) x
DISCO_DEV float2 cfma_cj(float2 a, float2 b, float2 acc) {
    const float2 t = __ffma2_rn(b, make_float2(a.x, a.x), acc);
    return __ffma2_rn(make_float2(b.y, -b.x), make_float2(a.y, a.y), t);
}
DISCO_DEV float2 cscale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
DISCO_DEV float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// acc += s * a   (s real)
DISCO_DEV float2 cfma_r(float s, float2 a, float2 acc) { return __ffma2_rn(make_float2(s, s), a, acc); }

// ---------------------------------------------------------------- shared-memory / TMA plumbing
DISCO_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

DISCO_DEV void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
DISCO_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
DISCO_DEV void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Wait for the phase with the given parity.  try_wait suspends the warp in hardware for up to the
// hinted time instead of spinning, so waiting warps do not steal issue slots from working ones.
DISCO_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(20000u)
            : "memory");
    } while (!done);
}
// 1-D bulk tensor-memory-accelerator copy global -> shared, completion on an mbarrier.
// dst, src 16-byte aligned, bytes a multiple of 16.
DISCO_DEV void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// Bulk prefetch of a contiguous global range into L2 (no shared-memory destination, no completion tracking):
// hides the DRAM latency of a later tma_load_1d of the same range.  src 16-byte aligned, bytes a multiple of 16.
DISCO_DEV void tma_prefetch_l2(const void* src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
DISCO_DEV void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
DISCO_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
DISCO_DEV void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// streaming (evict-first) 8-byte global store: outputs are written once and not re-read by this kernel
DISCO_DEV void st_stream(float2* p, float2 v) { __stcs(p, v); }
DISCO_DEV float ld_stream(const float* p) { return __ldcs(p); }
DISCO_DEV float2 ld_stream(const float2* p) { return __ldcs(p); }

}  // namespace disco
