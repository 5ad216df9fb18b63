// dd_tma.cuh -- 1-D bulk asynchronous copies (TMA engine, cp.async.bulk) global -> shared with mbarrier
// completion, as raw PTX for sm_100a.  Source, destination and size must be multiples of 16 bytes.
#pragma once
#if defined(__CUDACC__)
__device__ __forceinline__ unsigned dd_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dd_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dd_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void dd_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dd_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dd_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dd_smem_u32(dst)), "l"(src), "r"(bytes), "r"(dd_smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: a failed try without one comes back after a few dozen cycles, and 21 waiting warps
// per SM spinning on it issue a quarter of the gallery stream's instructions; with the hint the warp sleeps in hardware.
#ifndef DD_MBAR_SUSPEND_NS
#define DD_MBAR_SUSPEND_NS 20000
#endif
__device__ __forceinline__ void dd_mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok = 0;
    while (!ok) {
#if DD_MBAR_SUSPEND_NS > 0
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(dd_smem_u32(bar)), "r"(parity), "r"((unsigned)DD_MBAR_SUSPEND_NS) : "memory");
#else
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(dd_smem_u32(bar)), "r"(parity) : "memory");
#endif
    }
}
__device__ __forceinline__ void dd_mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
#endif
