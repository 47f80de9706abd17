// Shared device helpers for the sm_100a kernels: mbarrier / bulk-copy (TMA) PTX wrappers and the
// layout of the coefficient stream the Wigner kernels consume.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "bq_internal.h"

namespace bq {

// --------------------------------------------------------------------------------------------
// Coefficient stream (one per cutoff N, consumed strictly front to back by the Wigner kernels)
//
//   record 0                      header:   c = 2*rho[0, N-1]                     (w0 seed)
//   for L = N-2 ... 0, K = N-L:   K step records, r = 0..K-1, j = K-1-r, k = j+2
//                                     c  = f*rho[j, j+L]   (f = 2 off the main diagonal, else 1)
//                                     na = -sqrt((k-1)(L+k-1)/((L+k)k))
//                                     ne = -(L+2k-1)/sqrt((L+k)k)
//                                     s  =  1/sqrt((L+k)k)
//                                 1 tail record:  ne = -sqrt(L+1), s = 1/sqrt(L+1)
//
// The rho-dependent part (c, 16 B) and the constant part (na, ne, s, pad: 32 B) are two parallel
// arrays so that a batch of frames shares one constant table.  Follows the recurrences of
// qutip 5.2.2 `_wigner_clenshaw` / `_wig_laguerre_val` behind src/bosonic_qiskit/wigner.py:190.
// --------------------------------------------------------------------------------------------
// (stream_records / diag_pos live in bq_internal.h so the host code can size buffers)

// --------------------------------------------------------------------------------------------
// mbarrier + cp.async.bulk (1-D TMA) wrappers
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy, completion counted in bytes on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace bq
