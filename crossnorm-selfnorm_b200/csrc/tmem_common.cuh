// tmem_common.cuh -- tensor memory as on-chip storage: geometry of the group pipeline kernels (selfnorm_tmem.cu,
// site_tmem.cu) and the tcgen05 allocation / load / store primitives they use.
#pragma once

#include "selfnorm_fold.cuh"

namespace cnsn {
namespace flow {

constexpr int kTmT = 128;               // threads per group: one TMEM lane each
constexpr int kTmGroups = 4;            // groups per CTA
constexpr int kTmCta = kTmT * kTmGroups;
constexpr int kTmCols = 128;            // TMEM columns per group (4 x 128 = all 512 of the SM)

__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint4& v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 :: "r"(taddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 tmem_ld4(uint32_t taddr) {
    uint4 v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(taddr) : "memory");
    return v;
}
// tcgen05.ld is asynchronous: the destination registers are valid after the wait.  The registers are in/out operands of
// the wait so that the compiler cannot schedule their uses above it.
template <int K>
__device__ __forceinline__ void tmem_wait_ld(uint4 (&v)[K]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < K; ++k) asm volatile("" : "+r"(v[k].x), "+r"(v[k].y), "+r"(v[k].z), "+r"(v[k].w));
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }


// One warp of the CTA allocates all of the SM's tensor memory (may block while another kernel holds columns: the CTA
// has not taken a ticket yet, nobody waits for it); every thread returns the base address.
__device__ __forceinline__ uint32_t tmem_alloc_all(unsigned* s_tmem) {
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(s_tmem)), "r"(kTmCols * kTmGroups) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *s_tmem;
}
__device__ __forceinline__ void tmem_free_all(uint32_t tbase) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                         // every group is done with tensor memory
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(kTmCols * kTmGroups) : "memory");
}

}  // namespace flow
}  // namespace cnsn
