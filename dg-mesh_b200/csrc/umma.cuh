// umma.cuh -- thin inline-PTX layer over the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// Shared-memory operand layout used throughout (the "interleaved", no-swizzle canonical K-major
// layout of the UMMA matrix descriptor): an operand tile of R rows x K bf16 is stored as
//     unit16(row, kb) at  kb * (R * 16)  +  (row >> 3) * 128  +  (row & 7) * 16      [bytes]
// i.e. 8x8 "core matrices" of 128 contiguous bytes, all core matrices of one 8-wide k-block
// contiguous.  Descriptor: SBO (stride between 8-row groups) = 128 B, LBO (stride between the two
// k-blocks of one K=16 instruction) = R*16 B.  Field layout per cute/arch/mma_sm100_desc.hpp.
// The same bytes read as an MN-major operand (rows = the contraction index): core matrix = 8 rows
// (K) x 8 elements (M/N); LBO = stride between 8-row K groups = 128 B, SBO = stride between
// 8-element M/N groups = R*16 B (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "common.cuh"

namespace dgm {
namespace umma {

__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // [0,14)  start address
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;    // [16,30) leading byte offset
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;    // [32,46) stride byte offset
  d |= (uint64_t)1 << 46;                              // [46,48) descriptor version (Blackwell)
  // base_offset = 0, lbo_mode = 0, layout_type [61,64) = 0 (SWIZZLE_NONE / interleave)
  return d;
}

// kind::f16 instruction descriptor: D = fp32, A = B = bf16
__host__ __device__ constexpr uint32_t instr_desc_bf16(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
  return (1u << 4)                 // c_format  = F32
         | (1u << 7)               // a_format  = BF16
         | (1u << 10)              // b_format  = BF16
         | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16)  // 0 = K-major, 1 = MN-major
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // the allocating warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the tensor core (async proxy)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues for the CTA
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 consecutive accumulator columns of this warp's 32 TMEM lanes -> registers (lane = row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 16-byte async copy global -> shared with zero fill (src_bytes = 0 writes zeros)
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace umma
}  // namespace dgm
