// tcgen05 (5th-gen tensor core) primitives for sm_100a, hand-written inline PTX:
// TMEM allocation, shared-memory matrix descriptors, kind::tf32 MMA issue, commit -> mbarrier,
// TMEM -> register loads.  Bit layouts follow the PTX ISA "tcgen05" matrix/instruction
// descriptors (cross-checked against the CuTe headers shipped in the image:
// cute/arch/mma_sm100_desc.hpp, mma_sm100_umma.hpp, copy_sm100.hpp, tmem_allocator_sm100.hpp).
//
// Operand layout used throughout (both A = activations and B = weights): K-major,
// SWIZZLE_128B, 32-bit elements consumed as TF32 (validated on B200 by tests/test_gpu_umma.py).
// A tile with R rows (edges for A, output features for B) and K reduction columns is stored as
// K/32 panels; panel p holds, for every row, the 32 consecutive k in [32p, 32p+32) as one
// 128-byte line, 8-row groups of 1024 bytes, 16-byte chunks XOR-swizzled with row%8:
//     float index(row, k) = p*R*32 + row*32 + ((((k%32)/4) ^ (row%8))*4) + k%4        (p = k/32)
// SBO (between 8-row groups) = 1024 B; one tcgen05.mma.kind::tf32 consumes K = 8 (32 bytes of
// every row): the descriptor start address advances by 32 B inside a panel.  Panels must be
// 1024-byte aligned.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// float index of element (row, k) inside a K-major operand array with R rows
__host__ __device__ __forceinline__ int opk_idx(int row, int k, int R) {
  return (k >> 5) * (R * 32) + row * 32 + ((((k & 31) >> 2) ^ (row & 7)) << 2) + (k & 3);
}
// ---- TMEM ------------------------------------------------------------------------------------
// one full warp; writes the TMEM base address to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (1-D, contiguous), completion signalled on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- descriptors ------------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major SWIZZLE_128B (see header comment)
__device__ __forceinline__ uint64_t make_desc_k_sw128(const void* p) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(p) >> 4) & 0x3FFF);               // start address     bits [0,14)
  d |= (uint64_t)1 << 16;                                       // leading byte off. (unused for swizzled K-major)
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;                // stride byte off.  bits [32,46): 8 rows x 128 B
  d |= (uint64_t)1 << 46;                                       // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                                       // layout type SWIZZLE_128B
  return d;
}
// same from a shared-space byte address (lets the caller keep the address in a uniform register)
__device__ __forceinline__ uint64_t make_desc_k_sw128_addr(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((1024u >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t"
      "}" : "+r"(pred) : "r"(0xffffffffu));
  return pred != 0;
}
// instruction descriptor: D=F32, A=B=TF32, both K-major, M=128, N
__host__ __device__ constexpr uint32_t make_idesc_tf32(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; single elected thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// same with the A operand in tensor memory (lane = row m, 8 consecutive 32-bit columns = the K slice)
__device__ __forceinline__ void mma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// registers -> TMEM: this warp's 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"
               ::"r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                 "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                 "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
                 "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float a, float b, float c, float d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%4], {%0,%1,%2,%3};"
               ::"r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)), "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, float a) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%1], {%0};" ::"r"(__float_as_uint(a)), "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// all previously issued MMAs of this thread arrive on the mbarrier when complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns --------------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// two back-to-back 16-column loads (32 consecutive columns) with a single wait
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// issue only (no wait): pair with tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// split for 3xTF32: hi = a rounded to TF32 (round-to-nearest on the magnitude, so the residual is
// signed and the split is unbiased), lo = a - hi (exact in fp32; the tensor core keeps its top 11 bits)
__host__ __device__ __forceinline__ float tf32_hi(float a) {
#ifdef __CUDA_ARCH__
  return __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xFFFFE000u);
#else
  uint32_t b; memcpy(&b, &a, 4); b = (b + 0x1000u) & 0xFFFFE000u; float r; memcpy(&r, &b, 4); return r;
#endif
}
__host__ __device__ __forceinline__ float tf32_lo(float a) { return a - tf32_hi(a); }

}  // namespace umma
