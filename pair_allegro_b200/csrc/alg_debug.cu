// Debug / unit-test entry points of the C-ABI library (not part of the force path).
//   alg_debug_umma_gemm : C[128][N] = A^T W on the tcgen05 tensor cores (kind::tf32, 1 or 3
//   passes), operands staged exactly the way the pipeline kernels stage them (umma.cuh).
#include <cstdio>
#include <vector>

#include "../../include/allegro_b200.h"
#include <cuda_bf16.h>

#include "umma.cuh"

extern "C" ALG_API int alg_debug_umma_gemm2(const float* A, const float* W, float* C, int K, int N, int passes, int variant, float* dump128);

namespace {
using umma::opk_idx;
using umma::make_desc_k_sw128;
__global__ void __launch_bounds__(256, 1) k_umma_test(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ C,
                                                      int K, int N, int passes, int variant, float* __restrict__ dump) {
  extern __shared__ __align__(1024) float sm_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  float* sm = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~uintptr_t(1023));
  float* A_hi = sm;
  float* A_lo = A_hi + 128 * K;
  float* W_hi = A_lo + 128 * K;
  float* W_lo = W_hi + N * K;
  const bool kmajor = true;
  const bool bf16ns = variant & 2;
  if (variant & 4) {   // TMEM store/load round trip only
    if (warp == 0) umma::tmem_alloc(&tmem_base_s, 128);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tm = tmem_base_s;
    const int q = warp & 3, half = warp >> 2;
    const int m = q * 32 + lane;
    for (int c0 = half * 64; c0 < (half + 1) * 64; c0 += 16) {
      uint32_t r[16];
      for (int i = 0; i < 16; ++i) r[i] = __float_as_uint((float)(m * 1000 + c0 + i));
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"
                   ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                     "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(tm + ((uint32_t)(q * 32) << 16) + (uint32_t)c0) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    for (int c0 = half * 64; c0 < (half + 1) * 64; c0 += 16) {
      float v[16];
      umma::tmem_ld16(tm + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      for (int i = 0; i < 16; ++i) dump[(size_t)m * 128 + c0 + i] = v[i];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tm, 128);
    return;
  }
  if (bf16ns) {
    // bf16, K-major, no swizzle: core matrix = 8 rows x 16 B; LBO (K dir) = 128 B, SBO (row-group dir) = K/8*128 B
    __nv_bfloat16* Ab = reinterpret_cast<__nv_bfloat16*>(A_hi);
    __nv_bfloat16* Wb = reinterpret_cast<__nv_bfloat16*>(W_hi);
    for (int i = t; i < 128 * K; i += 256) {
      const int k = i / 128, m = i % 128;
      Ab[(m / 8) * (K * 8) + (k / 8) * 64 + (m % 8) * 8 + (k % 8)] = __float2bfloat16(A[i]);
    }
    for (int i = t; i < N * K; i += 256) {
      const int k = i / N, n = i % N;
      Wb[(n / 8) * (K * 8) + (k / 8) * 64 + (n % 8) * 8 + (k % 8)] = __float2bfloat16(W[i]);
    }
  } else
  for (int i = t; i < 128 * K; i += 256) {
    const int k = i / 128, m = i % 128;
    const float a = A[i];
    const int o = opk_idx(m, k, 128);
    A_hi[o] = umma::tf32_hi(a);
    A_lo[o] = umma::tf32_lo(a);
  }
  if (!bf16ns)
  for (int i = t; i < N * K; i += 256) {
    const int k = i / N, n = i % N;
    const float w = W[i];
    const int o = opk_idx(n, k, N);
    W_hi[o] = umma::tf32_hi(w);
    W_lo[o] = umma::tf32_lo(w);
  }
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 128);
  if (t == 0) umma::mbar_init(&bar, 1);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (t == 0 && (variant & 8)) printf("smem_raw %x sm %x A_hi %x W_hi %x tmem %x bar %x\n", umma::smem_u32(sm_raw), umma::smem_u32(sm), umma::smem_u32(A_hi), umma::smem_u32(W_hi), tmem, umma::smem_u32(&bar));
  if (t == 0 && bf16ns) {
    // kind::f16: A=B=BF16 (format 1), D=F32, K-major both, M=128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t acc = 0;
    for (int ks = 0; ks < K / 16; ++ks) {
      auto mk = [&](const void* p) {
        uint64_t d = 0;
        d |= (uint64_t)((umma::smem_u32(p) >> 4) & 0x3FFF);
        d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;                       // LBO: next core matrix along K
        d |= (uint64_t)((((uint32_t)K / 8u * 128u) >> 4) & 0x3FFF) << 32;   // SBO: next 8-row group
        d |= (uint64_t)1 << 46;
        return d;
      };
      const uint64_t da = mk(reinterpret_cast<const char*>(A_hi) + ks * 256);
      const uint64_t db = mk(reinterpret_cast<const char*>(W_hi) + ks * 256);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      acc = 1;
    }
    umma::mma_commit(&bar);
  } else
  if (t == 0) {
    uint32_t idesc = umma::make_idesc_tf32(N);
    uint32_t acc = 0;
    for (int p = 0; p < passes; ++p) {
      const float* Ap = (passes == 3 && p == 0) ? A_lo : A_hi;
      const float* Wp = (passes == 3 && p == 1) ? W_lo : W_hi;
      for (int kg = 0; kg < K / 8; ++kg) {
        uint64_t da, db;
        da = make_desc_k_sw128(Ap + (kg >> 2) * (128 * 32) + (kg & 3) * 8);
        db = make_desc_k_sw128(Wp + (kg >> 2) * (N * 32) + (kg & 3) * 8);
        umma::mma_tf32(tmem, da, db, idesc, acc);
        acc = 1;
      }
    }
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::fence_after_sync();
  {
    const int q = warp & 3, half = warp >> 2;
    const int m = q * 32 + lane;
    for (int c0 = half * (N / 2); c0 < (half + 1) * (N / 2); c0 += 16) {
      float v[16];
      umma::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      for (int i = 0; i < 16; ++i) C[(size_t)m * N + c0 + i] = v[i];
    }
    if (dump) {   // all 128 allocated columns of every lane
      for (int c0 = half * 64; c0 < (half + 1) * 64; c0 += 16) {
        float v[16];
        umma::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        for (int i = 0; i < 16; ++i) dump[(size_t)m * 128 + c0 + i] = v[i];
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 128);
}
}  // namespace

extern "C" ALG_API int alg_debug_umma_gemm(const float* A, const float* W, float* C, int K, int N, int passes) {
  return alg_debug_umma_gemm2(A, W, C, K, N, passes, 0, nullptr);
}

extern "C" ALG_API int alg_debug_umma_gemm2(const float* A, const float* W, float* C, int K, int N, int passes, int variant, float* dump128) {
  if (!A || !W || !C || K % 32 || N % 32 || N > 128 || K < 32 || (passes != 1 && passes != 3)) return ALG_EINVAL;
  if (K % 32) return ALG_EINVAL;
  if ((variant & 2) && K % 16) return ALG_EINVAL;
  const size_t smem = sizeof(float) * 2 * (size_t)(128 + N) * K + 2048;
  if (smem > 227 * 1024) return ALG_EINVAL;
  float *dA, *dW, *dC;
  if (cudaMalloc(&dA, sizeof(float) * 128 * K) != cudaSuccess) return ALG_ECUDA;
  cudaMalloc(&dW, sizeof(float) * N * K);
  cudaMalloc(&dC, sizeof(float) * 128 * N);
  cudaMemcpy(dA, A, sizeof(float) * 128 * K, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W, sizeof(float) * N * K, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k_umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  float* dD = nullptr;
  if (dump128) { cudaMalloc(&dD, sizeof(float) * 128 * 128); cudaMemset(dD, 0, sizeof(float) * 128 * 128); }
  k_umma_test<<<1, 256, smem>>>(dA, dW, dC, K, N, passes, variant, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) cudaMemcpy(C, dC, sizeof(float) * 128 * N, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && dump128) cudaMemcpy(dump128, dD, sizeof(float) * 128 * 128, cudaMemcpyDeviceToHost);
  cudaFree(dA); cudaFree(dW); cudaFree(dC); if (dD) cudaFree(dD);
  if (e != cudaSuccess) { fprintf(stderr, "alg_debug_umma_gemm: %s\n", cudaGetErrorString(e)); return ALG_ECUDA; }
  return ALG_OK;
}

// ---- tcgen05.mma issue-rate microbenchmark -----------------------------------------------------------
// every CTA repeats `iters` times: issue 8*groups kind::tf32 MMAs (M=128, N, K=8 each; operands = whatever is in
// shared memory; group g accumulates into TMEM accumulator g % nacc), commit, (sync_each ? wait : continue).
// cycles[0] = clock64 ticks of CTA 0 from the first issue to the last completion wake-up, cycles[1] = time spent
// inside the issue loops.  blocks_per_sm = 2 shows the contention the pipeline kernels see.
namespace {
__global__ void __launch_bounds__(128, 2) k_mma_rate(int N, int nacc, int groups, int iters, int sync_each, long long* cycles) {
  extern __shared__ __align__(1024) float sm_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  float* sm = sm_raw + (((1024u - (umma::smem_u32(sm_raw) & 1023u)) & 1023u) >> 2);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  for (int i = threadIdx.x; i < 128 * 64 + 256 * 64; i += blockDim.x) sm[i] = 1.0f;
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 256);
  if (threadIdx.x == 0) umma::mbar_init(&bar, 1);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  long long t0 = clock64(), tissue = 0;
  uint32_t ph = 0;
  const uint32_t sbase = __shfl_sync(0xffffffffu, umma::smem_u32(sm), 0);
  const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base_s, 0);
  const uint32_t mb = __shfl_sync(0xffffffffu, umma::smem_u32(&bar), 0);
  const uint64_t dA = umma::make_desc_k_sw128_addr(sbase), dW = umma::make_desc_k_sw128_addr(sbase + 128 * 64 * 4);
  const uint32_t idesc = umma::make_idesc_tf32(N);
  const uint32_t wpan = (uint32_t)(N * 32 * 4) >> 4;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (warp == 0) {
      if (umma::elect_one()) {
        const long long ti = clock64();
        uint32_t acc = 0;
#pragma unroll 1
        for (int g = 0; g < groups; ++g) {
          const uint32_t td = tm + acc * (uint32_t)N;
          umma::mma_tf32(td, dA, dW, idesc, 1);
          umma::mma_tf32(td, dA + 2, dW + 2, idesc, 1);
          umma::mma_tf32(td, dA + 4, dW + 4, idesc, 1);
          umma::mma_tf32(td, dA + 6, dW + 6, idesc, 1);
          umma::mma_tf32(td, dA + 1024, dW + wpan, idesc, 1);
          umma::mma_tf32(td, dA + 1026, dW + wpan + 2, idesc, 1);
          umma::mma_tf32(td, dA + 1028, dW + wpan + 4, idesc, 1);
          umma::mma_tf32(td, dA + 1030, dW + wpan + 6, idesc, 1);
          acc = (acc + 1 == (uint32_t)nacc) ? 0 : acc + 1;
        }
        if (sync_each || it == iters - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mb) : "memory");
        tissue += clock64() - ti;
      }
      __syncwarp();
    }
    if (sync_each || it == iters - 1) {
      umma::mbar_wait(&bar, ph);
      ph ^= 1;
      umma::fence_after_sync();
    }
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = clock64() - t0;
  if (tissue && blockIdx.x == 0) cycles[1] = tissue;
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base_s, 256);
}
}  // namespace

extern "C" ALG_API int alg_debug_mma_rate(int N, int nacc, int groups, int iters, int sync_each, int nblocks, long long* cycles2) {
  long long* d;
  if (cudaMalloc(&d, 2 * sizeof(long long)) != cudaSuccess) return -1;
  cudaMemset(d, 0, 2 * sizeof(long long));
  const int smem = (128 * 64 + 256 * 64) * 4 + 1024;
  cudaFuncSetAttribute(k_mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_mma_rate<<<nblocks, 128, smem>>>(N, nacc, groups, iters, sync_each, d);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(cycles2, d, 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? 0 : -(int)e;
}

// ---- A operand in tensor memory: correctness + rate -------------------------------------------------
// C[128][N] = A^T W with A (given as A[k*128+m]) split hi/lo and written to TMEM by the thread that owns row m
// (columns [128,128+K) = hi, [192,192+K) = lo), W in shared memory as in the pipeline, D in TMEM columns [0,N).
// rate_iters > 0 additionally repeats the 3-pass MMA group `rate_iters` times and reports cycles per MMA.
namespace {
__global__ void __launch_bounds__(128, 2) k_umma_ta_test(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ C,
                                                         int K, int N, int passes, int rate_iters, long long* cycles) {
  extern __shared__ __align__(1024) float sm_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  float* sm = sm_raw + (((1024u - (umma::smem_u32(sm_raw) & 1023u)) & 1023u) >> 2);
  float* W_hi = sm;
  float* W_lo = sm + 64 * 64;
  const int t = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  for (int i = t; i < N * K; i += blockDim.x) {
    const int k = i / N, n = i % N;
    const float w = W[i];
    const int o = umma::opk_idx(n, k, N);
    W_hi[o] = umma::tf32_hi(w);
    W_lo[o] = umma::tf32_lo(w);
  }
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 256);
  if (t == 0) umma::mbar_init(&bar, 1);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  {  // row m = t: K values -> TMEM
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int k0 = 0; k0 < K; k0 += 16) {
      float hi[16], lo[16];
      for (int i = 0; i < 16; ++i) { const float a = A[(size_t)(k0 + i) * 128 + t]; hi[i] = umma::tf32_hi(a); lo[i] = a - hi[i]; }
      umma::tmem_st16(lane_base + 128 + k0, hi);
      umma::tmem_st16(lane_base + 192 + k0, lo);
    }
    umma::tmem_st_wait();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  long long t0 = 0;
  uint32_t ph = 0;
  const uint32_t sbase = __shfl_sync(0xffffffffu, umma::smem_u32(sm), 0);
  const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
  const uint32_t mb = __shfl_sync(0xffffffffu, umma::smem_u32(&bar), 0);
  const uint64_t dWh = umma::make_desc_k_sw128_addr(sbase), dWl = umma::make_desc_k_sw128_addr(sbase + 64 * 64 * 4);
  const uint32_t idesc = umma::make_idesc_tf32(N);
  const uint32_t wpan = (uint32_t)(N * 32 * 4) >> 4;
  const int reps = rate_iters > 0 ? rate_iters : 1;
  if (warp == 0) {
    if (umma::elect_one()) {
      t0 = clock64();
      auto group = [&](uint32_t ta, uint64_t db, uint32_t acc) {      // K/8 MMAs, constant offsets
        umma::mma_tf32_ta(tm, ta, db, idesc, acc);
        umma::mma_tf32_ta(tm, ta + 8, db + 2, idesc, 1);
        umma::mma_tf32_ta(tm, ta + 16, db + 4, idesc, 1);
        umma::mma_tf32_ta(tm, ta + 24, db + 6, idesc, 1);
        if (K > 32) {
          umma::mma_tf32_ta(tm, ta + 32, db + wpan, idesc, 1);
          umma::mma_tf32_ta(tm, ta + 40, db + wpan + 2, idesc, 1);
          umma::mma_tf32_ta(tm, ta + 48, db + wpan + 4, idesc, 1);
          umma::mma_tf32_ta(tm, ta + 56, db + wpan + 6, idesc, 1);
        }
      };
#pragma unroll 1
      for (int it = 0; it < reps; ++it) {
        const uint32_t acc0 = (it == 0) ? 0u : 1u;
        if (passes == 3) {
          group(tm + 192, dWh, acc0);
          group(tm + 128, dWl, 1);
          group(tm + 128, dWh, 1);
        } else {
          group(tm + 128, dWh, acc0);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mb) : "memory");
    }
    __syncwarp();
  }
  umma::mbar_wait(&bar, ph);
  umma::fence_after_sync();
  if (t == 0 && blockIdx.x == 0 && cycles) cycles[0] = clock64() - t0;
  {
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16];
      umma::tmem_ld16(lane_base + c0, v);
      if (blockIdx.x == 0) for (int i = 0; i < 16; ++i) C[(size_t)t * N + c0 + i] = v[i];
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}
}  // namespace

extern "C" ALG_API int alg_debug_umma_gemm_ta(const float* A, const float* W, float* C, int K, int N, int passes, int rate_iters, int nblocks,
                                              long long* cycles) {
  if (K % 16 || K > 64 || N % 16 || N > 64) return ALG_EINVAL;
  float *dA, *dW, *dC; long long* dcy;
  cudaMalloc(&dA, sizeof(float) * 128 * K); cudaMalloc(&dW, sizeof(float) * N * K); cudaMalloc(&dC, sizeof(float) * 128 * N); cudaMalloc(&dcy, 8);
  cudaMemcpy(dA, A, sizeof(float) * 128 * K, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W, sizeof(float) * N * K, cudaMemcpyHostToDevice);
  cudaMemset(dcy, 0, 8);
  const int smem = 2 * 64 * 64 * 4 + 1024;
  cudaFuncSetAttribute(k_umma_ta_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_umma_ta_test<<<nblocks, 128, smem>>>(dA, dW, dC, K, N, passes, rate_iters, dcy);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(C, dC, sizeof(float) * 128 * N, cudaMemcpyDeviceToHost);
  if (cycles) cudaMemcpy(cycles, dcy, 8, cudaMemcpyDeviceToHost);
  cudaFree(dA); cudaFree(dW); cudaFree(dC); cudaFree(dcy);
  return e == cudaSuccess ? 0 : -(int)e;
}
