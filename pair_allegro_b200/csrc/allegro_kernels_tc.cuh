// Tensor-core (tcgen05) variant of the five fused edge-tile kernels, l_max = 1 and 2.
//
// Same staging and the same HBM buffers as allegro_kernels.cuh (F0 / FK / T / BK / B0, kernel
// boundaries at the per-centre environment sums), but every dense contraction runs on the
// 5th-generation tensor cores:
//   * tile = 128 edges = the M dimension of one tcgen05.mma (cta_group::1), 256 threads;
//     thread (m = t%128, half = t/128) owns edge row m and one half of the output columns in
//     every epilogue (TMEM lane m is readable by warps w with w%4 == m/32);
//   * A operands (activations) are written by their producer thread straight into the K-major
//     SWIZZLE_128B shared-memory layout of umma.cuh, B operands (weights) are pre-swizzled on
//     the host and fetched by one TMA bulk copy (cp.async.bulk) per GEMM block, overlapped with
//     the previous epilogue;
//   * accumulators live in TMEM.  k_t_tc (forward + backward of the last layer in one kernel)
//     keeps z1, z2 and the pre-envelope MLP output m in TMEM and re-evaluates act'(z) from there;
//     k_f0_tc / k_fk_tc write act'(z1), act'(z2), m to the activation record a.ZD so that
//     k_b0_tc / k_bk_tc run the backward without recomputing the forward;
//   * MMAs are issued by one elected lane of warp 0 from warp-uniform descriptors (tc_mma);
//   * strict mode = 3xTF32 (a = hi + lo split of both operands, lo*hi + hi*lo + hi*hi
//     accumulated in fp32): error ~5e-7, i.e. fp32-level (tests/test_gpu_umma.py);
//     fast mode = single TF32 pass;
//   * l_max = 2: the l-indexed width 32*(l_max+1) = 96 is processed as 64-wide blocks (DimsTC::NB).
#pragma once
#include "alg_common.cuh"
#include "tp_gen.cuh"
#include "umma.cuh"

#ifndef ALG_TP_DBUF_L3
#define ALG_TP_DBUF_L3 0
#endif
#ifndef ALG_TP_FWD_DBUF_L3
#define ALG_TP_FWD_DBUF_L3 1
#endif
#ifndef ALG_TP_L1_PREFETCH
#define ALG_TP_L1_PREFETCH 0
#endif
#ifndef ALG_TP_DBUF_BELOW
#define ALG_TP_DBUF_BELOW 3
#endif
namespace alg {

// the large building blocks: inlined.  (Measured on the B200: as real functions -- __noinline__, one copy each -- the
// chunked pipeline went from 21.5 to 33.8 ms on a 256 k-atom box: the context struct lands in local memory and every
// call spills the live accumulators.)
#define ALG_NI __forceinline__

struct TcMat { const float* hi; const float* lo; int N, K; };   // smem image: K/32 panels x N rows x 32 floats
// every image is one MMA block: N <= 64 rows, K <= 64 columns (16 KB) so that a CTA needs only
// ~104 KB of shared memory and two CTAs share an SM (their phases overlap)
// matrices whose l-indexed dimension (width 32*(l_max+1)) exceeds 64 are split into blocks b = 0,1:
// block b covers l in [2b, 2b+2) i.e. columns [64b, 64b + BW(b))
struct TcLayerW { TcMat m0x, m0s[2], m1, m2, env[2], m2_b, m1_b, m0_bx, m0_bs[2], env_b[2]; };
struct TcW {
  TcMat two0, two1, two2, emb[2], two2_b, two1_b, two0_b, emb_b[2], ro0, ro0_b;
  TcLayerW layer[3];
  int passes;     // 3 = strict (3xTF32), 1 = fast (TF32)
};

template <int L> struct DimsTC {
  static constexpr int TM = 128;
  static constexpr int NSH = (L + 1) * (L + 1);
  static constexpr int NL = L + 1;
  static constexpr int ENVW = NL * U;
  static constexpr int SIN = S + NL * U;
  static constexpr int F = NSH * U;
  static constexpr int CPH = NT / TM;      // 2
  static constexpr int CPT = U / CPH;      // 16
  static constexpr int NB = (NL + 1) / 2;               // 64-wide blocks of the l-indexed width
  static constexpr int WS = 64 + 1;                     // W_s stride (one block at a time)
  static constexpr int CHU = (L == 1) ? 16 : 8;         // channels per dGamma staging pass
  static constexpr int FC = NSH * CHU;
  static constexpr int DGS = FC + 1;
  static constexpr int TB = (L == 1) ? 4 : 1;           // tensor-product channels whose loads are batched
  static constexpr int MINB = (L == 3) ? 1 : 2;         // resident CTAs per SM (__launch_bounds__)
  __host__ __device__ static constexpr int bw(int b) { return (ENVW - 64 * b) < 64 ? (ENVW - 64 * b) : 64; }   // block width
  __host__ __device__ static constexpr int lhi(int b) { return (2 * b + 2) < NL ? (2 * b + 2) : NL; }            // block covers l in [2b, lhi)
};

template <int L> struct SmemTC {
  using D = DimsTC<L>;
  static constexpr int TM = 128;
  // the A operand lives in tensor memory; [oOPH, oWBH) is a 64-KB scratch region (env-weight / ds / dG staging,
  // reduction scratch) whose two halves double as the hi / lo image of weight buffer 2 (tc_load_w2)
  static constexpr int OPF = TM * 64;
  static constexpr int WBF = 4096;                     // weight block capacity: N*K <= 64*64
  // scratch floats: l_max <= 2: 2*OPF (64 KB, two CTAs per SM); l_max = 3: dG staging [128][129] + ds rows [128][128]
  // (130 KB -> one CTA per SM, which is what the persistent fused kernel wants anyway)
  static constexpr int DGPAD = ((D::DGS * TM + 127) / 128) * 128;
  static constexpr int SCR = (L <= 2) ? 2 * OPF : ((DGPAD + D::ENVW * TM + 255) / 256) * 256;   // weight images behind it: 1024-byte aligned
  static constexpr int oOPH = 0;
  static constexpr int oOPL = oOPH + OPF;
  static constexpr int oWBH = SCR;
  static constexpr int oWBL = oWBH + WBF;
  static constexpr int oY = oWBL + WBF;
  static constexpr int oDY = oY + D::NSH * TM;
  static constexpr int oU = oDY + D::NSH * TM;
  static constexpr int oC = oU + TM;                   // int
  static constexpr int oZZ = oC + TM;                  // int
  static constexpr int oE = oZZ + TM;                  // 4*TM floats: per-half partials, E_e, du partial
  static constexpr int GSROWS = (L == 1) ? 12 : 0;     // staged per-centre rows (Gamma / dGamma); 0 = always read from global
  static constexpr int oGS = oE + 4 * TM;
  static constexpr int GSSTRIDE = D::F + 4;            // staged row stride: +4 floats so that the rows of different centres start in
                                                       // different banks (lanes of one warp read feature f of 1-3 distinct centres)
  static constexpr int oSEG = oGS + GSROWS * GSSTRIDE; // int: segment table seg[TM+1], nseg, warp counts (TM + 16 ints)
  static constexpr int oBAR = oSEG + TM + 16;          // 3 mbarriers + tmem pointer (8 floats: mbar, wbar, tmem ptr, wbar2), fused kernel: tile number
  static constexpr int TOTAL = oBAR + 12 + 12;         // + the fused kernel's batch / loop state (FS_*)
  static constexpr size_t BYTES = (size_t)TOTAL * sizeof(float) + 1024;   // + alignment slack
  // buffers inside the scratch / weight regions (live ranges never overlap a weight block that is still in use)
  static constexpr int oWS = oOPL;                                 // env weights of one block, edge-major [128][65]
  static constexpr int oDS = (L == 1) ? oWBH : DGPAD;              // ds rows [q*U+u][128] for the tensor-product backward
                                                                   // (l_max >= 2: right behind the dG staging)
  static constexpr int oDG = oOPH;                                 // dG staging [128][DGS]
  static_assert(L >= 1 && L <= 3, "tensor-core pipeline: l_max = 1..3");
  static_assert(D::WS * TM <= OPF + 2 * WBF, "W_s must fit OPL + weight region");
  static_assert(L <= 2 || oWS + D::WS * TM <= SCR, "l_max = 3: W_s inside the scratch region");
  static_assert(oDS + D::ENVW * TM <= oY, "DS_s must end before the persistent small arrays");
  static_assert(L <= 2 || oDS + D::ENVW * TM <= SCR, "l_max = 3: DS_s inside the scratch region (the weight buffers stay live)");
  static_assert(oDG + D::DGS * TM <= oDS || L == 1, "dG staging must not overlap DS_s");
  static_assert(L != 1 || D::DGS * TM <= 2 * OPF, "dG staging (l_max = 1) spans OPH + OPL");
  static_assert(L == 3 || BYTES <= 113 * 1024, "two CTAs per SM");
  static_assert(BYTES <= 227 * 1024, "shared memory per CTA");
  static_assert(oWBH % 256 == 0 && oWBL % 256 == 0 && oOPL % 256 == 0, "SWIZZLE_128B operand images need 1024-byte alignment");
};

// per-centre rows (Gamma_k or dGamma_k) of the centres touched by this tile: staged in shared memory
// when the tile spans <= GSROWS consecutive centre slots (the normal case), else read from global
struct RowSrc {
  const float* base;   // staged: smem row of centre cmin ; else global row of centre c0
  int c_origin;        // cmin or c0
  int stride;          // floats between consecutive centres (staged rows are padded: see tc_stage_rows)
  __device__ __forceinline__ const float* row(int centre, int /*F*/) const { return base + (size_t)(centre - c_origin) * stride; }
};


struct TcCtx {
  float* sm;
  uint32_t tmem;
  uint64_t* mbar;
  uint64_t* wbar;
  uint64_t* wbar2;     // second weight buffer (tc_load_w2 / tc_mma_pair)
  uint32_t mph, wph, wph2;
  int passes;
  int m, half, q;
  int c0;              // first centre slot of the Gamma / dGamma rows this CTA addresses (chunk origin, or the tile's first centre)
  size_t goff;         // float offset of this CTA's private Gamma / dGamma rows (fused kernel; 0 in the chunked pipeline)
};

template <int L> __device__ __forceinline__ TcCtx tc_begin(float* sm_raw, const TcW& tw) {
  using SM = SmemTC<L>;
  TcCtx c;
  // 1024-byte alignment by pointer arithmetic on the __shared__ array (an integer round trip would hide the
  // address space from the compiler and turn every LDS/STS into a generic LD/ST)
  c.sm = sm_raw + (((1024u - (umma::smem_u32(sm_raw) & 1023u)) & 1023u) >> 2);
  c.mbar = reinterpret_cast<uint64_t*>(c.sm + SM::oBAR);
  c.wbar = c.mbar + 1;
  c.wbar2 = c.mbar + 3;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(c.mbar + 2);
  const int t = threadIdx.x;
  if ((t >> 5) == 0) umma::tmem_alloc(tptr, 256);
  if (t == 0) { umma::mbar_init(c.mbar, 1); umma::mbar_init(c.wbar, 1); umma::mbar_init(c.wbar2, 1); }
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  c.tmem = *tptr;
  c.mph = c.wph = c.wph2 = 0;
  c.passes = tw.passes;
  c.m = t & 127; c.half = t >> 7; c.q = (t >> 5) & 3;
  c.c0 = 0; c.goff = 0;
  return c;
}
__device__ __forceinline__ void tc_end(TcCtx& c) {
  umma::fence_before_sync();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) umma::tmem_dealloc(c.tmem, 256);
}

// thread 0: fetch the next GEMM's weight image(s) with TMA bulk copies (call after the previous MMA completed)
template <int L> __device__ __forceinline__ void tc_load_w(TcCtx& c, const TcMat& w) {
  using SM = SmemTC<L>;
  if (threadIdx.x == 0) {
    const uint32_t bytes = (uint32_t)(w.N * w.K) * 4u;
    umma::mbar_expect_tx(c.wbar, c.passes == 3 ? 2 * bytes : bytes);
    umma::bulk_g2s(c.sm + SM::oWBH, w.hi, bytes, c.wbar);
    if (c.passes == 3) umma::bulk_g2s(c.sm + SM::oWBL, w.lo, bytes, c.wbar);
  }
}

// all threads: operands are written -> one elected lane of warp 0 issues the MMAs -> everybody waits for completion.
// Everything the issuing code touches is made provably warp-uniform (warp index and base addresses pass through
// __shfl_sync, the branch is on the warp index, the lane is chosen by elect.sync): the descriptors then live in
// uniform registers and every tcgen05.mma is a single UTCHMMA -- with a per-thread `threadIdx.x == 0` branch
// ptxas wraps each MMA in an ELECT / R2UR waterfall loop (~170 cycles per MMA measured).
// second weight buffer = the scratch region at the start of the shared-memory plan (oOPH / oOPL), free whenever no
// W_s / DS_s / dG staging is live: lets two GEMMs that share their A operand be issued as one group (tc_mma_pair)
template <int L> __device__ __forceinline__ void tc_load_w2(TcCtx& c, const TcMat& w) {
  using SM = SmemTC<L>;
  if (threadIdx.x == 0) {
    const uint32_t bytes = (uint32_t)(w.N * w.K) * 4u;
    umma::mbar_expect_tx(c.wbar2, c.passes == 3 ? 2 * bytes : bytes);
    umma::bulk_g2s(c.sm + SM::oOPH, w.hi, bytes, c.wbar2);
    if (c.passes == 3) umma::bulk_g2s(c.sm + SM::oOPL, w.lo, bytes, c.wbar2);
  }
}

// TMEM column map (256 columns per CTA): A operand hi [0,64), lo [64,128); accumulators [128,192) and [192,256)
constexpr uint32_t TC_AHI = 0, TC_ALO = 64, TC_ACC = 128, TC_ACC2 = 192;
template <int L> __device__ __forceinline__ void tc_mma8(uint32_t td, uint32_t ta, uint64_t db, uint32_t idesc, uint32_t wpan, int K, uint32_t acc) {
  // A: 8 k = 8 TMEM columns per MMA; B: start-address field in 16-byte units, 8 k = 32 B = 2 units, panel stride wpan
  umma::mma_tf32_ta(td, ta, db, idesc, acc);
  umma::mma_tf32_ta(td, ta + 8, db + 2, idesc, 1);
  umma::mma_tf32_ta(td, ta + 16, db + 4, idesc, 1);
  umma::mma_tf32_ta(td, ta + 24, db + 6, idesc, 1);
  if (K > 32) {
    umma::mma_tf32_ta(td, ta + 32, db + wpan, idesc, 1);
    umma::mma_tf32_ta(td, ta + 40, db + wpan + 2, idesc, 1);
    umma::mma_tf32_ta(td, ta + 48, db + wpan + 4, idesc, 1);
    umma::mma_tf32_ta(td, ta + 56, db + wpan + 6, idesc, 1);
  }
}
// MMA issue of one GEMM block: a REAL function (one copy), called by warp 0 only, all arguments warp-uniform scalars.
// Inlined, the descriptor arithmetic + 24 UTCHMMA of every call site were ~4.5 KB of code that 7 of 8 warps skip: half of
// the T phase's 87 KB, which made the two resident CTAs of the fused kernel (in different phases) miss the instruction cache.
//   wh / wl : shared-space byte addresses of the weight image (hi / lo); tm: TMEM base; commit_mbar: 0 = no commit
static __device__ __noinline__ void tc_issue(uint32_t wh, uint32_t wl, uint32_t tm, int passes, int K, int N, uint32_t dcol, uint32_t accumulate,
                                      uint32_t commit_mbar) {
  const uint32_t td = tm + dcol;
  const uint64_t dWh = umma::make_desc_k_sw128_addr(wh), dWl = umma::make_desc_k_sw128_addr(wl);
  const uint32_t idesc = umma::make_idesc_tf32(N);
  const uint32_t wpan = (uint32_t)(N * 32 * 4) >> 4;               // weight panel stride in 16-byte units
  if (umma::elect_one()) {
    if (passes == 3) {                                              // lo*hi, hi*lo, hi*hi
      tc_mma8<1>(td, tm + TC_ALO, dWh, idesc, wpan, K, accumulate);
      tc_mma8<1>(td, tm + TC_AHI, dWl, idesc, wpan, K, 1);
      tc_mma8<1>(td, tm + TC_AHI, dWh, idesc, wpan, K, 1);
    } else {
      tc_mma8<1>(td, tm + TC_AHI, dWh, idesc, wpan, K, accumulate);
    }
    if (commit_mbar) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(commit_mbar) : "memory");
  }
  __syncwarp();
}
template <int L> __device__ __forceinline__ void tc_mma(TcCtx& c, int K, int N, uint32_t dcol, uint32_t accumulate = 0) {
  using SM = SmemTC<L>;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp == 0) umma::mbar_wait(c.wbar, c.wph);      // weight block landed (requested one epilogue ago)
  umma::tmem_st_wait();                                // this thread's operand columns are in tensor memory
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    umma::fence_after_sync();
    const uint32_t sbase = __shfl_sync(0xffffffffu, umma::smem_u32(c.sm), 0);
    const uint32_t tm = __shfl_sync(0xffffffffu, c.tmem, 0);
    const uint32_t mbar = __shfl_sync(0xffffffffu, umma::smem_u32(c.mbar), 0);
    const int passes = __shfl_sync(0xffffffffu, c.passes, 0);
    tc_issue(sbase + SM::oWBH * 4, sbase + SM::oWBL * 4, tm, passes, K, N, dcol, accumulate, mbar);
  }
  c.wph ^= 1;
  umma::mbar_wait(c.mbar, c.mph);
  c.mph ^= 1;
  umma::fence_after_sync();
}

// two GEMMs on the SAME A operand (K columns), weights in buffer 1 (N1 -> dcol1) and buffer 2 (N2 -> dcol2):
// one barrier, one MMA group, one commit, one wake-up
template <int L> __device__ __forceinline__ void tc_mma_pair(TcCtx& c, int K, int N1, uint32_t dcol1, int N2, uint32_t dcol2) {
  using SM = SmemTC<L>;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp == 0) { umma::mbar_wait(c.wbar, c.wph); umma::mbar_wait(c.wbar2, c.wph2); }
  umma::tmem_st_wait();
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    umma::fence_after_sync();
    const uint32_t sbase = __shfl_sync(0xffffffffu, umma::smem_u32(c.sm), 0);
    const uint32_t tm = __shfl_sync(0xffffffffu, c.tmem, 0);
    const uint32_t mbar = __shfl_sync(0xffffffffu, umma::smem_u32(c.mbar), 0);
    const int passes = __shfl_sync(0xffffffffu, c.passes, 0);
    tc_issue(sbase + SM::oWBH * 4, sbase + SM::oWBL * 4, tm, passes, K, N1, dcol1, 0, 0);
    tc_issue(sbase + SM::oOPH * 4, sbase + SM::oOPL * 4, tm, passes, K, N2, dcol2, 0, mbar);
  }
  c.wph ^= 1;
  c.wph2 ^= 1;
  umma::mbar_wait(c.mbar, c.mph);
  c.mph ^= 1;
  umma::fence_after_sync();
}

// this thread's row m, 16 columns starting at absolute TMEM column col
__device__ __forceinline__ void tc_ld16(const TcCtx& c, uint32_t col, float* v) {
  umma::tmem_ld16(c.tmem + ((uint32_t)(c.q * 32) << 16) + col, v);
}
// write 4 consecutive k (k4 % 4 == 0) of row m into the A operand in tensor memory (hi [+ lo]).
// hi / lo by Veltkamp splitting with packed-pair arithmetic (4 FFMA2/FMUL2 per PAIR): t = 8193 a, hi = t - (t - a) keeps the
// top 11 significant bits of a (round to nearest -> exactly representable in TF32), lo = a - hi is exact.
__device__ __forceinline__ void tf32_split2(float2 a, float2& hi, float2& lo) {
  const float2 m1 = make_float2(-1.f, -1.f);
  const float2 t = __fmul2_rn(a, make_float2(8193.f, 8193.f));
  const float2 u = __ffma2_rn(a, m1, t);
  hi = __ffma2_rn(u, m1, t);
  lo = __ffma2_rn(hi, m1, a);
}
template <int L> __device__ __forceinline__ void op_put4(const TcCtx& c, int k4, float a, float b, float d, float e) {
  const uint32_t t0 = c.tmem + ((uint32_t)(c.q * 32) << 16) + (uint32_t)k4;
  float2 h0, l0, h1, l1;
  tf32_split2(make_float2(a, b), h0, l0);
  tf32_split2(make_float2(d, e), h1, l1);
  umma::tmem_st4(t0 + TC_AHI, h0.x, h0.y, h1.x, h1.y);
  if (c.passes == 3) umma::tmem_st4(t0 + TC_ALO, l0.x, l0.y, l1.x, l1.y);
}
// one element of this thread's own row
template <int L> __device__ __forceinline__ void op_put1(const TcCtx& c, int /*row == c.m*/, int k, float a) {
  const uint32_t t0 = c.tmem + ((uint32_t)(c.q * 32) << 16) + (uint32_t)k;
  const float ah = umma::tf32_hi(a);
  umma::tmem_st1(t0 + TC_AHI, ah);
  if (c.passes == 3) umma::tmem_st1(t0 + TC_ALO, a - ah);
}
// epilogue over this thread's half of NC columns: fn(n, v0..v3) for 4 consecutive columns n..n+3
template <class Fn> __device__ __forceinline__ void tc_epi(const TcCtx& c, uint32_t dcol, int NC, Fn fn) {
  // every epilogue of this pipeline covers 64 columns: this thread's half = 32 columns, one TMEM load
  const int c0 = c.half * 32;
  float v[32];
  umma::tmem_ld32(c.tmem + ((uint32_t)(c.q * 32) << 16) + dcol + c0, v);
#pragma unroll
  for (int i = 0; i < 32; i += 4) fn(c0 + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
  (void)NC;
}
// block epilogue: bw = 64 (32 columns per thread) or 32 (16 columns per thread)
template <class Fn> __device__ __forceinline__ void tc_epi_bw(const TcCtx& c, uint32_t dcol, int bw, Fn fn) {
  if (bw == 64) { tc_epi(c, dcol, 64, fn); return; }
  const int c0 = c.half * 16;
  float v[16];
  umma::tmem_ld16(c.tmem + ((uint32_t)(c.q * 32) << 16) + dcol + c0, v);
#pragma unroll
  for (int i = 0; i < 16; i += 4) fn(c0 + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
}
// "row4" tile layout of the arrays that are only ever touched through the epilogue mapping (x^k, dX and the
// activation record ZD): element (row n, edge m) lives at ((n/4)*128 + m)*4 + n%4, so a thread's 4 consecutive
// rows are one 128-bit access and a warp covers 512 contiguous bytes
__device__ __forceinline__ void st_row4(float* g, int n4, int m, float a, float b, float d, float e) {
  *reinterpret_cast<float4*>(g + (((n4 >> 2) * 128 + m) << 2)) = make_float4(a, b, d, e);
}
// load a 64-row tile array (x^k, dX; row4 layout) of this tile into operand columns [0,64)
// split form: issue the 8 loads early (no TcCtx needed), write the operand later
__device__ __forceinline__ void ld_rows_x8(const float* g /*tile base, row4 layout*/, float4 (&v)[8]) {
  const float4* gp = reinterpret_cast<const float4*>(g) + ((threadIdx.x >> 7) * 8) * 128 + (threadIdx.x & 127);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = gp[i * 128];
}
template <int L> __device__ __forceinline__ void op_put_x8(const TcCtx& c, const float4 (&v)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) op_put4<L>(c, c.half * 32 + 4 * i, v[i].x, v[i].y, v[i].z, v[i].w);
}
template <int L> __device__ __forceinline__ void op_load_rows64(const TcCtx& c, const float* g /*tile base, row4 layout*/) {
  float4 v[8];
  ld_rows_x8(g, v);                                             // all loads in flight before the first use
  op_put_x8<L>(c, v);
}
// a block of `bw` rows (64 or 32) of a plain tile-SoA array [row][128] -> operand columns [0,bw)
template <int L> __device__ __forceinline__ void op_load_rows_bw(const TcCtx& c, const float* g, int bw) {
  if (bw == 64) {      // plain tile-SoA rows [row][128] (w0 / dw0 are indexed by channel in the tensor-product drivers)
    float v[32];
    const float* gp = g + (c.half * 32) * 128 + c.m;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gp[i * 128];
#pragma unroll
    for (int i = 0; i < 32; i += 4) op_put4<L>(c, c.half * 32 + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
    return;
  }
  float v[16];
  const float* gp = g + (c.half * 16) * 128 + c.m;
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = gp[i * 128];
#pragma unroll
  for (int i = 0; i < 16; i += 4) op_put4<L>(c, c.half * 16 + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
}
// this thread's 32 values (its column half) of a 64-row tile array in row4 layout, all loads issued together
template <bool CG = false>
__device__ __forceinline__ void ld_rows32(const TcCtx& c, const float* g, float* v) {
  const float4* gp = reinterpret_cast<const float4*>(g) + (c.half * 8) * 128 + c.m;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 q = CG ? __ldcg(gp + i * 128) : gp[i * 128];
    v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
  }
}

// all threads; c_s must be published.  Ends with a barrier.
template <int L> __device__ ALG_NI RowSrc tc_stage_rows(const TcCtx& c, const float* gbase, int c0, int nvalid) {
  using D = DimsTC<L>; using SM = SmemTC<L>;
  const int* c_s = reinterpret_cast<const int*>(c.sm + SM::oC);
  const int cmin = c_s[0];
  const int span = c_s[nvalid - 1] - cmin + 1;
  RowSrc r;
  if (span <= SM::GSROWS) {
    float* GS = c.sm + SM::oGS;
    const float4* src = reinterpret_cast<const float4*>(gbase + (size_t)(cmin - c0) * D::F);
    constexpr int RV = D::F / 4;                       // float4 per row
    for (int i = threadIdx.x; i < span * RV; i += NT) {
      const int rw = i / RV, cv = i - rw * RV;
      reinterpret_cast<float4*>(GS + rw * SM::GSSTRIDE)[cv] = src[i];
    }
    r.base = GS; r.c_origin = cmin; r.stride = SM::GSSTRIDE;
  } else {
    r.base = gbase; r.c_origin = c0; r.stride = D::F;
  }
  __syncthreads();
  return r;
}

// geometry of row m (both halves compute, half 0 publishes Y_s, u_s, c_s, zz_s)
// the first global loads of every kernel (edge vector, centre slot), issued before the TMEM allocation /
// barrier of tc_begin so that their DRAM latency overlaps the CTA start-up
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
template <bool L1> __device__ __forceinline__ void prefetch_lx(const void* p) { if (L1) prefetch_l1(p); else prefetch_l2(p); }
struct GeomIn { float4 rv; int centre; };
__device__ __forceinline__ GeomIn tc_geom_load(const ChunkArgs& a, int es, int nvalid) {
  const int e = es + min((int)(threadIdx.x & 127), nvalid - 1);
  GeomIn gi;
  gi.rv = a.rvec[e];
  gi.centre = a.edge_c[e];
  return gi;
}
// pull a contiguous per-tile array into L2 at kernel start (one 128-byte line per thread and step): the kernels
// are latency-bound, their later loads then hit L2 instead of DRAM
__device__ __forceinline__ void tc_prefetch(const float* p, int nfloats) {
  for (int i = threadIdx.x * 32; i < nfloats; i += NT * 32) prefetch_l2(p + i);
}
// the per-centre row (Gamma / dGamma, F floats) of this thread's centre: lane m pulls line (m mod F/32)
template <int L> __device__ __forceinline__ void tc_prefetch_row(const float* gbase, int c0, const GeomIn& gi) {
  using D = DimsTC<L>;
  prefetch_l2(gbase + (size_t)(gi.centre - c0) * D::F + (threadIdx.x % (D::F / 32)) * 32);
}
template <int L> __device__ __forceinline__ Geom tc_geom(const ChunkArgs& a, const ModelW& w, const TcCtx& c, const GeomIn& gi) {
  using D = DimsTC<L>; using SM = SmemTC<L>; constexpr int TM = 128;
  const float4 rv = gi.rv;
  Geom g;
  const int zz = __float_as_int(rv.w);
  g.zi = zz & 255; g.zj = zz >> 8;
  g.r = sqrtf(rv.x * rv.x + rv.y * rv.y + rv.z * rv.z);
  g.x = rv.x / g.r; g.y = rv.y / g.r; g.z = rv.z / g.r;
  g.rc = w.rc[g.zi * MAXT + g.zj];
  float dudx;
  poly_cutoff(g.r / g.rc, w.p, g.u, dudx);
  g.dudr = dudx / g.rc;
  if (c.half == 0) {
    float Y[D::NSH];
    sph_harm<L>(g.x, g.y, g.z, Y);
#pragma unroll
    for (int k = 0; k < D::NSH; ++k) c.sm[SM::oY + k * TM + c.m] = Y[k];
    c.sm[SM::oU + c.m] = g.u;
    reinterpret_cast<int*>(c.sm + SM::oC)[c.m] = gi.centre;
    reinterpret_cast<int*>(c.sm + SM::oZZ)[c.m] = zz;
  }
  return g;
}

// env-weight GEMM output of block b (TMEM columns [dcol, dcol+bw)) -> W_s (edge-major, stride WS)
template <int L> __device__ __forceinline__ void tc_env_to_ws(const TcCtx& c, uint32_t dcol, int bw) {
  using D = DimsTC<L>; using SM = SmemTC<L>;
  float* W_s = c.sm + SM::oWS;
  tc_epi_bw(c, dcol, bw, [&](int n, float v0, float v1, float v2, float v3) {
    float* p = W_s + c.m * D::WS + n;
    p[0] = v0; p[1] = v1; p[2] = v2; p[3] = v3;
  });
}
// Gamma partial sums of the features whose l lies in block b (columns of W_s = (l-2b)*U+u)
template <int L> __device__ __forceinline__ void tc_env_sum(const ChunkArgs& a, const ModelW& w, const TcCtx& c, int tile, int es, int b,
                                                             float* gamma) {
  using D = DimsTC<L>; using SM = SmemTC<L>; constexpr int TM = 128;
  const float* W_s = c.sm + SM::oWS;
  const float* Y_s = c.sm + SM::oY;
  const int* c_s = reinterpret_cast<const int*>(c.sm + SM::oC);
  const int* seg = reinterpret_cast<const int*>(c.sm + SM::oSEG);
  const bool contin = a.rowptr[c_s[0]] < es;
  float* carry = a.carry + (size_t)tile * D::F;
  const float sc = w.inv_sqrt_n;
  const int lm_lo = (2 * b) * (2 * b), lm_hi = D::lhi(b) * D::lhi(b);
  const int nseg = seg[TM + 1];
  const int NFb = (lm_hi - lm_lo) * U;
  for (int wi = threadIdx.x; wi < NFb * nseg; wi += NT) {
    const int f = wi % NFb, sgm = wi / NFb;
    const int lm = lm_lo + f / U, u = f % U;
    const float* wcol = W_s + (lsel(lm) - 2 * b) * U + u;
    const float* ycol = Y_s + lm * TM;
    const int e0 = seg[sgm], e1 = seg[sgm + 1];
    float acc = 0.f;
    int e = e0;
    for (; e + 4 <= e1; e += 4) {
      const float v0 = wcol[e * D::WS] * ycol[e], v1 = wcol[(e + 1) * D::WS] * ycol[e + 1];
      const float v2 = wcol[(e + 2) * D::WS] * ycol[e + 2], v3 = wcol[(e + 3) * D::WS] * ycol[e + 3];
      acc += v0; acc += v1; acc += v2; acc += v3;
    }
    for (; e < e1; ++e) acc += wcol[e * D::WS] * ycol[e];
    const int fg = lm * U + u;
    if (sgm == 0 && contin) carry[fg] = acc * sc; else gamma[(size_t)(c_s[e0] - c.c0) * D::F + fg] = acc * sc;
  }
}
// env-weight accumulators (block 0 in TMEM columns col0, block 1 in col1) -> Gamma
template <int L> __device__ ALG_NI void tc_env_finish(const ChunkArgs& a, const ModelW& w, TcCtx& c, int tile, int es, float* gamma,
                                                                uint32_t col0, uint32_t col1) {
  using D = DimsTC<L>;
#pragma unroll 1
  for (int b = 0; b < D::NB; ++b) {
    tc_env_to_ws<L>(c, b == 0 ? col0 : col1, D::bw(b));
    __syncthreads();
    tc_env_sum<L>(a, w, c, tile, es, b, gamma);
    if (b + 1 < D::NB) __syncthreads();              // W_s is rewritten by the next block
  }
}
// env linear of x (operand [0,64)) for all blocks -> Gamma.  In: weight block env[0] requested (buffer 1) and, for
// l_max = 2, env[1] requested into weight buffer 2 (tc_load_w2): both blocks run as one MMA group.  W_s aliases
// weight buffer 2 and is only written after the group has completed.
template <int L> __device__ ALG_NI void tc_env_all(const ChunkArgs& a, const ModelW& w, TcCtx& c, const TcMat* env, int tile, int es,
                                                             float* gamma) {
  using D = DimsTC<L>;
  if constexpr (D::NB == 1) tc_mma<L>(c, 64, D::bw(0), TC_ACC);
  else tc_mma_pair<L>(c, 64, D::bw(0), TC_ACC, D::bw(1), TC_ACC2);
  tc_env_finish<L>(a, w, c, tile, es, gamma, TC_ACC, TC_ACC2);
  (void)env;
}

// s-part operand column of (scalar path q, channel u): K index inside the 64-wide "s" block
__device__ __forceinline__ int s_col(int q, int u) { return q * U + u; }

// per-channel component vectors of V^k / dV^k (N components of edge e, channel block base `g`): when the (padded)
// component count is a multiple of 4 the components are packed in groups of four per edge ("comp4":
// ((cc/4)*128 + e)*4 + cc%4) so that a channel is NP/4 128-bit accesses and a warp covers 512 contiguous bytes;
// otherwise plain [cc][128].  N % 4 == 3 (7, 31: the l_max = 1 / 3 full-parity sets) is padded by one component: a quarter
// of the memory instructions of the l_max = 3 tensor-product phases for 3 % more bytes.
__host__ __device__ constexpr int vpad(int n) { return (n % 4 == 3) ? n + 1 : n; }
template <int N> __device__ __forceinline__ void vec_load(const float* g, int e, float* v) {
  constexpr int NP = vpad(N);
  if constexpr (NP % 4 == 0) {
#pragma unroll
    for (int q = 0; q < NP / 4; ++q) {
      const float4 t = *reinterpret_cast<const float4*>(g + ((q * 128 + e) << 2));
      v[4 * q] = t.x;
      if (4 * q + 1 < N) v[4 * q + 1] = t.y;
      if (4 * q + 2 < N) v[4 * q + 2] = t.z;
      if (4 * q + 3 < N) v[4 * q + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int cc = 0; cc < N; ++cc) v[cc] = g[cc * 128 + e];
  }
}
template <int N> __device__ __forceinline__ void vec_store(float* g, int e, const float* v) {
  constexpr int NP = vpad(N);
  if constexpr (NP % 4 == 0) {
#pragma unroll
    for (int q = 0; q < NP / 4; ++q)
      *reinterpret_cast<float4*>(g + ((q * 128 + e) << 2)) =
          make_float4(v[4 * q], 4 * q + 1 < N ? v[4 * q + 1] : 0.f, 4 * q + 2 < N ? v[4 * q + 2] : 0.f, 4 * q + 3 < N ? v[4 * q + 3] : 0.f);
  } else {
#pragma unroll
    for (int cc = 0; cc < N; ++cc) g[cc * 128 + e] = v[cc];
  }
}
template <int N, bool L1 = false> __device__ __forceinline__ void vec_prefetch(const float* g, int e) {
  constexpr int NP = vpad(N);
  if constexpr (NP % 4 == 0) {
#pragma unroll
    for (int q = 0; q < NP / 4; ++q) prefetch_lx<L1>(g + ((q * 128 + e) << 2));
  } else {
#pragma unroll
    for (int cc = 0; cc < N; ++cc) prefetch_lx<L1>(g + cc * 128 + e);
  }
}
// raw global inputs of one tensor-product channel: FIRST layers read the L+1 embed weights w0[l][u]
// (V^0 = w0 (x) Y is formed in registers), later layers read V^k[u][DIN]
template <int L, bool FIRST, int DIN> struct VinRaw {
  static constexpr int N = FIRST ? (L + 1) : DIN;
  float v[N];
  // pull the lines a later issue() will read into L2 (no registers held)
  template <bool L1 = false>
  __device__ __forceinline__ static void prefetch(const ChunkArgs& a, int tile, int k, int e, int u) {
    using D = DimsTC<L>; constexpr int TM = 128;
    if (FIRST) {
      const float* W0g = a.W0 + (size_t)tile * D::ENVW * TM;
#pragma unroll
      for (int l = 0; l <= L; ++l) prefetch_lx<L1>(W0g + (l * U + u) * TM + e);
    } else {
      vec_prefetch<DIN, L1>(a.V[k] + ((size_t)tile * U + u) * vpad(DIN) * TM, e);
    }
  }
  __device__ __forceinline__ void issue(const ChunkArgs& a, int tile, int k, int e, int u) {
    using D = DimsTC<L>; constexpr int TM = 128;
    if (FIRST) {
      const float* W0g = a.W0 + (size_t)tile * D::ENVW * TM;
#pragma unroll
      for (int l = 0; l <= L; ++l) v[l] = W0g[(l * U + u) * TM + e];
    } else {
      vec_load<DIN>(a.V[k] + ((size_t)tile * U + u) * vpad(DIN) * TM, e, v);
    }
  }
  __device__ __forceinline__ void expand(int e, const float* Y_s, float* Vin) const {
    constexpr int TM = 128;
    if (FIRST) {
#pragma unroll
      for (int l = 0; l <= L; ++l) {
#pragma unroll
        for (int lm = l * l; lm < (l + 1) * (l + 1); ++lm) Vin[lm] = v[l] * Y_s[lm * TM + e];
      }
    } else {
#pragma unroll
      for (int cc = 0; cc < DIN; ++cc) Vin[cc] = v[cc];
    }
  }
};

// ============================================================================================
// tensor product drivers: thread (edge e = m, channel half uh = half) owns a contiguous block of channels,
// processed in batches of TB channels.  The global loads of batch s+1 are issued before batch s is
// evaluated (register double buffering), so only the first batch exposes the memory latency.
// ============================================================================================
// forward for K-block b of the "s" operand: s[q] for l = q in [2b, lhi(b)) -> operand column (q-2b)*U+u.
// With WANT_V (block 0 only) the full product is evaluated and V^{k+1} stored; otherwise only the
// scalar paths (identical order in every kind: path q pairs irrep (q,(-1)^q) of V with l2 = q).
template <int L, char KIND, bool FIRST, bool WANT_V>
__device__ ALG_NI void tc_tp_forward(const ChunkArgs& a, const LayerW& lw, const TcCtx& c, int tile, int k, const RowSrc& gsrc, int b) {
  using D = DimsTC<L>; using SM = SmemTC<L>; using TP = tpgen::TP<L, KIND>; using TPA = tpgen::TP<L, 'A'>; constexpr int TM = 128;
  constexpr int TB = D::TB;
  constexpr int NS = D::CPT / TB;
  static_assert(NS % 2 == 0, "double buffering");
  using Raw = VinRaw<L, FIRST, TP::DIN>;
  const float* Y_s = c.sm + SM::oY;
  const int* c_s = reinterpret_cast<const int*>(c.sm + SM::oC);
  const int e = c.m, uh = c.half;
  const float* gam = gsrc.row(c_s[e], D::F);
  float* Vng = WANT_V ? a.V[k + 1] + (size_t)tile * U * vpad(TP::DOUT) * TM : nullptr;
  const int q_lo = 2 * b, q_hi = D::lhi(b);
  struct In { Raw vin; float G[D::NSH]; };
  auto issue = [&](int s, In (&r)[TB]) {
#pragma unroll
    for (int bb = 0; bb < TB; ++bb) {
      const int u = uh * D::CPT + s * TB + bb;
      r[bb].vin.issue(a, tile, k, e, u);
#pragma unroll
      for (int lm = 0; lm < D::NSH; ++lm) r[bb].G[lm] = gam[lm * U + u];
    }
  };
  auto eval = [&](int s, const In (&r)[TB]) {
    float sq[TP::N0][TB];                               // scalar paths of the batch (4 consecutive channels -> one 128-bit store)
#pragma unroll
    for (int bb = 0; bb < TB; ++bb) {
      const int u = uh * D::CPT + s * TB + bb;
      float Vin[TP::DIN], Vout[TP::DOUT], sc[TP::N0];
      r[bb].vin.expand(e, Y_s, Vin);
      if (WANT_V) TP::template fwd<U>(Vin, r[bb].G, lw.omega + u, Vout, sc);
      else TPA::template fwd<U>(Vin, r[bb].G, nullptr, nullptr, sc);
#pragma unroll
      for (int q = 0; q < TP::N0; ++q) sq[q][bb] = sc[q];
      if (WANT_V) vec_store<TP::DOUT>(Vng + (size_t)u * vpad(TP::DOUT) * TM, e, Vout);
    }
    const int u0 = uh * D::CPT + s * TB;
#pragma unroll
    for (int q = 0; q < TP::N0; ++q) {
      if (q >= q_lo && q < q_hi) {
        if constexpr (TB == 4) op_put4<L>(c, (q - q_lo) * U + u0, sq[q][0], sq[q][1], sq[q][2], sq[q][3]);
        else {
#pragma unroll
          for (int bb = 0; bb < TB; ++bb) op_put1<L>(c, e, (q - q_lo) * U + u0 + bb, sq[q][bb]);
        }
      }
    }
  };
  constexpr bool DBUF = ALG_TP_FWD_DBUF_L3 || L < 3;
  In ra[TB], rb[DBUF ? TB : 1];
  if (DBUF) issue(0, ra);
  if (b == 0) {
#pragma unroll 1
    for (int i = DBUF ? TB : 0; i < D::CPT; ++i) Raw::prefetch(a, tile, k, e, uh * D::CPT + i);
  }
  if constexpr (DBUF) {
#pragma unroll 1
    for (int s = 0; s < NS; s += 2) {
      issue(s + 1, rb);
      eval(s, ra);
      if (s + 2 < NS) issue(s + 2, ra);
      eval(s + 1, rb);
    }
  } else {
#pragma unroll 1
    for (int s = 0; s < NS; ++s) {
      issue(s, ra);
      eval(s, ra);
    }
  }
}

// backward over all channels in passes of CHU, dG segmented sum -> dgamma_out.
// ds is read from DS_s ([q*U+u][128]); dG staged in the OPH/OPL regions.
template <int L, char KIND, bool FIRST, bool HAS_DVOUT>
__device__ ALG_NI void tc_tp_backward(const ChunkArgs& a, const LayerW& lw, const TcCtx& c, int tile, int k, int es, int nvalid,
                                               const float* dVnext, float* dVprev,
                                               float* dgamma_out, float* dYp, const RowSrc& gsrc) {
  using D = DimsTC<L>; using SM = SmemTC<L>; using TP = tpgen::TP<L, KIND>; constexpr int TM = 128;
  constexpr int TB = D::TB;
  constexpr int BPP = D::CHU / D::CPH / TB;           // batches per pass
  constexpr int NPASS = U / D::CHU;
  static_assert(D::CHU / D::CPH % TB == 0 && BPP % 2 == 0, "batching / double buffering");
  using Raw = VinRaw<L, FIRST, TP::DIN>;
  struct In { Raw vin; float dv[HAS_DVOUT ? TP::DOUT : 1]; };
  const float* DS_s = c.sm + SM::oDS;
  float* DG = c.sm + SM::oDG;
  const float* Y_s = c.sm + SM::oY;
  const int* c_s = reinterpret_cast<const int*>(c.sm + SM::oC);
  const int e = c.m, uh = c.half;
  const float* gam = gsrc.row(c_s[e], D::F);
  float* W0g = a.W0 + (size_t)tile * D::ENVW * TM;
  if (FIRST) {
#pragma unroll
    for (int lm = 0; lm < D::NSH; ++lm) dYp[lm] = 0.f;
  }
  constexpr int CPP = D::CHU / D::CPH;               // channels per pass and thread (contiguous block per half)
  auto chan = [&](int pass, int j) { return pass * D::CHU + uh * CPP + j; };
  auto issue = [&](int pass, int jb, In (&r)[TB]) {
#pragma unroll
    for (int bb = 0; bb < TB; ++bb) {
      const int u = chan(pass, jb * TB + bb);
      r[bb].vin.issue(a, tile, k, e, u);
      if (HAS_DVOUT) vec_load<TP::DOUT>(dVnext + ((size_t)tile * U + u) * vpad(TP::DOUT) * TM, e, r[bb].dv);
    }
  };
  auto eval = [&](int pass, int jb, const In (&r)[TB]) {
#pragma unroll
    for (int bb = 0; bb < TB; ++bb) {
      const int ul = uh * CPP + jb * TB + bb;
      const int u = pass * D::CHU + ul;
      float Vin[TP::DIN], G[D::NSH], ds[TP::N0], dVin[TP::DIN], dG[D::NSH];
      r[bb].vin.expand(e, Y_s, Vin);
#pragma unroll
      for (int lm = 0; lm < D::NSH; ++lm) G[lm] = gam[lm * U + u];
#pragma unroll
      for (int q = 0; q < TP::N0; ++q) ds[q] = DS_s[s_col(q, u) * TM + e];
      TP::template bwd<U>(Vin, G, lw.omega + u, r[bb].dv, ds, dVin, dG);
      if (FIRST) {
#pragma unroll
        for (int l = 0; l <= L; ++l) {
          const float wv = r[bb].vin.v[l];
          float dw = 0.f;
#pragma unroll
          for (int lm = l * l; lm < (l + 1) * (l + 1); ++lm) { dw += dVin[lm] * Y_s[lm * TM + e]; dYp[lm] += dVin[lm] * wv; }
          W0g[(l * U + u) * TM + e] = dw;      // in place: w0 -> dw0
        }
      } else {
        vec_store<TP::DIN>(dVprev + ((size_t)tile * U + u) * vpad(TP::DIN) * TM, e, dVin);
      }
#pragma unroll
      for (int lm = 0; lm < D::NSH; ++lm) DG[e * D::DGS + lm * D::CHU + ul] = dG[lm];
    }
  };
  // l_max = 3: one channel of a full-parity product already needs ~100 live registers (Vin, dVin, G, dG, dVout); a second
  // set of prefetched inputs pushes the arrays into local memory, so the loads are issued right before their evaluation
  // (the lines were pulled into L2 by the prefetch loop below)
  constexpr bool DBUF = ALG_TP_DBUF_L3 || L < ALG_TP_DBUF_BELOW;
  In ra[TB], rb[DBUF ? TB : 1];
  if (DBUF) issue(0, 0, ra);
#pragma unroll 1
  for (int i = DBUF ? TB : 0; i < U / D::CPH; ++i) {    // all remaining channels of this thread -> L2
    const int u = chan(i / CPP, i % CPP);
    Raw::prefetch(a, tile, k, e, u);
    if (HAS_DVOUT) {
      vec_prefetch<TP::DOUT>(dVnext + ((size_t)tile * U + u) * vpad(TP::DOUT) * TM, e);
    }
  }
#pragma unroll 1
  for (int pass = 0; pass < NPASS; ++pass) {
    if constexpr (DBUF) {
#pragma unroll 1
      for (int jb = 0; jb < BPP; jb += 2) {
        issue(pass, jb + 1, rb);
        eval(pass, jb, ra);
        if (jb + 2 < BPP) issue(pass, jb + 2, ra);
        else if (pass + 1 < NPASS) issue(pass + 1, 0, ra);
        eval(pass, jb + 1, rb);
      }
    } else {
#pragma unroll 1
      for (int jb = 0; jb < BPP; ++jb) {
        issue(pass, jb, ra);
        if (ALG_TP_L1_PREFETCH && (jb + 1 < BPP || pass + 1 < NPASS)) {      // next channel: L2 -> L1 while this one is evaluated
          const int un = jb + 1 < BPP ? chan(pass, (jb + 1) * TB) : chan(pass + 1, 0);
#pragma unroll
          for (int bb = 0; bb < TB; ++bb) {
            Raw::template prefetch<true>(a, tile, k, e, un + bb);
            if (HAS_DVOUT) vec_prefetch<TP::DOUT, true>(dVnext + ((size_t)tile * U + un + bb) * vpad(TP::DOUT) * TM, e);
          }
        }
        eval(pass, jb, ra);
      }
    }
    __syncthreads();
    {
      const int* seg = reinterpret_cast<const int*>(c.sm + SM::oSEG);
      const bool contin = a.rowptr[c_s[0]] < es;
      float* carry = a.carry + (size_t)tile * D::F;
      segsum_items<D::FC, TM>(c_s, seg,
          [&](int ee, int f) { return DG[ee * D::DGS + f]; },
          [&](int centre, int f, bool first, float v) {
            const int lm = f / D::CHU, ul = f % D::CHU;
            const int fg = lm * U + pass * D::CHU + ul;
            if (first && contin) carry[fg] = v; else dgamma_out[(size_t)(centre - c.c0) * D::F + fg] = v;
          });
    }
    __syncthreads();
  }
  (void)nvalid;
}

// dY: DY_s (phase-2 part, smem) + the two channel-halves' partials (FIRST layers) -> global dY
template <int L, bool ASSIGN>
__device__ ALG_NI void tc_dy_store(const ChunkArgs& a, const TcCtx& c, int tile, const float* dYp, bool have) {
  using D = DimsTC<L>; using SM = SmemTC<L>; constexpr int TM = 128;
  float* P = c.sm + SM::oOPH;     // [2][NSH][128]
  if (have) {
#pragma unroll
    for (int lm = 0; lm < D::NSH; ++lm) P[(c.half * D::NSH + lm) * TM + c.m] = dYp[lm];
  }
  __syncthreads();
  float* dYg = a.dY + (size_t)tile * D::NSH * TM;
  const float* DY_s = c.sm + SM::oDY;
  for (int i = threadIdx.x; i < D::NSH * TM; i += NT) {
    float v = DY_s[i];
    if (have) v += P[i] + P[D::NSH * TM + i];
    if (ASSIGN) dYg[i] = v; else dYg[i] += v;
  }
  __syncthreads();
}

// every GEMM of a chain accumulates into the same TMEM block: its epilogue has consumed the previous result
// (and the next MMA is issued behind a CTA barrier) before it is overwritten
constexpr uint32_t TC_Z1 = TC_ACC, TC_Z2 = TC_ACC, TC_M = TC_ACC, TC_SCR = TC_ACC;
struct NoBias { __device__ __forceinline__ float operator()(int) const { return 0.f; } };

// activation record of one MLP evaluation kept for the backward kernels (tile-SoA, [3][64][128] per tile):
// rows [0,64) = act'(z1 + bias), [64,128) = act'(z2), [128,192) = m (pre-envelope output)
constexpr int ZD_ROWS = 3 * 64;
// hidden layers of an MLP whose layer-0 pre-activation z1 is complete in TMEM (w1 requested):
// leaves z2 and m (pre-envelope output) in TMEM, requests `next`.  With STORE the activation
// derivatives are written to `zd` so that the backward kernels need no recomputation.
template <int L, bool STORE, class Bias>
__device__ ALG_NI void tc_mlp_hidden_fwd(TcCtx& c, const TcMat& w2, const TcMat& next, Bias bias, float* zd = nullptr) {
  constexpr int TM = 128;
  tc_epi(c, TC_Z1, 64, [&](int n, float v0, float v1, float v2, float v3) {
    if constexpr (STORE) {
      float2 r01, r23, d01, d23;
      silu_act2(make_float2(v0 + bias(n), v1 + bias(n + 1)), r01, d01);
      silu_act2(make_float2(v2 + bias(n + 2), v3 + bias(n + 3)), r23, d23);
      op_put4<L>(c, n, r01.x, r01.y, r23.x, r23.y);
      st_row4(zd, n, c.m, d01.x, d01.y, d23.x, d23.y);
    } else {
      const float2 r01 = silu_act2(make_float2(v0 + bias(n), v1 + bias(n + 1))), r23 = silu_act2(make_float2(v2 + bias(n + 2), v3 + bias(n + 3)));
      op_put4<L>(c, n, r01.x, r01.y, r23.x, r23.y);
    }
  });
  tc_mma<L>(c, 64, 64, TC_Z2);
  tc_load_w<L>(c, w2);
  tc_epi(c, TC_Z2, 64, [&](int n, float v0, float v1, float v2, float v3) {
    if constexpr (STORE) {
      float2 r01, r23, d01, d23;
      silu_act2(make_float2(v0, v1), r01, d01);
      silu_act2(make_float2(v2, v3), r23, d23);
      op_put4<L>(c, n, r01.x, r01.y, r23.x, r23.y);
      st_row4(zd + 64 * TM, n, c.m, d01.x, d01.y, d23.x, d23.y);
    } else {
      const float2 r01 = silu_act2(make_float2(v0, v1)), r23 = silu_act2(make_float2(v2, v3));
      op_put4<L>(c, n, r01.x, r01.y, r23.x, r23.y);
    }
  });
  tc_mma<L>(c, 64, 64, TC_M);
  tc_load_w<L>(c, next);
}

// backward through the hidden layers with the STORED derivatives: dm in operand [0,64), w2_b requested:
// dz2 = (dm W2^T) act'(z2); dz1 = (dz2 W1^T) act'(z1) -> operand; requests `next`
// CG: the record was written by this very thread earlier in the same kernel (k_t_tc): read it with ld.global.cg
// next2 (optional): weights of a GEMM that shares the operand with `next` -> weight buffer 2 (tc_mma_pair)
template <int L, bool CG = false>
__device__ ALG_NI void tc_mlp_bwd_hidden_st(TcCtx& c, const TcMat& w1_b, const TcMat& next, const float* zd, const TcMat* next2 = nullptr) {
  constexpr int TM = 128;
  float d[32];
  ld_rows32<CG>(c, zd + 64 * TM, d);                  // act'(z2): in flight during the MMA
  tc_mma<L>(c, 64, 64, TC_SCR);
  tc_load_w<L>(c, w1_b);
  tc_epi(c, TC_SCR, 64, [&](int n, float v0, float v1, float v2, float v3) {
    const int j = n - c.half * 32;
    op_put4<L>(c, n, v0 * d[j], v1 * d[j + 1], v2 * d[j + 2], v3 * d[j + 3]);
  });
  ld_rows32<CG>(c, zd, d);                            // act'(z1 + bias)
  tc_mma<L>(c, 64, 64, TC_SCR);
  tc_load_w<L>(c, next);
  if (next2) tc_load_w2<L>(c, *next2);
  tc_epi(c, TC_SCR, 64, [&](int n, float v0, float v1, float v2, float v3) {
    const int j = n - c.half * 32;
    op_put4<L>(c, n, v0 * d[j], v1 * d[j + 1], v2 * d[j + 2], v3 * d[j + 3]);
  });
}

// dz1 in operand, m0_bx requested: dX(global) += dz1 W0x^T ; ds = dz1 W0s^T -> DS_s.
// (l_max = 2: the first ds block is held in registers until the MMA that still reads the weight region DS_s aliases
// has completed.)
template <int L>
__device__ ALG_NI void tc_din(TcCtx& c, const TcLayerW& tl, float* dXg) {
  using D = DimsTC<L>; using SM = SmemTC<L>; constexpr int TM = 128;
  float dp[32];
  ld_rows32(c, dXg, dp);                              // in flight during the MMAs
  // dz1 W0x^T -> TC_ACC and dz1 W0s[block 0]^T -> TC_ACC2 as one MMA group (m0_bs[0] was requested into weight
  // buffer 2 by the caller's tc_load_w2 -- see tc_mlp_bwd_hidden_st)
  tc_mma_pair<L>(c, 64, 64, TC_ACC, 64, TC_ACC2);
  if (D::NB > 1) tc_load_w<L>(c, tl.m0_bs[1]);
  tc_epi(c, TC_ACC, 64, [&](int n, float v0, float v1, float v2, float v3) {
    const int j = n - c.half * 32;
    st_row4(dXg, n, c.m, dp[j] + v0, dp[j + 1] + v1, dp[j + 2] + v2, dp[j + 3] + v3);
  });
  float* DS_s = c.sm + SM::oDS;
  if constexpr (D::NB == 1) {
    // DS_s aliases only weight buffer 1, dead once the MMA group has completed
    tc_epi(c, TC_ACC2, 64, [&](int n, float v0, float v1, float v2, float v3) {
      float* p = DS_s + n * TM + c.m;
      p[0] = v0; p[TM] = v1; p[2 * TM] = v2; p[3 * TM] = v3;
    });
  } else {
    float dsr[32];                                    // block 0 is held in registers: DS_s aliases the weights of block 1's MMA
    tc_epi(c, TC_ACC2, 64, [&](int n, float v0, float v1, float v2, float v3) {
      const int j = n - c.half * 32;
      dsr[j] = v0; dsr[j + 1] = v1; dsr[j + 2] = v2; dsr[j + 3] = v3;
    });
    tc_mma<L>(c, 64, D::bw(1), TC_ACC);               // same operand (dz1); weights are dead afterwards
#pragma unroll
    for (int j = 0; j < 32; ++j) DS_s[(c.half * 32 + j) * TM + c.m] = dsr[j];
    tc_epi_bw(c, TC_ACC, D::bw(1), [&](int n, float v0, float v1, float v2, float v3) {
      float* p = DS_s + (64 + n) * TM + c.m;
      p[0] = v0; p[TM] = v1; p[2 * TM] = v2; p[3 * TM] = v3;
    });
  }
  __syncthreads();
}

// phase 2 of layer kk.  In: x^kk in operand [0,64), weight block env[0] requested, dsrc = dGamma_kk rows.
// For every block b: w_b = x env_b (TMEM) -> dw_b (operand), dY partial; then dxacc += dw_b env_b^T.
// With emb_b != nullptr (layer 0 in B0) the embed backward dw0 emb^T is accumulated as well.
// Out: dxacc[32] = this thread's 32 columns of the correction to dx^kk, DY_s = d/dY of the env sum.
// Requests `next` before the last epilogue.
template <int L, class Pre>
__device__ __forceinline__ void tc_phase2(const ChunkArgs& a, const ModelW& w, TcCtx& c, const TcLayerW& tl, const TcMat* emb_b, int tile,
                                          const float* Xtile, const TcMat& next, const RowSrc& dsrc, float* dxacc, Pre pre) {
  using D = DimsTC<L>; using SM = SmemTC<L>; constexpr int TM = 128;
  const float* Y_s = c.sm + SM::oY;
  const int* c_s = reinterpret_cast<const int*>(c.sm + SM::oC);
  float* DY_s = c.sm + SM::oDY;
  const float* dgam = dsrc.row(c_s[c.m], D::F);
  float dYp[D::NSH];
#pragma unroll
  for (int lm = 0; lm < D::NSH; ++lm) dYp[lm] = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) dxacc[i] = 0.f;
  // w of ALL blocks must be computed from x before the operand is overwritten by dw: TMEM scratch holds
  // one block, so x is re-staged from global for every further block (NB <= 2)
#pragma unroll 1
  for (int b = 0; b < D::NB; ++b) {
    tc_mma<L>(c, 64, D::bw(b), TC_SCR);                // w_b = x env_b
    tc_load_w<L>(c, tl.env_b[b]);
#pragma unroll
    for (int l = 2 * b; l < D::lhi(b); ++l) {
      float wv[16], dw[16];
      tc_ld16(c, TC_SCR + (l - 2 * b) * U + c.half * 16, wv);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int u = c.half * 16 + i;
        float acc = 0.f;
#pragma unroll
        for (int lm = l * l; lm < (l + 1) * (l + 1); ++lm) {
          const float dg = dgam[lm * U + u] * w.inv_sqrt_n;
          acc += dg * Y_s[lm * TM + c.m];
          dYp[lm] += dg * wv[i];
        }
        dw[i] = acc;
      }
#pragma unroll
      for (int i = 0; i < 16; i += 4) op_put4<L>(c, (l - 2 * b) * U + c.half * 16 + i, dw[i], dw[i + 1], dw[i + 2], dw[i + 3]);
    }
    if (b + 1 == D::NB && !emb_b) pre();               // caller's global loads fly during the last MMA
    tc_mma<L>(c, D::bw(b), 64, TC_SCR);                // dw_b env_b^T
    if (b + 1 < D::NB) tc_load_w<L>(c, tl.env[b + 1]);
    else if (emb_b) tc_load_w<L>(c, emb_b[0]);
    else tc_load_w<L>(c, next);
    if (b + 1 < D::NB) {                                // scratch is needed for the next block's w: drain to registers
      tc_epi(c, TC_SCR, 64, [&](int n, float v0, float v1, float v2, float v3) {
        const int j = n - c.half * 32;
        dxacc[j] += v0; dxacc[j + 1] += v1; dxacc[j + 2] += v2; dxacc[j + 3] += v3;
      });
      op_load_rows64<L>(c, Xtile);                      // x again for the next block's w
    }
  }
  if (emb_b) {                                          // embed backward accumulates onto the last block's product in TMEM
#pragma unroll 1
    for (int b = 0; b < D::NB; ++b) {
      op_load_rows_bw<L>(c, a.W0 + ((size_t)tile * D::ENVW + 64 * b) * TM, D::bw(b));     // dw0 block
      if (b + 1 == D::NB) pre();
      tc_mma<L>(c, D::bw(b), 64, TC_SCR, 1);
      if (b + 1 < D::NB) tc_load_w<L>(c, emb_b[b + 1]); else tc_load_w<L>(c, next);
    }
  }
  tc_epi(c, TC_SCR, 64, [&](int n, float v0, float v1, float v2, float v3) {
    const int j = n - c.half * 32;
    dxacc[j] += v0; dxacc[j + 1] += v1; dxacc[j + 2] += v2; dxacc[j + 3] += v3;
  });
  if (c.half == 1) {
#pragma unroll
    for (int lm = 0; lm < D::NSH; ++lm) DY_s[lm * TM + c.m] = dYp[lm];
  }
  __syncthreads();
  if (c.half == 0) {
#pragma unroll
    for (int lm = 0; lm < D::NSH; ++lm) DY_s[lm * TM + c.m] += dYp[lm];
  }
}

// ============================================================================================
// F0
// ============================================================================================
// L2 prefetch of the per-tile state a phase reads (issued before the geometry: the phases are latency-bound)
template <int L> __device__ __forceinline__ void fk_prefetch(const ChunkArgs& a, const TcCtx* c, int tile, int k, const GeomIn& gi, int c0, size_t goff) {
  tc_prefetch(a.X[k] + (size_t)tile * S * 128, S * 128);
  tc_prefetch_row<L>(a.gamma[k] + goff, c0, gi);
  (void)c;
}
template <int L> __device__ __forceinline__ void bk_prefetch(const ChunkArgs& a, int tile, int k, const GeomIn& gi, int c0, size_t goff) {
  using D = DimsTC<L>; constexpr int TM = 128;
  tc_prefetch(a.X[k + 1] + (size_t)tile * S * TM, S * TM);
  tc_prefetch_row<L>(a.dgamma[k + 1] + goff, c0, gi);
  tc_prefetch_row<L>(a.gamma[k] + goff, c0, gi);
  tc_prefetch(a.dX + (size_t)tile * S * TM, S * TM);
  tc_prefetch(a.ZD[k + 1] + (size_t)tile * ZD_ROWS * TM, ZD_ROWS * TM);
  tc_prefetch(a.dY + (size_t)tile * D::NSH * TM, D::NSH * TM);
  tc_prefetch(a.du + (size_t)tile * TM, TM);
}
template <int L> __device__ __forceinline__ void b0_prefetch(const ChunkArgs& a, int tile, const GeomIn& gi, int c0, size_t goff) {
  using D = DimsTC<L>; constexpr int TM = 128;
  tc_prefetch(a.X[0] + (size_t)tile * S * TM, S * TM);
  tc_prefetch_row<L>(a.dgamma[0] + goff, c0, gi);
  tc_prefetch(a.dX + (size_t)tile * S * TM, S * TM);
  tc_prefetch(a.ZD[0] + (size_t)tile * ZD_ROWS * TM, ZD_ROWS * TM);
  tc_prefetch(a.W0 + (size_t)tile * D::ENVW * TM, D::ENVW * TM);
  tc_prefetch(a.dY + (size_t)tile * D::NSH * TM, D::NSH * TM);
  tc_prefetch(a.du + (size_t)tile * TM, TM);
}

// Every phase is a device function ("body") that starts from the published tile geometry (Y_s, u_s, c_s, zz_s, segment
// table; see tc_tile_begin) and the per-thread Geom: the chunked pipeline wraps each body in its own kernel (tile =
// blockIdx.x), the fused persistent kernel (k_fused_tc) runs all bodies of a tile back to back in one CTA.
// `tile` indexes the per-tile scratch arrays (chunk tile number, or the CTA slot in the fused kernel).
template <int L>
__device__ __forceinline__ void f0_body(const ChunkArgs& a, const ModelW& w, const TcW& tw, TcCtx& c, const Geom& g, int tile, int es, int nvalid) {
  using D = DimsTC<L>; using SM = SmemTC<L>; constexpr int TM = 128;
  tc_load_w<L>(c, tw.two0);
  {  // Bessel*u -> operand columns [0,32) (zero padded beyond num_bessels)
    // sin((n+1) theta) by the Chebyshev recurrence s_{n+1} = 2 cos(theta) s_n - s_{n-1}: one sincosf per edge
    if (c.half == 0) {
      float s1, c1;
      sincosf(3.14159265358979323846f * (g.r / g.rc), &s1, &c1);
      const float sc = sqrtf(2.0f / g.rc) / g.r * g.u, c2 = 2.0f * c1;
      float sp = 0.f, sn = s1;
#pragma unroll
      for (int k4 = 0; k4 < MAXB; k4 += 4) {
        float b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          b[i] = (k4 + i) < w.B ? sn * sc : 0.f;
          const float nx = c2 * sn - sp;
          sp = sn; sn = nx;
        }
        op_put4<L>(c, k4, b[0], b[1], b[2], b[3]);
      }
    } else {
#pragma unroll
      for (int k4 = 16; k4 < 32; k4 += 4) op_put4<L>(c, k4, 0.f, 0.f, 0.f, 0.f);
    }
  }
  tc_mma<L>(c, 32, 64, TC_Z1);
  tc_load_w<L>(c, tw.two1);
  const float* w0 = w.two.w[0];
  const float* wi = w0 + g.zi * H; const float* wj = w0 + (w.T + g.zj) * H;
  float* zd = a.ZD[0] + (size_t)tile * ZD_ROWS * TM;
  tc_mlp_hidden_fwd<L, true>(c, tw.two2, tw.emb[0], [&](int n) { return __ldg(wi + n) + __ldg(wj + n); }, zd);
  tc_load_w2<L>(c, D::NB == 1 ? tw.layer[0].env[0] : tw.emb[1]);      // partner of the emb[0] GEMM (same operand x^0)
  {
    float* X0g = a.X[0] + (size_t)tile * S * TM;
    tc_epi(c, TC_M, 64, [&](int n, float v0, float v1, float v2, float v3) {
      st_row4(zd + 128 * TM, n, c.m, v0, v1, v2, v3);
      v0 *= g.u; v1 *= g.u; v2 *= g.u; v3 *= g.u;
      op_put4<L>(c, n, v0, v1, v2, v3);
      st_row4(X0g, n, c.m, v0, v1, v2, v3);
    });
  }
  {  // embed linear -> w0 in HBM/L2, env linear of layer 0 -> Gamma_0: both read the operand x^0
    float* W0g = a.W0 + (size_t)tile * D::ENVW * TM;
    auto w0_store = [&](int b, uint32_t col) {
      tc_epi_bw(c, col, D::bw(b), [&](int n, float v0, float v1, float v2, float v3) {
        float* p = W0g + (size_t)(64 * b + n) * TM + c.m;
        p[0] = v0; p[TM] = v1; p[2 * TM] = v2; p[3 * TM] = v3;
      });
    };
    if constexpr (D::NB == 1) {
      // l_max = 1: emb (buffer 1) and env_0 (buffer 2) as one MMA group
      tc_mma_pair<L>(c, 64, D::bw(0), TC_ACC, D::bw(0), TC_ACC2);
      w0_store(0, TC_ACC);
      tc_env_finish<L>(a, w, c, tile, es, (a.gamma[0] + c.goff), TC_ACC2, TC_ACC2);
    } else {
      // l_max = 2: the two emb blocks as one group, then the two env blocks
      tc_mma_pair<L>(c, 64, D::bw(0), TC_ACC, D::bw(1), TC_ACC2);
      tc_load_w<L>(c, tw.layer[0].env[0]);
      tc_load_w2<L>(c, tw.layer[0].env[1]);
      w0_store(0, TC_ACC);
      w0_store(1, TC_ACC2);
      tc_env_all<L>(a, w, c, tw.layer[0].env, tile, es, (a.gamma[0] + c.goff));
    }
  }
  (void)nvalid;
}
// geometry of the tile -> shared memory (Y_s, u_s, c_s, zz_s) + segment table; returns this thread's Geom
template <int L>
__device__ __forceinline__ Geom tc_tile_begin(const ChunkArgs& a, const ModelW& w, const TcCtx& c, const GeomIn& gi, int nvalid) {
  using SM = SmemTC<L>;
  const Geom g = tc_geom<L>(a, w, c, gi);
  __syncthreads();
  seg_build<128>(reinterpret_cast<const int*>(c.sm + SM::oC), nvalid, reinterpret_cast<int*>(c.sm + SM::oSEG));
  return g;
}
template <int L>
__global__ void __launch_bounds__(NT, DimsTC<L>::MINB) k_f0_tc(const __grid_constant__ ChunkArgs a, const __grid_constant__ ModelW w, const __grid_constant__ TcW tw) {
  using D = DimsTC<L>; constexpr int TM = 128;
  extern __shared__ __align__(1024) float sm_raw[];
  // persistent over the tiles of the chunk: tile = blockIdx.x, blockIdx.x + gridDim.x, ... (TMEM / mbarriers set up once per CTA)
  const ChunkBounds cb = chunk_bounds(a);
  const int ntiles = (cb.e1 - cb.e0 + TM - 1) / TM;
  if ((int)blockIdx.x >= ntiles) return;               // also: the empty chunks behind the last edge of a device-built plan
  TcCtx c = tc_begin<L>(sm_raw, tw);
  c.c0 = cb.c0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int es = cb.e0 + tile * TM;
    const int nvalid = min(TM, cb.e1 - es);
    const GeomIn gi = tc_geom_load(a, es, nvalid);

    const Geom g = tc_tile_begin<L>(a, w, c, gi, nvalid);
    f0_body<L>(a, w, tw, c, g, tile, es, nvalid);
    __syncthreads();
  }
  tc_end(c);
  (void)sizeof(D);
}

// layer-0 GEMM of a latent MLP: z1 = [x || s] W0 as accumulating K-blocks (s blocks first, then x)
// in: weight block m0s[0] requested, geometry published.  out: z1 in TMEM, m1 requested.
template <int L, char KIND, bool FIRST, bool WANT_V>
__device__ __forceinline__ void tc_latent_z1(const ChunkArgs& a, const LayerW& lw, const TcLayerW& tl, TcCtx& c, int tile, int k,
                                             const float* Xg, const RowSrc& gsrc) {
  using D = DimsTC<L>;
  float4 xv[8];                                       // x^k rows: loads fly during the last s-block MMA
  tc_tp_forward<L, KIND, FIRST, WANT_V>(a, lw, c, tile, k, gsrc, 0);
  if (D::NB == 1) ld_rows_x8(Xg, xv);
  tc_mma<L>(c, D::bw(0), 64, TC_Z1, 0);
  if (D::NB > 1) {
    tc_load_w<L>(c, tl.m0s[1]);
    tc_tp_forward<L, KIND, FIRST, false>(a, lw, c, tile, k, gsrc, 1);
    ld_rows_x8(Xg, xv);
    tc_mma<L>(c, D::bw(D::NB - 1), 64, TC_Z1, 1);
  }
  tc_load_w<L>(c, tl.m0x);
  op_put_x8<L>(c, xv);
  tc_mma<L>(c, 64, 64, TC_Z1, 1);
  tc_load_w<L>(c, tl.m1);
}

// ============================================================================================
// FK
// ============================================================================================
template <int L, char KIND, bool FIRST>
__device__ __forceinline__ void fk_body(const ChunkArgs& a, const ModelW& w, const TcW& tw, TcCtx& c, const Geom& g, int tile, int es, int nvalid, const int k) {
  using D = DimsTC<L>; constexpr int TM = 128;
  const LayerW& lw = w.layer[k];
  const TcLayerW& tl = tw.layer[k];
  tc_load_w<L>(c, tl.m0s[0]);
  const float* Xg = a.X[k] + (size_t)tile * S * TM;
  const RowSrc gsrc = tc_stage_rows<L>(c, (a.gamma[k] + c.goff), c.c0, nvalid);
  tc_latent_z1<L, KIND, FIRST, true>(a, lw, tl, c, tile, k, Xg, gsrc);
  float* zd = a.ZD[k + 1] + (size_t)tile * ZD_ROWS * TM;
  tc_mlp_hidden_fwd<L, true>(c, tl.m2, tw.layer[k + 1].env[0], NoBias(), zd);
  if (D::NB > 1) tc_load_w2<L>(c, tw.layer[k + 1].env[1]);
  {
    float* Xng = a.X[k + 1] + (size_t)tile * S * TM;
    float xp[32];
    ld_rows32(c, Xg, xp);
    tc_epi(c, TC_M, 64, [&](int n, float v0, float v1, float v2, float v3) {
      const int j = n - c.half * 32;
      st_row4(zd + 128 * TM, n, c.m, v0, v1, v2, v3);
      const float x0 = lw.a * xp[j] + lw.b * v0 * g.u, x1 = lw.a * xp[j + 1] + lw.b * v1 * g.u;
      const float x2 = lw.a * xp[j + 2] + lw.b * v2 * g.u, x3 = lw.a * xp[j + 3] + lw.b * v3 * g.u;
      op_put4<L>(c, n, x0, x1, x2, x3);
      st_row4(Xng, n, c.m, x0, x1, x2, x3);
    });
  }
  tc_env_all<L>(a, w, c, tw.layer[k + 1].env, tile, es, (a.gamma[k + 1] + c.goff));
}
template <int L, char KIND, bool FIRST>
__global__ void __launch_bounds__(NT, DimsTC<L>::MINB) k_fk_tc(const __grid_constant__ ChunkArgs a, const __grid_constant__ ModelW w, const __grid_constant__ TcW tw, const int k) {
  using D = DimsTC<L>; constexpr int TM = 128;
  extern __shared__ __align__(1024) float sm_raw[];
  // persistent over the tiles of the chunk: tile = blockIdx.x, blockIdx.x + gridDim.x, ... (TMEM / mbarriers set up once per CTA)
  const ChunkBounds cb = chunk_bounds(a);
  const int ntiles = (cb.e1 - cb.e0 + TM - 1) / TM;
  if ((int)blockIdx.x >= ntiles) return;               // also: the empty chunks behind the last edge of a device-built plan
  TcCtx c = tc_begin<L>(sm_raw, tw);
  c.c0 = cb.c0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int es = cb.e0 + tile * TM;
    const int nvalid = min(TM, cb.e1 - es);
    const GeomIn gi = tc_geom_load(a, es, nvalid);
    fk_prefetch<L>(a, &c, tile, k, gi, cb.c0, 0);
    const Geom g = tc_tile_begin<L>(a, w, c, gi, nvalid);
    fk_body<L, KIND, FIRST>(a, w, tw, c, g, tile, es, nvalid, k);
    __syncthreads();
  }
  tc_end(c);
  (void)sizeof(D);
}

// ============================================================================================
// T
// ============================================================================================
template <int L, bool FIRST>
__device__ __forceinline__ void t_body(const ChunkArgs& a, const ModelW& w, const TcW& tw, TcCtx& c, const Geom& g, int tile, int es, int nvalid, const int k) {
  using D = DimsTC<L>; using SM = SmemTC<L>; constexpr int TM = 128;
  const LayerW& lw = w.layer[k];
  const TcLayerW& tl = tw.layer[k];
  tc_load_w<L>(c, tl.m0s[0]);
  const float* Xg = a.X[k] + (size_t)tile * S * TM;
  const int* c_s = reinterpret_cast<const int*>(c.sm + SM::oC);
  const RowSrc gsrc = tc_stage_rows<L>(c, (a.gamma[k] + c.goff), c.c0, nvalid);
  tc_latent_z1<L, 'A', FIRST, false>(a, lw, tl, c, tile, k, Xg, gsrc);
  // act'(z1), act'(z2), m of this layer: written and read back by the same thread (the accumulators and the
  // A operand occupy all of this CTA's tensor memory)
  float* zd = a.ZD[k + 1] + (size_t)tile * ZD_ROWS * TM;
  tc_mlp_hidden_fwd<L, true>(c, tl.m2, tw.ro0, NoBias(), zd);
  {
    float xp[32];
    ld_rows32(c, Xg, xp);
    tc_epi(c, TC_M, 64, [&](int n, float v0, float v1, float v2, float v3) {        // x^n -> operand
      const int j = n - c.half * 32;
      st_row4(zd + 128 * TM, n, c.m, v0, v1, v2, v3);
      op_put4<L>(c, n, lw.a * xp[j] + lw.b * v0 * g.u, lw.a * xp[j + 1] + lw.b * v1 * g.u,
                 lw.a * xp[j + 2] + lw.b * v2 * g.u, lw.a * xp[j + 3] + lw.b * v3 * g.u);
    });
  }
  tc_mma<L>(c, 64, R, TC_SCR);                       // readout hidden
  tc_load_w<L>(c, tw.ro0_b);
  float* e_s = c.sm + SM::oE;
  {
    const float ge = w.gscale[g.zi];
    float v[16];
    tc_ld16(c, TC_SCR + c.half * 16, v);
    float ee = 0.f, dz[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float wq = __ldg(w.ro1 + c.half * 16 + i);
      float d;
      const float r = silu_act(v[i], d);
      ee += wq * r;
      dz[i] = ge * wq * d;
    }
#pragma unroll
    for (int i = 0; i < 16; i += 4) op_put4<L>(c, c.half * 16 + i, dz[i], dz[i + 1], dz[i + 2], dz[i + 3]);
    e_s[c.half * TM + c.m] = ee;
  }
  float mv[32];
  ld_rows32<true>(c, zd + 128 * TM, mv);             // m: in flight during the MMA
  tc_mma<L>(c, R, 64, TC_SCR);                       // dx^n = dz ro0^T   (the barrier inside orders e_s)
  tc_load_w<L>(c, tl.m2_b);
  if (c.half == 0) {
    const float ee = e_s[c.m] + e_s[TM + c.m];
    e_s[2 * TM + c.m] = ee;                          // final E_e (read after the next barrier)
    if (c.m < nvalid && a.edge_energy) a.edge_energy[es + c.m] = ee;
  }
  float* dXg = a.dX + (size_t)tile * S * TM;
  {
    float dup = 0.f;
    tc_epi(c, TC_SCR, 64, [&](int n, float v0, float v1, float v2, float v3) {
      const int j = n - c.half * 32;
      float v[4] = {v0, v1, v2, v3};
      st_row4(dXg, n, c.m, lw.a * v0, lw.a * v1, lw.a * v2, lw.a * v3);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        v[q] *= lw.b;                 // dxt
        dup += v[q] * mv[j + q];
        v[q] *= g.u;                  // dm
      }
      op_put4<L>(c, n, v[0], v[1], v[2], v[3]);
    });
    if (c.half == 1) e_s[3 * TM + c.m] = dup;
    __syncthreads();
    if (c.half == 0) a.du[(size_t)tile * TM + c.m] = dup + e_s[3 * TM + c.m];
  }
  {  // E_i raw sums (double): one thread per centre run, edges in order (deterministic)
    const int* seg = reinterpret_cast<const int*>(c.sm + SM::oSEG);
    const int nseg = seg[TM + 1];
    const bool contin = a.rowptr[c_s[0]] < es;
    for (int sgm = NT - 1 - threadIdx.x; sgm < nseg; sgm += NT) {     // highest threads first: not the MMA-issuing thread
      double acc = 0.0;
      for (int e = seg[sgm]; e < seg[sgm + 1]; ++e) acc += (double)e_s[2 * TM + e];
      if (sgm == 0 && contin) a.ecarry[tile] = acc; else a.esum[c_s[seg[sgm]]] = acc;
    }
  }
  tc_mlp_bwd_hidden_st<L, true>(c, tl.m1_b, tl.m0_bx, zd, &tl.m0_bs[0]);
  tc_din<L>(c, tl, dXg);
  float dYp[D::NSH];
  tc_tp_backward<L, 'A', FIRST, false>(a, lw, c, tile, k, es, nvalid, nullptr, FIRST ? nullptr : a.dV[k & 1], (a.dgamma[k] + c.goff), dYp, gsrc);
  for (int i = threadIdx.x; i < D::NSH * TM; i += NT) c.sm[SM::oDY + i] = 0.f;
  __syncthreads();
  tc_dy_store<L, true>(a, c, tile, dYp, FIRST);
}
template <int L, bool FIRST>
__global__ void __launch_bounds__(NT, DimsTC<L>::MINB) k_t_tc(const __grid_constant__ ChunkArgs a, const __grid_constant__ ModelW w, const __grid_constant__ TcW tw, const int k) {
  using D = DimsTC<L>; constexpr int TM = 128;
  extern __shared__ __align__(1024) float sm_raw[];
  // persistent over the tiles of the chunk: tile = blockIdx.x, blockIdx.x + gridDim.x, ... (TMEM / mbarriers set up once per CTA)
  const ChunkBounds cb = chunk_bounds(a);
  const int ntiles = (cb.e1 - cb.e0 + TM - 1) / TM;
  if ((int)blockIdx.x >= ntiles) return;               // also: the empty chunks behind the last edge of a device-built plan
  TcCtx c = tc_begin<L>(sm_raw, tw);
  c.c0 = cb.c0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int es = cb.e0 + tile * TM;
    const int nvalid = min(TM, cb.e1 - es);
    const GeomIn gi = tc_geom_load(a, es, nvalid);
    fk_prefetch<L>(a, &c, tile, k, gi, cb.c0, 0);
    const Geom g = tc_tile_begin<L>(a, w, c, gi, nvalid);
    t_body<L, FIRST>(a, w, tw, c, g, tile, es, nvalid, k);
    __syncthreads();
  }
  tc_end(c);
  (void)sizeof(D);
}

// ============================================================================================
// BK
// ============================================================================================
template <int L, char KIND, bool FIRST>
__device__ __forceinline__ void bk_body(const ChunkArgs& a, const ModelW& w, const TcW& tw, TcCtx& c, const Geom& g, int tile, int es, int nvalid, const int k) {
  using D = DimsTC<L>; using SM = SmemTC<L>; constexpr int TM = 128;
  const float* Xn = a.X[k + 1] + (size_t)tile * S * TM;
  const LayerW& lw = w.layer[k];
  const TcLayerW& tl = tw.layer[k];
  tc_load_w<L>(c, tw.layer[k + 1].env[0]);
  op_load_rows64<L>(c, Xn);
  float* dXg = a.dX + (size_t)tile * S * TM;
  const RowSrc dsrc = tc_stage_rows<L>(c, (a.dgamma[k + 1] + c.goff), c.c0, nvalid);
  float dxn[32];                                      // complete dx^{k+1} of this thread's 32 columns (kept in registers)
  const float* zd = a.ZD[k + 1] + (size_t)tile * ZD_ROWS * TM;      // act'(z1), act'(z2), m of layer k (written by FK)
  {
    float dp[32], mv[32];
    tc_phase2<L>(a, w, c, tw.layer[k + 1], nullptr, tile, Xn, tl.m2_b, dsrc, dxn,
                 [&] { ld_rows32(c, dXg, dp); ld_rows32(c, zd + 128 * TM, mv); });
    float dup = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float v[4], dxa[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float dx1 = dxn[i + q] + dp[i + q];
        dxa[q] = lw.a * dx1;
        const float dxt = lw.b * dx1;
        dup += dxt * mv[i + q];
        v[q] = dxt * g.u;
      }
      st_row4(dXg, c.half * 32 + i, c.m, dxa[0], dxa[1], dxa[2], dxa[3]);
      op_put4<L>(c, c.half * 32 + i, v[0], v[1], v[2], v[3]);
    }
    float* e_s = c.sm + SM::oE;
    if (c.half == 1) e_s[c.m] = dup;
    __syncthreads();
    if (c.half == 0) a.du[(size_t)tile * TM + c.m] += dup + e_s[c.m];
  }
  const RowSrc gsrc = tc_stage_rows<L>(c, (a.gamma[k] + c.goff), c.c0, nvalid);    // dGamma rows are consumed (barrier above)
  tc_mlp_bwd_hidden_st<L>(c, tl.m1_b, tl.m0_bx, zd, &tl.m0_bs[0]);
  tc_din<L>(c, tl, dXg);
  float dYp[D::NSH];
  tc_tp_backward<L, KIND, FIRST, true>(a, lw, c, tile, k, es, nvalid, a.dV[(k + 1) & 1], FIRST ? nullptr : a.dV[k & 1], (a.dgamma[k] + c.goff), dYp, gsrc);
  tc_dy_store<L, false>(a, c, tile, dYp, FIRST);
}
template <int L, char KIND, bool FIRST>
__global__ void __launch_bounds__(NT, DimsTC<L>::MINB) k_bk_tc(const __grid_constant__ ChunkArgs a, const __grid_constant__ ModelW w, const __grid_constant__ TcW tw, const int k) {
  using D = DimsTC<L>; constexpr int TM = 128;
  extern __shared__ __align__(1024) float sm_raw[];
  // persistent over the tiles of the chunk: tile = blockIdx.x, blockIdx.x + gridDim.x, ... (TMEM / mbarriers set up once per CTA)
  const ChunkBounds cb = chunk_bounds(a);
  const int ntiles = (cb.e1 - cb.e0 + TM - 1) / TM;
  if ((int)blockIdx.x >= ntiles) return;               // also: the empty chunks behind the last edge of a device-built plan
  TcCtx c = tc_begin<L>(sm_raw, tw);
  c.c0 = cb.c0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int es = cb.e0 + tile * TM;
    const int nvalid = min(TM, cb.e1 - es);
    const GeomIn gi = tc_geom_load(a, es, nvalid);
    bk_prefetch<L>(a, tile, k, gi, cb.c0, 0);
    const Geom g = tc_tile_begin<L>(a, w, c, gi, nvalid);
    bk_body<L, KIND, FIRST>(a, w, tw, c, g, tile, es, nvalid, k);
    __syncthreads();
  }
  tc_end(c);
  (void)sizeof(D);
}

// ============================================================================================
// B0
// ============================================================================================
template <int L>
__device__ __forceinline__ void b0_body(const ChunkArgs& a, const ModelW& w, const TcW& tw, TcCtx& c, const Geom& g, int tile, int es, int nvalid) {
  using D = DimsTC<L>; using SM = SmemTC<L>; constexpr int TM = 128;
  const float* X0 = a.X[0] + (size_t)tile * S * TM;
  tc_load_w<L>(c, tw.layer[0].env[0]);
  // ---- phase 2 of layer 0 and the embed backward: dx0 = dX + dw env0^T + dw0 emb^T
  op_load_rows64<L>(c, X0);
  const RowSrc dsrc = tc_stage_rows<L>(c, (a.dgamma[0] + c.goff), c.c0, nvalid);
  float dx0[32];
  const float* dXg = a.dX + (size_t)tile * S * TM;
  const float* zd = a.ZD[0] + (size_t)tile * ZD_ROWS * TM;          // act'(z1 + bias), act'(z2), m0 of the two-body MLP (written by F0)
  float du_tot;
  {
    float dp[32], mv[32];
    tc_phase2<L>(a, w, c, tw.layer[0], tw.emb_b, tile, X0, tw.two2_b, dsrc, dx0,
                 [&] { ld_rows32(c, dXg, dp); ld_rows32(c, zd + 128 * TM, mv); });
    float dup = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float dxv = dx0[i + q] + dp[i + q];
        dup += dxv * mv[i + q];
        v[q] = dxv * g.u;
      }
      op_put4<L>(c, c.half * 32 + i, v[0], v[1], v[2], v[3]);
    }
    float* e_s = c.sm + SM::oE;
    e_s[c.half * TM + c.m] = dup;
    __syncthreads();
    du_tot = e_s[c.m] + e_s[TM + c.m] + a.du[(size_t)tile * TM + c.m];
  }
  tc_mlp_bwd_hidden_st<L>(c, tw.two1_b, tw.two0_b, zd);
  // Bessel basis and its radial derivative (for d/dr of bessel * u)
  float bes[MAXB], dbes[MAXB];
  {
    // sin / cos((n+1) theta) by the Chebyshev recurrence (one sincosf per edge); only half 0 consumes them
    const float pref = sqrtf(2.0f / g.rc), ir = 1.0f / g.r, irc = 1.0f / g.rc;
    float s1, c1;
    sincosf(3.14159265358979323846f * (g.r * irc), &s1, &c1);
    const float c2 = 2.0f * c1;
    float sp = 0.f, sn = s1, cp = 1.f, cs = c1;
#pragma unroll
    for (int n = 0; n < MAXB; ++n) {
      const float kn = (float)(n + 1) * 3.14159265358979323846f;
      const bool on = n < w.B;
      bes[n] = on ? pref * sn * ir : 0.f;
      dbes[n] = on ? pref * (kn * irc * cs * ir - sn * ir * ir) : 0.f;
      const float ns = c2 * sn - sp, nc = c2 * cs - cp;
      sp = sn; sn = ns; cp = cs; cs = nc;
    }
  }
  tc_mma<L>(c, 64, 32, TC_SCR);                          // d(bessel*u) = dz1 W0[bessel rows]^T  (N padded to 32)
  float gx = 0.f, gy = 0.f, gz = 0.f;
  float* G3 = c.sm + SM::oOPH;                           // [3][128] g components ; virial rows after
  float* VR = G3 + 3 * TM;                               // [6][128]
  if (c.half == 0) {
    float v[16];
    tc_ld16(c, TC_SCR, v);
    float dr = du_tot * g.dudr;
#pragma unroll
    for (int n = 0; n < MAXB; ++n) dr += v[n] * (dbes[n] * g.u + bes[n] * g.dudr);
    float dYt[D::NSH];
    const float* dYg = a.dY + (size_t)tile * D::NSH * TM;
    const float* DY_s = c.sm + SM::oDY;
#pragma unroll
    for (int lm = 0; lm < D::NSH; ++lm) dYt[lm] = dYg[lm * TM + c.m] + DY_s[lm * TM + c.m];
    float qx, qy, qz;
    sph_harm_vjp<L>(g.x, g.y, g.z, dYt, qx, qy, qz);
    const float nq = g.x * qx + g.y * qy + g.z * qz;
    const float ir = 1.0f / g.r;
    gx = dr * g.x + (qx - g.x * nq) * ir;
    gy = dr * g.y + (qy - g.y * nq) * ir;
    gz = dr * g.z + (qz - g.z * nq) * ir;
    if (c.m >= nvalid) { gx = gy = gz = 0.f; }
  }
  __syncthreads();                                       // operand reads (MMA) are complete; reuse OPH
  if (c.half == 0) {
    const int t = c.m;
    G3[0 * TM + t] = gx; G3[1 * TM + t] = gy; G3[2 * TM + t] = gz;
    if (t < nvalid) {
      const int e = es + t;
      if (a.edge_grad) { a.edge_grad[3 * (size_t)e + 0] = gx; a.edge_grad[3 * (size_t)e + 1] = gy; a.edge_grad[3 * (size_t)e + 2] = gz; }
      const int j = a.edge_j[e];
      atomicAdd(a.facc + 3 * (size_t)j + 0, (unsigned long long)__double2ll_rn(-(double)gx * FIX_SCALE));
      atomicAdd(a.facc + 3 * (size_t)j + 1, (unsigned long long)__double2ll_rn(-(double)gy * FIX_SCALE));
      atomicAdd(a.facc + 3 * (size_t)j + 2, (unsigned long long)__double2ll_rn(-(double)gz * FIX_SCALE));
      const float rx = g.x * g.r, ry = g.y * g.r, rz = g.z * g.r;
      VR[0 * TM + t] = -rx * gx; VR[1 * TM + t] = -ry * gy; VR[2 * TM + t] = -rz * gz;
      VR[3 * TM + t] = -0.5f * (rx * gy + ry * gx); VR[4 * TM + t] = -0.5f * (rx * gz + rz * gx); VR[5 * TM + t] = -0.5f * (ry * gz + rz * gy);
    } else {
#pragma unroll
      for (int q = 0; q < 6; ++q) VR[q * TM + t] = 0.f;
    }
  }
  __syncthreads();
  {
    const int t = threadIdx.x;
    const int* c_s = reinterpret_cast<const int*>(c.sm + SM::oC);
    const int* seg = reinterpret_cast<const int*>(c.sm + SM::oSEG);
    const int nseg = seg[TM + 1];
    // F_i += sum of g_e over each centre run (fixed order), one atomic per (centre, component)
    for (int w2 = t; w2 < 3 * nseg; w2 += NT) {
      const int q = w2 % 3, sgm = w2 / 3;
      double acc = 0.0;
      for (int e = seg[sgm]; e < seg[sgm + 1]; ++e) acc += (double)G3[q * TM + e];
      atomicAdd(a.facc + 3 * (size_t)a.ilist[c_s[seg[sgm]]] + q, (unsigned long long)__double2ll_rn(acc * FIX_SCALE));
    }
    // virial: exact integer (fixed-point) warp reduction, one atomic per warp and component
    if (a.vacc && t >= 128) {
      const int e = t - 128;
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        long long v = __double2ll_rn((double)VR[q * TM + e] * VIR_SCALE);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((t & 31) == 0) atomicAdd(a.vacc + q, (unsigned long long)v);
      }
    }
  }
}
template <int L>
__global__ void __launch_bounds__(NT, DimsTC<L>::MINB) k_b0_tc(const __grid_constant__ ChunkArgs a, const __grid_constant__ ModelW w, const __grid_constant__ TcW tw) {
  using D = DimsTC<L>; constexpr int TM = 128;
  extern __shared__ __align__(1024) float sm_raw[];
  // persistent over the tiles of the chunk: tile = blockIdx.x, blockIdx.x + gridDim.x, ... (TMEM / mbarriers set up once per CTA)
  const ChunkBounds cb = chunk_bounds(a);
  const int ntiles = (cb.e1 - cb.e0 + TM - 1) / TM;
  if ((int)blockIdx.x >= ntiles) return;               // also: the empty chunks behind the last edge of a device-built plan
  TcCtx c = tc_begin<L>(sm_raw, tw);
  c.c0 = cb.c0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int es = cb.e0 + tile * TM;
    const int nvalid = min(TM, cb.e1 - es);
    const GeomIn gi = tc_geom_load(a, es, nvalid);
    b0_prefetch<L>(a, tile, gi, cb.c0, 0);
    const Geom g = tc_tile_begin<L>(a, w, c, gi, nvalid);
    b0_body<L>(a, w, tw, c, g, tile, es, nvalid);
    __syncthreads();
  }
  tc_end(c);
  (void)sizeof(D);
}

// ============================================================================================
// Fused persistent kernel.  The edge list is cut into centre-aligned BATCHES of at most B*128 edges (plan built on the
// device by k_plan); a batch is B edge-aligned 128-edge tiles, exactly the layout of a (tiny) chunk of the chunked
// pipeline.  One CTA takes a batch from a queue and runs it phase by phase -- all its tiles through F0, then FK, ... --
// so the code of one phase stays in the instruction cache for the whole batch, and combines the per-centre sums of
// rows that straddle two tiles itself (fused_fixup) before the next phase starts: no kernel boundaries, no fix-up
// launches, no host synchronisation.  The inter-phase state lives in CTA-private scratch slots that are rewritten for
// every batch.  TMEM and the mbarriers are set up once per CTA.
// ============================================================================================
struct FusedPlan {
  const int* batch_c0;    // [nbatch + 1] first centre slot of every batch (batch_c0[nbatch] = nlocal)
  int* info;              // [0] = nbatch, [1] = max degree, [2] = E, [3] = edge capacity overflow flag, [5] = batch queue, [6] = tiles
  int batch;              // B: tiles per batch (scratch slots per CTA); a centre needs at most B*128 edges
  int num_sms;
  unsigned* sm_phase;     // [number of SMs] phase tickets of the CTAs resident on each SM (zeroed before the launch), or nullptr
};
// Keep the CTAs that share an SM in the SAME phase: every CTA adds one ticket per phase it enters and waits (bounded) until
// all its co-residents have entered that phase too.  Two CTAs in different phases execute ~130 KB of different code and
// evict each other from the instruction cache (measured: 15 % of all warp stalls were "no instruction"; the chunked
// pipeline, where an SM only ever runs one phase at a time, has none).  The wait is bounded by a cycle budget, so a CTA
// without a partner (odd residency, partner already finished) only loses the alignment, never deadlocks.
__device__ __forceinline__ void fused_phase_align(unsigned* sm_phase, unsigned& my_phase, unsigned residents) {
  if (sm_phase == nullptr || residents < 2) return;
  if (threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    ++my_phase;
    atomicAdd(sm_phase + smid, 1u);
    const unsigned want = my_phase * residents;
    const long long t0 = clock64();
    while (*reinterpret_cast<volatile unsigned*>(sm_phase + smid) < want && clock64() - t0 < 400000) { }
  }
  __syncthreads();
}
// rows of centres whose CSR row straddles tiles of this batch: the tile in which the row STARTS adds the carries of the
// following tiles in tile order (same rule as k_fixup of the chunked pipeline)
template <int NF>
__device__ __forceinline__ void fused_fixup(const ChunkArgs& a, const TcCtx& c, int e0, int e1, int nb, int slot0, float* out) {
  // one warp per tile boundary (a batch has at most 63 of them); the lanes share the row's features
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 1
  for (int tb = warp; tb + 1 < nb; tb += NT / 32) {
    const int es = e0 + tb * 128, ee = min(es + 128, e1);
    const int ce = a.edge_c[ee - 1];
    const int rb = a.rowptr[ce], re = a.rowptr[ce + 1];
    if (rb < es || re <= ee) continue;                          // row does not start here, or ends here
    for (int f = lane; f < NF; f += 32) {
      float acc = out[(size_t)(ce - c.c0) * NF + f];
      for (int t2 = tb + 1; t2 < nb && e0 + t2 * 128 < re; ++t2) acc += a.carry[(size_t)(slot0 + t2) * NF + f];
      out[(size_t)(ce - c.c0) * NF + f] = acc;
    }
  }
  __syncthreads();
}
__device__ __forceinline__ void fused_fixup_e(const ChunkArgs& a, int e0, int e1, int nb, int slot0) {
  for (int tb = threadIdx.x; tb + 1 < nb; tb += NT) {
    const int es = e0 + tb * 128, ee = min(es + 128, e1);
    const int ce = a.edge_c[ee - 1];
    const int rb = a.rowptr[ce], re = a.rowptr[ce + 1];
    if (rb < es || re <= ee) continue;
    double acc = a.esum[ce];
    for (int t2 = tb + 1; t2 < nb && e0 + t2 * 128 < re; ++t2) acc += a.ecarry[slot0 + t2];
    a.esum[ce] = acc;
  }
}
// The batch / loop state lives in shared memory (volatile), not in registers: the phase bodies were tuned to exactly
// 128 (255) registers as stand-alone kernels; ~20 more values live across them made ptxas spill (800-byte frames for
// l_max = 2) and the fused kernel up to 1.7x slower than the per-phase kernels.  Every tile-phase rebuilds its context
// from shared memory, so nothing but the kernel parameters (constant bank) is live across a phase body.
enum { FS_E0 = 0, FS_E1, FS_NB, FS_TB, FS_PHASE, FS_MPH, FS_WPH, FS_WPH2, FS_C0, FS_BATCH, FS_COUNT };
template <int L> __device__ __forceinline__ volatile int* fused_state(float* sm_raw) {
  float* sm = sm_raw + (((1024u - (umma::smem_u32(sm_raw) & 1023u)) & 1023u) >> 2);
  return reinterpret_cast<volatile int*>(sm + SmemTC<L>::oBAR) + 12;
}
template <int L> __device__ __forceinline__ TcCtx fused_ctx(float* sm_raw, const TcW& tw, int batch) {
  using SM = SmemTC<L>;
  TcCtx c;
  c.sm = sm_raw + (((1024u - (umma::smem_u32(sm_raw) & 1023u)) & 1023u) >> 2);
  c.mbar = reinterpret_cast<uint64_t*>(c.sm + SM::oBAR);
  c.wbar = c.mbar + 1;
  c.wbar2 = c.mbar + 3;
  c.tmem = *reinterpret_cast<volatile uint32_t*>(c.mbar + 2);
  volatile int* fs = reinterpret_cast<volatile int*>(c.sm + SM::oBAR) + 12;
  c.mph = (uint32_t)fs[FS_MPH]; c.wph = (uint32_t)fs[FS_WPH]; c.wph2 = (uint32_t)fs[FS_WPH2];
  c.passes = tw.passes;
  const int t = threadIdx.x;
  c.m = t & 127; c.half = t >> 7; c.q = (t >> 5) & 3;
  c.c0 = fs[FS_C0];
  c.goff = (size_t)((int)blockIdx.x * batch) * 128 * DimsTC<L>::F;
  return c;
}
__device__ __forceinline__ void fused_phase_align_s(unsigned* sm_phase, volatile int* fs, unsigned residents) {
  if (sm_phase == nullptr || residents < 2) return;
  if (threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const unsigned mine = (unsigned)fs[FS_PHASE] + 1u;
    fs[FS_PHASE] = (int)mine;
    atomicAdd(sm_phase + smid, 1u);
    const unsigned want = mine * residents;
    const long long t0 = clock64();
    while (*reinterpret_cast<volatile unsigned*>(sm_phase + smid) < want && clock64() - t0 < 400000) { }
  }
  __syncthreads();
}
template <int L, int NLAYERS>
__global__ void __launch_bounds__(NT, DimsTC<L>::MINB) k_fused_tc(const __grid_constant__ ChunkArgs a, const __grid_constant__ ModelW w, const __grid_constant__ TcW tw,
                                                    const __grid_constant__ FusedPlan plan) {
  using D = DimsTC<L>; constexpr int TM = 128;
  extern __shared__ __align__(1024) float sm_raw[];
  if (plan.info[1] > plan.batch * TM || plan.info[3] != 0) return;      // a centre with more edges than a batch holds / edge arrays too small: the host falls back
  {
    TcCtx c = tc_begin<L>(sm_raw, tw);
    volatile int* fs = fused_state<L>(sm_raw);
    if (threadIdx.x < FS_COUNT) fs[threadIdx.x] = 0;
    __syncthreads();
    (void)c;
  }
  // one phase of one tile of the batch: geometry -> shared memory, then the phase body on the tile's private scratch slot
#define ALG_PHASE(PRE, ...)                                                                        \
  fused_phase_align_s(plan.sm_phase, fused_state<L>(sm_raw), (gridDim.x + plan.num_sms - 1) / plan.num_sms); \
  _Pragma("unroll 1") for (;;) {                                                                    \
    volatile int* fs = fused_state<L>(sm_raw);                                                      \
    const int tb = fs[FS_TB];                                                                       \
    if (tb >= fs[FS_NB]) break;                                                                     \
    TcCtx c = fused_ctx<L>(sm_raw, tw, plan.batch);                                                 \
    const int es = fs[FS_E0] + tb * TM, nvalid = min(TM, fs[FS_E1] - es), slot = (int)blockIdx.x * plan.batch + tb; \
    const GeomIn gi = tc_geom_load(a, es, nvalid);                                                  \
    PRE;                                                                                            \
    const Geom g = tc_tile_begin<L>(a, w, c, gi, nvalid);                                           \
    __VA_ARGS__;                                                                                    \
    __syncthreads();                                                                                \
    if (threadIdx.x == 0) { fs[FS_TB] = fs[FS_TB] + 1; fs[FS_MPH] = (int)c.mph; fs[FS_WPH] = (int)c.wph; fs[FS_WPH2] = (int)c.wph2; } \
    __syncthreads();                                                                                \
  }                                                                                                 \
  __syncthreads();               /* every thread has seen the terminating tile index before it is reset */ \
  if (threadIdx.x == 0) fused_state<L>(sm_raw)[FS_TB] = 0;                                          \
  __syncthreads();
#define ALG_FIX(ptr)                                                                                \
  {                                                                                                 \
    volatile int* fs = fused_state<L>(sm_raw);                                                      \
    const TcCtx c = fused_ctx<L>(sm_raw, tw, plan.batch);                                           \
    fused_fixup<D::F>(a, c, fs[FS_E0], fs[FS_E1], fs[FS_NB], (int)blockIdx.x * plan.batch, (ptr) + c.goff); \
  }
#pragma unroll 1
  for (;;) {
    // dynamic batch queue: the result does not depend on which CTA runs a batch (fixed-point accumulation), so the
    // kernel balances itself whatever the number of resident CTAs is
    {
      volatile int* fs = fused_state<L>(sm_raw);
      if (threadIdx.x == 0) {
        const int b = atomicAdd(plan.info + 5, 1);
        fs[FS_BATCH] = b;
        if (b < plan.info[0]) {
          const int bc0 = plan.batch_c0[b];
          const int e0 = a.rowptr[bc0], e1 = a.rowptr[plan.batch_c0[b + 1]];
          fs[FS_E0] = e0; fs[FS_E1] = e1; fs[FS_NB] = (e1 - e0 + TM - 1) / TM; fs[FS_TB] = 0; fs[FS_C0] = bc0;
        }
      }
      __syncthreads();
      if (fs[FS_BATCH] >= plan.info[0]) {
        if (plan.sm_phase && threadIdx.x == 0) { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); atomicAdd(plan.sm_phase + smid, 1u << 24); }
        break;
      }
      if (fs[FS_NB] <= 0) {                                     // only centres without neighbours: hand in the tickets of the skipped phases
        if (plan.sm_phase && threadIdx.x == 0) {
          unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
          fs[FS_PHASE] = fs[FS_PHASE] + 2 * NLAYERS + 1;
          atomicAdd(plan.sm_phase + smid, 2u * NLAYERS + 1u);
        }
        __syncthreads();
        continue;
      }
    }
    ALG_PHASE((void)0, f0_body<L>(a, w, tw, c, g, slot, es, nvalid));
    ALG_FIX(a.gamma[0]);
    if constexpr (NLAYERS == 1) {
      ALG_PHASE(fk_prefetch<L>(a, &c, slot, 0, gi, c.c0, c.goff), (t_body<L, true>(a, w, tw, c, g, slot, es, nvalid, 0)));
    } else if constexpr (NLAYERS == 2) {
      ALG_PHASE(fk_prefetch<L>(a, &c, slot, 0, gi, c.c0, c.goff), (fk_body<L, 'B', true>(a, w, tw, c, g, slot, es, nvalid, 0)));
      ALG_FIX(a.gamma[1]);
      ALG_PHASE(fk_prefetch<L>(a, &c, slot, 1, gi, c.c0, c.goff), (t_body<L, false>(a, w, tw, c, g, slot, es, nvalid, 1)));
    } else {
      ALG_PHASE(fk_prefetch<L>(a, &c, slot, 0, gi, c.c0, c.goff), (fk_body<L, 'C', true>(a, w, tw, c, g, slot, es, nvalid, 0)));
      ALG_FIX(a.gamma[1]);
      ALG_PHASE(fk_prefetch<L>(a, &c, slot, 1, gi, c.c0, c.goff), (fk_body<L, 'D', false>(a, w, tw, c, g, slot, es, nvalid, 1)));
      ALG_FIX(a.gamma[2]);
      ALG_PHASE(fk_prefetch<L>(a, &c, slot, 2, gi, c.c0, c.goff), (t_body<L, false>(a, w, tw, c, g, slot, es, nvalid, 2)));
    }
    {
      volatile int* fs = fused_state<L>(sm_raw);
      fused_fixup_e(a, fs[FS_E0], fs[FS_E1], fs[FS_NB], (int)blockIdx.x * plan.batch);
    }
    ALG_FIX(a.dgamma[NLAYERS - 1]);
    if constexpr (NLAYERS == 2) {
      ALG_PHASE(bk_prefetch<L>(a, slot, 0, gi, c.c0, c.goff), (bk_body<L, 'B', true>(a, w, tw, c, g, slot, es, nvalid, 0)));
      ALG_FIX(a.dgamma[0]);
    } else if constexpr (NLAYERS == 3) {
      ALG_PHASE(bk_prefetch<L>(a, slot, 1, gi, c.c0, c.goff), (bk_body<L, 'D', false>(a, w, tw, c, g, slot, es, nvalid, 1)));
      ALG_FIX(a.dgamma[1]);
      ALG_PHASE(bk_prefetch<L>(a, slot, 0, gi, c.c0, c.goff), (bk_body<L, 'C', true>(a, w, tw, c, g, slot, es, nvalid, 0)));
      ALG_FIX(a.dgamma[0]);
    }
    ALG_PHASE(b0_prefetch<L>(a, slot, gi, c.c0, c.goff), b0_body<L>(a, w, tw, c, g, slot, es, nvalid));
  }
#undef ALG_PHASE
#undef ALG_FIX
  {
    TcCtx c = fused_ctx<L>(sm_raw, tw, plan.batch);
    tc_end(c);
  }
}

}  // namespace alg
