// Shared definitions of the Allegro B200 engine: model constants, weight/argument structs,
// and the device building blocks (tile GEMM, activations, geometry, segmented sums).
//
// Tile model: edges are centre-sorted (CSR).  A chunk is a centre-aligned edge range
// [e0,e1); tile t of a chunk covers TM consecutive edges.  Per-edge state between kernels
// lives in HBM/L2 as tile-SoA arrays  buf[(tile*ROWS + row)*TM + e]  so that one warp reads
// 32 consecutive edges of one feature row (128 B, coalesced).  Inside a kernel a tile's
// activations sit in shared memory as [row][TM].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace alg {

constexpr int S = 64;      // num_scalar_features (latent width)
constexpr int H = 64;      // MLP hidden width
constexpr int U = 32;      // num_tensor_features
constexpr int R = 32;      // readout hidden width
constexpr int MAXB = 16;   // max num_bessels
constexpr int MAXT = 8;    // max model types
constexpr int NT = 256;    // threads per CTA
constexpr float ACT_C = 1.6765324703310909f;
constexpr double FIX_SCALE = 4294967296.0;          // 2^32 fixed point for force accumulation
constexpr double FIX_INV = 1.0 / 4294967296.0;
constexpr double VIR_SCALE = 16777216.0;            // 2^24 fixed point for the virial

template <int L> struct Dims {
  static constexpr int TM = 64;                     // edges per tile (2 CTAs/SM for L <= 2)
  static constexpr int MINB = (L == 3) ? 1 : 2;     // resident CTAs per SM targeted by __launch_bounds__
  static constexpr int NSH = (L + 1) * (L + 1);
  static constexpr int NL = L + 1;
  static constexpr int ENVW = NL * U;               // env / embed linear width, column = l*U+u
  static constexpr int SIN = S + NL * U;            // latent MLP input width (x || s), n0 = L+1
  static constexpr int F = NSH * U;                 // Gamma features per centre, f = lm*U+u
  static constexpr int CPH = NT / TM;               // channel phases of the (edge,channel) mapping
  static constexpr int CPT = U / CPH;               // channels per thread
  static constexpr int WS = ENVW + 1;               // stride of the edge-major W_s buffer
  static constexpr int CHU = (L == 1) ? 32 : ((L == 2) ? 16 : 8);   // channels per dGamma pass
  static constexpr int FC = NSH * CHU;
  static constexpr int DGS = FC + 1;
};

__host__ __device__ constexpr int lsel(int lm) { return lm >= 9 ? 3 : (lm >= 4 ? 2 : (lm >= 1 ? 1 : 0)); }

struct MLPW {
  const float* w[3];    // w[i]  : [K_i][N_i] row-major (in-features major)
  const float* wt[3];   // wt[i] : [N_i][K_i] (transposed, for input gradients)
};
struct LayerW {
  const float* env;     // [S][ENVW]
  const float* env_t;   // [ENVW][S]
  const float* omega;   // [npaths][U]
  MLPW mlp;             // SIN -> H -> H -> S
  float a, b;           // residual mix: x' = a*x + b*xt, a = 1/sqrt(1+alpha^2), b = alpha*a
};
struct ModelW {
  MLPW two;             // (2T+B) -> H -> H -> S
  const float* emb;     // [S][ENVW]
  const float* emb_t;   // [ENVW][S]
  LayerW layer[3];
  const float* ro0;     // [S][R]
  const float* ro0_t;   // [R][S]
  const float* ro1;     // [R]
  float rc[MAXT * MAXT];        // per-edge-type cutoff (model types)
  float gscale[MAXT];           // inv_sqrt_n * per-type scale  (dE_tot/dE_edge)
  int T, B, nl;
  float p;
  float inv_sqrt_n;
};

// per-step edge arrays + per-chunk buffers
struct ChunkArgs {
  // per step
  const float4* rvec;   // [E] (x_j - x_i) fp32, w = bits(zi | zj<<8)
  const int* edge_j;    // [E] neighbour atom index
  const int* edge_c;    // [E] centre slot ii
  const int* rowptr;    // [nlocal+1]
  const int* ilist;     // [nlocal] centre slot -> atom index
  // chunk: edge range [e0, e1) and first centre slot c0 -- by value (host-built plan), or read from the device-built plan
  // entry plan[3*ci .. 3*ci+2] when `plan` is set (no host synchronisation; see k_chunk_plan in alg_api.cu)
  int e0, e1, c0;
  const int* plan;
  int ci;
  float* X[3];          // x^k, k = 0..nl-1     [tile][S][TM]
  float* W0;            // embed weights w0 / later dw0  [tile][ENVW][TM]
  float* V[3];          // V^k, k = 1..nl-1     [tile][U][DIM_k][TM]
  float* dX;            // [tile][S][TM]
  float* ZD[4];         // tensor-core pipeline: stored act'(z1), act'(z2), m per MLP stage (0 = two-body, 1+k = layer k)  [tile][3*64][TM]
  float* dV[2];         // ping-pong gradient of V  [tile][U][DIM][TM]
  float* dY;            // [tile][NSH][TM]
  float* du;            // [tile][TM]
  float* gamma[3];      // [centre - c0][F]
  float* dgamma[3];
  float* carry;         // [tile][F]
  double* ecarry;       // [tile]
  double* esum;         // [nlocal] raw sum of edge energies per centre slot
  float* edge_energy;   // [E] or nullptr
  float* edge_grad;     // [E][3] or nullptr (debug)
  unsigned long long* facc;   // [ntot][3] fixed-point force accumulators
  unsigned long long* vacc;   // [6] fixed-point virial accumulators
};

struct ChunkBounds { int e0, e1, c0; };
__device__ __forceinline__ ChunkBounds chunk_bounds(const ChunkArgs& a) {
  if (a.plan) return ChunkBounds{a.plan[3 * a.ci], a.plan[3 * a.ci + 1], a.plan[3 * a.ci + 2]};
  return ChunkBounds{a.e0, a.e1, a.c0};
}

// ------------------------------------------------------------------------------------------
// sigmoid via the SFU: ex2.approx (2 ulp) + rcp.approx (1 ulp); absolute error of s <~ 2e-7,
// far inside the strict fp32 tolerance (checked per stage by tests/test_gpu_parity.py::test_intermediates)
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sigmoid_fast(float z) { return rcp_approx(1.0f + __expf(-z)); }
__device__ __forceinline__ float silu_act(float z, float& d) {
  const float s = sigmoid_fast(z);
  d = ACT_C * s * (1.0f + z * (1.0f - s));
  return ACT_C * z * s;
}
__device__ __forceinline__ float silu_act(float z) {
  return ACT_C * z * sigmoid_fast(z);
}

// ---- packed-pair versions for the tensor-core epilogues (sm_100 FFMA2 / FMUL2 / FADD2: two fp32 lanes per instruction).
// The epilogues are bound by instruction issue, not by FP32 throughput: c*silu and its derivative cost 7 packed
// instructions + 4 MUFU per PAIR (5.5 per element) instead of ~14 scalar ones.
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// r = c*z*s, d = c*(s + z*s*(1-s)), s = sigmoid(z) = 1/(1 + 2^(-z*log2 e))  (same formulas as silu_act, up to rounding)
__device__ __forceinline__ void silu_act2(float2 z, float2& r, float2& d) {
  const float2 t = __fmul2_rn(z, make_float2(-1.4426950408889634f, -1.4426950408889634f));
  const float2 den = __fadd2_rn(make_float2(ex2_approx(t.x), ex2_approx(t.y)), make_float2(1.f, 1.f));
  const float2 s = make_float2(rcp_approx(den.x), rcp_approx(den.y));
  const float2 zs = __fmul2_rn(z, s);
  r = __fmul2_rn(zs, make_float2(ACT_C, ACT_C));
  const float2 oms = __ffma2_rn(s, make_float2(-1.f, -1.f), make_float2(1.f, 1.f));
  d = __fmul2_rn(__ffma2_rn(zs, oms, s), make_float2(ACT_C, ACT_C));
}
__device__ __forceinline__ float2 silu_act2(float2 z) {
  const float2 t = __fmul2_rn(z, make_float2(-1.4426950408889634f, -1.4426950408889634f));
  const float2 den = __fadd2_rn(make_float2(ex2_approx(t.x), ex2_approx(t.y)), make_float2(1.f, 1.f));
  const float2 s = make_float2(rcp_approx(den.x), rcp_approx(den.y));
  return __fmul2_rn(__fmul2_rn(z, s), make_float2(ACT_C, ACT_C));
}

// ------------------------------------------------------------------------------------------
// Tile GEMM on the FP32 pipe:  out(m, n) = sum_k A_s[k*TM + m] * W[k*ldw + col0 + n]
// for a TM x NC output block; A in shared memory ([K][TM]), W in global memory (L1-resident,
// broadcast reads).  Thread (mg, ng) owns rows {q*MG*4 + mg*4 + 0..3} and columns
// {ng*4 + 0..3}: all shared/global accesses are 128-bit and bank-conflict free.
// The epilogue functor receives (row m, column n, value).  No barrier inside.
template <int TM, int NC> struct GemmCfg {
  static constexpr int NG = NC / 4;
  static constexpr int MG = (NT / NG < TM / 4) ? NT / NG : TM / 4;
  static constexpr int Q = TM / (MG * 4);
  static constexpr int ACTIVE = MG * NG;
};

template <int TM, int NC, class Epi>
__device__ __forceinline__ void gemm_tile(const float* __restrict__ A_s, int K, const float* __restrict__ W,
                                          int ldw, int col0, Epi epi) {
  using C = GemmCfg<TM, NC>;
  const int t = threadIdx.x;
  if (t >= C::ACTIVE) return;
  const int mg = t % C::MG, ng = t / C::MG;
  float acc[C::Q][4][4];
#pragma unroll
  for (int q = 0; q < C::Q; ++q)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[q][i][j] = 0.f;
  const float* wp = W + col0 + ng * 4;
  const float* ap = A_s + mg * 4;
  constexpr int KB = 8;   // weight rows prefetched one block ahead (hides L2 latency of the weight stream)
  const int nkb = K / KB;
  float4 wn[KB];
  if (nkb > 0) {
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) wn[kk] = __ldg(reinterpret_cast<const float4*>(wp + (size_t)kk * ldw));
  }
#pragma unroll 1
  for (int kb = 0; kb < nkb; ++kb) {
    float4 wc[KB];
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) wc[kk] = wn[kk];
    if (kb + 1 < nkb) {
#pragma unroll
      for (int kk = 0; kk < KB; ++kk) wn[kk] = __ldg(reinterpret_cast<const float4*>(wp + (size_t)((kb + 1) * KB + kk) * ldw));
    }
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) {
      const int k = kb * KB + kk;
      const float w[4] = {wc[kk].x, wc[kk].y, wc[kk].z, wc[kk].w};
#pragma unroll
      for (int q = 0; q < C::Q; ++q) {
        const float4 a4 = *reinterpret_cast<const float4*>(ap + k * TM + q * (C::MG * 4));
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[q][i][j] = fmaf(a[i], w[j], acc[q][i][j]);
      }
    }
  }
  for (int k = nkb * KB; k < K; ++k) {   // K tail (num_bessels not a multiple of 8)
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)k * ldw));
    const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int q = 0; q < C::Q; ++q) {
      const float4 a4 = *reinterpret_cast<const float4*>(ap + k * TM + q * (C::MG * 4));
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[q][i][j] = fmaf(a[i], w[j], acc[q][i][j]);
    }
  }
#pragma unroll
  for (int q = 0; q < C::Q; ++q)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) epi(q * (C::MG * 4) + mg * 4 + i, ng * 4 + j, acc[q][i][j]);
}

// N-wide GEMM split into 64/32 column blocks
template <int TM, int N, class Epi>
__device__ __forceinline__ void gemm_tile_n(const float* __restrict__ A_s, int K, const float* __restrict__ W, int ldw, Epi epi) {
  static_assert(N % 32 == 0, "N must be a multiple of 32");
#pragma unroll 1
  for (int c0 = 0; c0 + 64 <= N; c0 += 64)
    gemm_tile<TM, 64>(A_s, K, W, ldw, c0, [&](int m, int n, float v) { epi(m, c0 + n, v); });
  if (N % 64) {
    constexpr int c0 = N - 32;
    gemm_tile<TM, 32>(A_s, K, W, ldw, c0, [&](int m, int n, float v) { epi(m, c0 + n, v); });
  }
}

// ------------------------------------------------------------------------------------------
// real spherical harmonics (component normalisation, m = -l..l) and their gradient wrt the
// unconstrained components of the unit vector (same definition as tools/gen_tables.py).
template <int L> __device__ __forceinline__ void sph_harm(float x, float y, float z, float* Y) {
  Y[0] = 1.f;
  if (L >= 1) {
    const float s3 = 1.7320508075688772f;
    Y[1] = s3 * y; Y[2] = s3 * z; Y[3] = s3 * x;
  }
  if (L >= 2) {
    const float s15 = 3.872983346207417f, s5h = 1.118033988749895f;
    Y[4] = s15 * x * y; Y[5] = s15 * y * z; Y[6] = s5h * (3.f * z * z - 1.f);
    Y[7] = s15 * x * z; Y[8] = 0.5f * s15 * (x * x - y * y);
  }
  if (L >= 3) {
    const float a = 2.091650066335189f, b = 10.246950765959598f, c = 1.620185174601965f, d = 1.3228756555322954f;
    Y[9] = a * y * (3.f * x * x - y * y); Y[10] = b * x * y * z; Y[11] = c * y * (5.f * z * z - 1.f);
    Y[12] = d * (5.f * z * z * z - 3.f * z); Y[13] = c * x * (5.f * z * z - 1.f);
    Y[14] = 0.5f * b * (x * x - y * y) * z; Y[15] = a * x * (x * x - 3.f * y * y);
  }
}
// q = sum_lm dY[lm] * grad_n Y_lm
template <int L> __device__ __forceinline__ void sph_harm_vjp(float x, float y, float z, const float* dY, float& qx, float& qy, float& qz) {
  qx = qy = qz = 0.f;
  if (L >= 1) {
    const float s3 = 1.7320508075688772f;
    qy += s3 * dY[1]; qz += s3 * dY[2]; qx += s3 * dY[3];
  }
  if (L >= 2) {
    const float s15 = 3.872983346207417f, s5 = 2.23606797749979f;
    qx += s15 * (dY[4] * y + dY[7] * z + dY[8] * x);
    qy += s15 * (dY[4] * x + dY[5] * z - dY[8] * y);
    qz += s15 * (dY[5] * y + dY[7] * x) + 3.f * s5 * z * dY[6];
  }
  if (L >= 3) {
    const float a = 2.091650066335189f, b = 10.246950765959598f, c = 1.620185174601965f, d = 1.3228756555322954f;
    const float x2 = x * x, y2 = y * y, z2 = z * z;
    qx += dY[9] * 6.f * a * x * y + dY[10] * b * y * z + dY[13] * c * (5.f * z2 - 1.f) + dY[14] * b * x * z +
          dY[15] * a * (3.f * x2 - 3.f * y2);
    qy += dY[9] * a * (3.f * x2 - 3.f * y2) + dY[10] * b * x * z + dY[11] * c * (5.f * z2 - 1.f) - dY[14] * b * y * z -
          dY[15] * 6.f * a * x * y;
    qz += dY[10] * b * x * y + dY[11] * 10.f * c * y * z + dY[12] * d * (15.f * z2 - 3.f) + dY[13] * 10.f * c * x * z +
          dY[14] * 0.5f * b * (x2 - y2);
  }
}

// polynomial cutoff envelope u(x) and du/dx, x = r/rc in [0,1]; u = 0 for x >= 1
__device__ __forceinline__ void poly_cutoff(float x, float p, float& u, float& dudx) {
  if (x >= 1.f) { u = 0.f; dudx = 0.f; return; }
  const float xp = powf(x, p - 1.f);          // x^(p-1)
  const float c0 = 0.5f * (p + 1.f) * (p + 2.f), c1 = p * (p + 2.f), c2 = 0.5f * p * (p + 1.f);
  const float x1 = xp * x, x2 = x1 * x, x3 = x2 * x;   // x^p, x^(p+1), x^(p+2)
  u = 1.f - c0 * x1 + c1 * x2 - c2 * x3;
  dudx = -c0 * p * xp + c1 * (p + 1.f) * x1 - c2 * (p + 2.f) * x2;
}

// ------------------------------------------------------------------------------------------
// Deterministic segmented sum over the edges of a tile, one thread per feature, edges visited
// in order.  The value of (edge e, feature f) is val(e, f).  A centre whose CSR row started in
// an earlier tile gets its partial written to carry[f] (combined later, in tile order, by
// fixup_carry_kernel); every other centre is written directly.  c_s = centre slot per edge.
template <int NF, class Val>
__device__ __forceinline__ void segsum_tile(const int* __restrict__ c_s, int nvalid, int es, const int* __restrict__ rowptr,
                                            float scale, float* __restrict__ out, int c0, float* __restrict__ carry, Val val) {
  const int cfirst = c_s[0];
  const bool contin = rowptr[cfirst] < es;
  for (int f = threadIdx.x; f < NF; f += NT) {
    int cur = cfirst;
    bool first = true;
    float acc = 0.f;
    for (int e = 0; e < nvalid; ++e) {
      const int c = c_s[e];
      if (c != cur) {
        if (first && contin) carry[f] = acc * scale; else out[(size_t)(cur - c0) * NF + f] = acc * scale;
        cur = c; acc = 0.f; first = false;
      }
      acc += val(e, f);
    }
    if (first && contin) carry[f] = acc * scale; else out[(size_t)(cur - c0) * NF + f] = acc * scale;
  }
}


// ------------------------------------------------------------------------------------------
// Segment table of a tile: seg[0..nseg] = first tile-local edge of every centre run
// (seg[nseg] = nvalid), built in parallel by the first TM threads (ballot + prefix).
// Shared scratch: seg[TM + 1] ints, then nseg, then up to 8 per-warp counts.
// Must be called by ALL threads of the CTA (contains barriers); c_s must be visible.
template <int TM>
__device__ __forceinline__ void seg_build(const int* __restrict__ c_s, int nvalid, int* __restrict__ seg) {
  int* nseg_p = seg + TM + 1;
  int* wcnt = seg + TM + 2;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  bool flag = false;
  unsigned m = 0;
  if (t < TM) {
    flag = t < nvalid && (t == 0 || c_s[t] != c_s[t - 1]);
    m = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) wcnt[warp] = __popc(m);
  }
  __syncthreads();
  if (t < TM) {
    int base = 0;
    for (int w2 = 0; w2 < warp; ++w2) base += wcnt[w2];
    if (flag) seg[base + __popc(m & ((1u << lane) - 1u))] = t;
    if (t == TM - 1) {
      const int total = base + __popc(m);
      *nseg_p = total;
      seg[total] = nvalid;
    }
  }
  __syncthreads();
}

// Deterministic segmented sum with the segment table: work item (feature f, segment s) is one
// thread that adds the segment's edges in order (no per-edge branch, unrolled loads).
//   put(centre_slot, f, is_first_segment_of_tile, value)
template <int NF, int TM, class Val, class Put>
__device__ __forceinline__ void segsum_items(const int* __restrict__ c_s, const int* __restrict__ seg, Val val, Put put) {
  const int nseg = seg[TM + 1];
  for (int w = threadIdx.x; w < NF * nseg; w += NT) {
    const int f = w % NF, sgm = w / NF;
    const int e0 = seg[sgm], e1 = seg[sgm + 1];
    float acc = 0.f;
    int e = e0;
    for (; e + 4 <= e1; e += 4) {
      const float v0 = val(e, f), v1 = val(e + 1, f), v2 = val(e + 2, f), v3 = val(e + 3, f);
      acc += v0; acc += v1; acc += v2; acc += v3;
    }
    for (; e < e1; ++e) acc += val(e, f);
    put(c_s[e0], f, sgm == 0, acc);
  }
}

}  // namespace alg
