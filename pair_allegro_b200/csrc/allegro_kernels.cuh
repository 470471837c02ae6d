// The five fused edge-tile kernels of the Allegro force evaluation (strict-fp32 FFMA path).
//
//   F0      geometry, Bessel x cutoff, two-body MLP -> x^0 ; embed linear -> w0 ; env linear 0
//           -> Gamma_0 partial sums                                        (per edge + seg. sum)
//   FK<k>   layer k (not last): TP(V^k, Gamma_k) -> s, V^{k+1}; latent MLP -> x^{k+1};
//           env linear k+1 -> Gamma_{k+1}
//   T       last layer forward, readout -> E_e, E_i ; backward "phase 1" of the last layer
//   BK<k>   backward: "phase 2" of layer k+1 (needs dGamma_{k+1}) + "phase 1" of layer k
//   B0      backward: phase 2 of layer 0, embed/two-body/geometry backward -> g_e = dE/dr_e,
//           force accumulation (+i, -j incl. ghosts) and virial
//
// Kernel boundaries are exactly the points where a per-centre environment sum (forward:
// Gamma, backward: dGamma) must be complete.  The chain rule implemented here is stated (and
// checked against autograd) in oracle/analytic_numpy.py; reference call site replaced:
// /root/reference/pair_nequip_allegro.cpp:425 (torchscript_model.forward) incl. its in-graph
// autograd.
#pragma once
#include "alg_common.cuh"
#include "tp_gen.cuh"

namespace alg {

template <int L> struct Smem {
  using D = Dims<L>;
  static constexpr int TM = D::TM;
  static constexpr int oY = 0;
  static constexpr int oDY = oY + D::NSH * TM;
  static constexpr int oU = oDY + D::NSH * TM;
  static constexpr int oMisc = oU + TM;            // 3 rows: c_s (int), zz_s (int), e_s (float)
  static constexpr int oSeg = oMisc + 3 * TM;      // 2 rows (int): segment table seg[TM+1], nseg, warp counts
  static constexpr int oIN = oSeg + 2 * TM;
  static constexpr int oA = oIN + D::SIN * TM;
  static constexpr int oB = oA + 64 * TM;
  static constexpr int oPAD = oB + 64 * TM;
  static constexpr int oC = oPAD + 4 * TM;
  static constexpr int oD = oC + 64 * TM;
  static constexpr int TOTAL = oD + 64 * TM;       // floats
  static constexpr size_t BYTES = (size_t)TOTAL * sizeof(float);
  static_assert(D::WS * TM <= 132 * TM, "W_s must fit A+B+PAD");
  static_assert(D::DGS * TM <= 196 * TM, "dGamma staging must fit A+B+PAD+C");
  static_assert(D::CPH * D::NSH <= 64, "dY partials must fit region D");
};

// ---- small helpers -------------------------------------------------------------------------
template <int TM> __device__ __forceinline__ void load_rows(float* __restrict__ dst, const float* __restrict__ src, int nrows) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int i = threadIdx.x; i < nrows * TM / 4; i += NT) d4[i] = s4[i];
}

struct Geom { float x, y, z, r, rc, u, dudr; int zi, zj; };

// per-edge geometry for thread t < TM: fills Y_s, u_s, c_s, zz_s (and returns the scalars)
template <int L> __device__ __forceinline__ Geom edge_geom(const ChunkArgs& a, const ModelW& w, int es, int nvalid, float* sm) {
  using D = Dims<L>; using SM = Smem<L>; constexpr int TM = D::TM;
  const int t = threadIdx.x;
  const int e = es + min(t, nvalid - 1);
  const float4 rv = a.rvec[e];
  Geom g;
  const int zz = __float_as_int(rv.w);
  g.zi = zz & 255; g.zj = zz >> 8;
  g.r = sqrtf(rv.x * rv.x + rv.y * rv.y + rv.z * rv.z);
  g.x = rv.x / g.r; g.y = rv.y / g.r; g.z = rv.z / g.r;
  g.rc = w.rc[g.zi * MAXT + g.zj];
  float dudx;
  poly_cutoff(g.r / g.rc, w.p, g.u, dudx);
  g.dudr = dudx / g.rc;
  float Y[D::NSH];
  sph_harm<L>(g.x, g.y, g.z, Y);
#pragma unroll
  for (int k = 0; k < D::NSH; ++k) sm[SM::oY + k * TM + t] = Y[k];
  sm[SM::oU + t] = g.u;
  reinterpret_cast<int*>(sm + SM::oMisc)[t] = a.edge_c[e];
  reinterpret_cast<int*>(sm + SM::oMisc)[TM + t] = zz;
  return g;
}

// env linear output in W_s (edge-major) -> Gamma partial sums of this tile
template <int L> __device__ __forceinline__ void env_sum(const ChunkArgs& a, const ModelW& w, int tile, int es, int nvalid,
                                                          float* sm, float* __restrict__ gamma) {
  using D = Dims<L>; using SM = Smem<L>; constexpr int TM = D::TM;
  const float* W_s = sm + SM::oA;
  const float* Y_s = sm + SM::oY;
  const int* c_s = reinterpret_cast<const int*>(sm + SM::oMisc);
  const int* seg = reinterpret_cast<const int*>(sm + SM::oSeg);
  const bool contin = a.rowptr[c_s[0]] < es;
  float* carry = a.carry + (size_t)tile * D::F;
  const float sc = w.inv_sqrt_n;
  segsum_items<D::F, TM>(c_s, seg,
      [&](int e, int f) { const int lm = f / U, u = f % U; return W_s[e * D::WS + lsel(lm) * U + u] * Y_s[lm * TM + e]; },
      [&](int centre, int f, bool first, float v) {
        if (first && contin) carry[f] = v * sc; else gamma[(size_t)(centre - chunk_bounds(a).c0) * D::F + f] = v * sc;
      });
  (void)nvalid;
}
template <int L> __device__ __forceinline__ void seg_setup(float* sm, int nvalid) {
  using SM = Smem<L>;
  seg_build<Dims<L>::TM>(reinterpret_cast<const int*>(sm + SM::oMisc), nvalid, reinterpret_cast<int*>(sm + SM::oSeg));
}

// ============================================================================================
// F0
// ============================================================================================
template <int L>
__global__ void __launch_bounds__(NT, Dims<L>::MINB) k_f0(const __grid_constant__ ChunkArgs a, const __grid_constant__ ModelW w) {
  using D = Dims<L>; using SM = Smem<L>; constexpr int TM = D::TM;
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, tile = blockIdx.x;
  const ChunkBounds cb = chunk_bounds(a);
  const int es = cb.e0 + tile * TM;
  if (es >= cb.e1) return;
  const int nvalid = min(TM, cb.e1 - es);
  float* IN = sm + SM::oIN; float* A = sm + SM::oA; float* B = sm + SM::oB; float* C = sm + SM::oC; float* Dd = sm + SM::oD;
  const float* u_s = sm + SM::oU;
  const int* zz_s = reinterpret_cast<const int*>(sm + SM::oMisc) + TM;
  if (t < TM) {
    const Geom g = edge_geom<L>(a, w, es, nvalid, sm);
    const float pref = sqrtf(2.0f / g.rc);
    const float xr = g.r / g.rc;
    for (int n = 0; n < w.B; ++n) {
      const float arg = (float)(n + 1) * (3.14159265358979323846f * xr);
      IN[n * TM + t] = pref * sinf(arg) / g.r * g.u;
    }
  }
  __syncthreads();
  seg_setup<L>(sm, nvalid);
  {  // two-body MLP layer 0: one-hot rows are added in the epilogue
    const float* w0 = w.two.w[0];
    const int T = w.T;
    gemm_tile<TM, 64>(IN, w.B, w0 + 2 * T * H, H, 0, [&](int m, int n, float v) {
      const int zz = zz_s[m];
      v += __ldg(w0 + (zz & 255) * H + n) + __ldg(w0 + (T + (zz >> 8)) * H + n);
      A[n * TM + m] = silu_act(v);
    });
  }
  __syncthreads();
  gemm_tile<TM, 64>(A, H, w.two.w[1], H, 0, [&](int m, int n, float v) { B[n * TM + m] = silu_act(v); });
  __syncthreads();
  float* X0g = a.X[0] + (size_t)tile * S * TM;
  gemm_tile<TM, 64>(B, H, w.two.w[2], S, 0, [&](int m, int n, float v) {
    v *= u_s[m];
    C[n * TM + m] = v;
    X0g[n * TM + m] = v;
  });
  __syncthreads();
  // embed linear -> w0 (kept in HBM/L2 for the first tensor product and its backward)
  float* W0g = a.W0 + (size_t)tile * D::ENVW * TM;
  gemm_tile_n<TM, D::ENVW>(C, S, w.emb, D::ENVW, [&](int m, int n, float v) { W0g[n * TM + m] = v; });
  // env linear of layer 0 -> W_s (edge-major) -> Gamma_0
  float* W_s = A;
  gemm_tile_n<TM, D::ENVW>(C, S, w.layer[0].env, D::ENVW, [&](int m, int n, float v) { W_s[m * D::WS + n] = v; });
  __syncthreads();
  env_sum<L>(a, w, tile, es, nvalid, sm, a.gamma[0]);
  (void)Dd;
}

// ============================================================================================
// tensor-product helpers (thread = (edge e, channel phase uh))
// ============================================================================================
template <int L, bool FIRST, int DIN>
__device__ __forceinline__ void load_vin(const ChunkArgs& a, int tile, int k, int e, int u, const float* Y_s, float* Vin) {
  using D = Dims<L>; constexpr int TM = D::TM;
  if (FIRST) {
    const float* W0g = a.W0 + (size_t)tile * D::ENVW * TM;
#pragma unroll
    for (int l = 0; l <= L; ++l) {
      const float wv = W0g[(l * U + u) * TM + e];
#pragma unroll
      for (int lm = l * l; lm < (l + 1) * (l + 1); ++lm) Vin[lm] = wv * Y_s[lm * TM + e];
    }
  } else {
    const float* Vg = a.V[k] + ((size_t)tile * U + u) * DIN * TM;
#pragma unroll
    for (int c = 0; c < DIN; ++c) Vin[c] = Vg[c * TM + e];
  }
}

// ============================================================================================
// FK: layer k forward (k < nl-1)
// ============================================================================================
template <int L, char KIND, bool FIRST>
__global__ void __launch_bounds__(NT, Dims<L>::MINB) k_fk(const __grid_constant__ ChunkArgs a, const __grid_constant__ ModelW w, const int k) {
  using D = Dims<L>; using SM = Smem<L>; using TP = tpgen::TP<L, KIND>; constexpr int TM = D::TM;
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, tile = blockIdx.x;
  const ChunkBounds cb = chunk_bounds(a);
  const int es = cb.e0 + tile * TM;
  if (es >= cb.e1) return;
  const int nvalid = min(TM, cb.e1 - es);
  float* IN = sm + SM::oIN; float* A = sm + SM::oA; float* B = sm + SM::oB; float* C = sm + SM::oC;
  const float* u_s = sm + SM::oU; const float* Y_s = sm + SM::oY;
  const int* c_s = reinterpret_cast<const int*>(sm + SM::oMisc);
  const LayerW& lw = w.layer[k];
  load_rows<TM>(IN, a.X[k] + (size_t)tile * S * TM, S);
  if (t < TM) edge_geom<L>(a, w, es, nvalid, sm);
  __syncthreads();
  seg_setup<L>(sm, nvalid);
  {  // tensor product: s -> IN rows S.., V^{k+1} -> global
    const int e = t % TM, uh = t / TM;
    const float* gam = a.gamma[k] + (size_t)(c_s[e] - chunk_bounds(a).c0) * D::F;
    float* Vng = a.V[k + 1] + (size_t)tile * U * TP::DOUT * TM;
#pragma unroll 1
    for (int i = 0; i < D::CPT; ++i) {
      const int u = uh + D::CPH * i;
      float Vin[TP::DIN], G[D::NSH], Vout[TP::DOUT], s[TP::N0];
      load_vin<L, FIRST, TP::DIN>(a, tile, k, e, u, Y_s, Vin);
#pragma unroll
      for (int lm = 0; lm < D::NSH; ++lm) G[lm] = gam[lm * U + u];
      TP::template fwd<U>(Vin, G, lw.omega + u, Vout, s);
#pragma unroll
      for (int q = 0; q < TP::N0; ++q) IN[(S + q * U + u) * TM + e] = s[q];
#pragma unroll
      for (int c = 0; c < TP::DOUT; ++c) Vng[(u * TP::DOUT + c) * TM + e] = Vout[c];
    }
  }
  __syncthreads();
  gemm_tile<TM, 64>(IN, D::SIN, lw.mlp.w[0], H, 0, [&](int m, int n, float v) { A[n * TM + m] = silu_act(v); });
  __syncthreads();
  gemm_tile<TM, 64>(A, H, lw.mlp.w[1], H, 0, [&](int m, int n, float v) { B[n * TM + m] = silu_act(v); });
  __syncthreads();
  float* Xng = a.X[k + 1] + (size_t)tile * S * TM;
  gemm_tile<TM, 64>(B, H, lw.mlp.w[2], S, 0, [&](int m, int n, float v) {
    const float xn = lw.a * IN[n * TM + m] + lw.b * v * u_s[m];
    C[n * TM + m] = xn;
    Xng[n * TM + m] = xn;
  });
  __syncthreads();
  float* W_s = A;
  gemm_tile_n<TM, D::ENVW>(C, S, w.layer[k + 1].env, D::ENVW, [&](int m, int n, float v) { W_s[m * D::WS + n] = v; });
  __syncthreads();
  env_sum<L>(a, w, tile, es, nvalid, sm, a.gamma[k + 1]);
}

// ============================================================================================
// shared backward pieces
// ============================================================================================
// MLP recompute (forward) for a latent MLP: IN -> (H1 in A, D1 in B) -> (H2 in C, D2 in D)
template <int L>
__device__ __forceinline__ void mlp_fwd_keep(const MLPW& mw, int K0, const float* w0base, float* sm, bool twobody, int T) {
  using SM = Smem<L>; constexpr int TM = Dims<L>::TM;
  float* IN = sm + SM::oIN; float* A = sm + SM::oA; float* B = sm + SM::oB; float* C = sm + SM::oC; float* Dd = sm + SM::oD;
  const int* zz_s = reinterpret_cast<const int*>(sm + SM::oMisc) + TM;
  if (twobody) {
    const float* w0 = mw.w[0];
    gemm_tile<TM, 64>(IN, K0, w0base, H, 0, [&](int m, int n, float v) {
      const int zz = zz_s[m];
      v += __ldg(w0 + (zz & 255) * H + n) + __ldg(w0 + (T + (zz >> 8)) * H + n);
      float d; A[n * TM + m] = silu_act(v, d); B[n * TM + m] = d;
    });
  } else {
    gemm_tile<TM, 64>(IN, K0, w0base, H, 0, [&](int m, int n, float v) { float d; A[n * TM + m] = silu_act(v, d); B[n * TM + m] = d; });
  }
  __syncthreads();
  gemm_tile<TM, 64>(A, H, mw.w[1], H, 0, [&](int m, int n, float v) { float d; C[n * TM + m] = silu_act(v, d); Dd[n * TM + m] = d; });
  __syncthreads();
}

// after: A = dxt*m products (for du), IN[0:S] = dm = dxt*u.  Then: du reduce, dz2 -> D, dz1 -> B.
template <int L>
__device__ __forceinline__ void mlp_bwd_hidden(const MLPW& mw, float* sm) {
  using SM = Smem<L>; constexpr int TM = Dims<L>::TM;
  float* IN = sm + SM::oIN; float* B = sm + SM::oB; float* Dd = sm + SM::oD;
  gemm_tile<TM, 64>(IN, S, mw.wt[2], H, 0, [&](int m, int n, float v) { Dd[n * TM + m] *= v; });
  __syncthreads();
  gemm_tile<TM, 64>(Dd, H, mw.wt[1], H, 0, [&](int m, int n, float v) { B[n * TM + m] *= v; });
  __syncthreads();
}

// tensor-product backward over all channels in passes of CHU channels, followed by the
// deterministic segmented sum of dG into dgamma_out.  ds is read from IN rows S...
//   FIRST : Vin = w0*Y ; dVin -> dw0 (overwrites W0 in place), dY partial returned via dYp
//   else  : Vin from V[k] ; dVin -> dVout_buf
template <int L, char KIND, bool FIRST, bool HAS_DVOUT>
__device__ __forceinline__ void tp_backward(const ChunkArgs& a, const LayerW& lw, int tile, int k, int es, int nvalid, float* sm,
                                            const float* __restrict__ dVnext, float* __restrict__ dVprev,
                                            float* __restrict__ dgamma_out, float* dYp /*[NSH] per thread, FIRST only*/) {
  using D = Dims<L>; using SM = Smem<L>; using TP = tpgen::TP<L, KIND>; constexpr int TM = D::TM;
  const int t = threadIdx.x;
  float* IN = sm + SM::oIN; float* DG = sm + SM::oA;
  const float* Y_s = sm + SM::oY;
  const int* c_s = reinterpret_cast<const int*>(sm + SM::oMisc);
  const int e = t % TM, uh = t / TM;
  const float* gam = a.gamma[k] + (size_t)(c_s[e] - chunk_bounds(a).c0) * D::F;
  float* W0g = a.W0 + (size_t)tile * D::ENVW * TM;
  if (FIRST) {
#pragma unroll
    for (int lm = 0; lm < D::NSH; ++lm) dYp[lm] = 0.f;
  }
#pragma unroll 1
  for (int pass = 0; pass < U / D::CHU; ++pass) {
#pragma unroll 1
    for (int ul = uh; ul < D::CHU; ul += D::CPH) {
      const int u = pass * D::CHU + ul;
      float Vin[TP::DIN], G[D::NSH], dVout[TP::DOUT], ds[TP::N0], dVin[TP::DIN], dG[D::NSH];
      load_vin<L, FIRST, TP::DIN>(a, tile, k, e, u, Y_s, Vin);
#pragma unroll
      for (int lm = 0; lm < D::NSH; ++lm) G[lm] = gam[lm * U + u];
      if (HAS_DVOUT) {
        const float* dVg = dVnext + ((size_t)tile * U + u) * TP::DOUT * TM;
#pragma unroll
        for (int c = 0; c < TP::DOUT; ++c) dVout[c] = dVg[c * TM + e];
      }
#pragma unroll
      for (int q = 0; q < TP::N0; ++q) ds[q] = IN[(S + q * U + u) * TM + e];
      TP::template bwd<U>(Vin, G, lw.omega + u, dVout, ds, dVin, dG);
      if (FIRST) {
#pragma unroll
        for (int l = 0; l <= L; ++l) {
          const float wv = W0g[(l * U + u) * TM + e];
          float dw = 0.f;
#pragma unroll
          for (int lm = l * l; lm < (l + 1) * (l + 1); ++lm) { dw += dVin[lm] * Y_s[lm * TM + e]; dYp[lm] += dVin[lm] * wv; }
          W0g[(l * U + u) * TM + e] = dw;      // in place: w0 -> dw0 (same thread read it above)
        }
      } else {
        float* dVp = dVprev + ((size_t)tile * U + u) * TP::DIN * TM;
#pragma unroll
        for (int c = 0; c < TP::DIN; ++c) dVp[c * TM + e] = dVin[c];
      }
#pragma unroll
      for (int lm = 0; lm < D::NSH; ++lm) DG[e * D::DGS + lm * D::CHU + ul] = dG[lm];
    }
    __syncthreads();
    {  // segmented sum of this pass's features: local f = lm*CHU+ul -> global lm*U + pass*CHU+ul
      const int* seg = reinterpret_cast<const int*>(sm + SM::oSeg);
      const bool contin = a.rowptr[c_s[0]] < es;
      float* carry = a.carry + (size_t)tile * D::F;
      segsum_items<D::FC, TM>(c_s, seg,
          [&](int ee, int f) { return DG[ee * D::DGS + f]; },
          [&](int centre, int f, bool first, float v) {
            const int lm = f / D::CHU, ul = f % D::CHU;
            const int fg = lm * U + pass * D::CHU + ul;
            if (first && contin) carry[fg] = v; else dgamma_out[(size_t)(centre - chunk_bounds(a).c0) * D::F + fg] = v;
          });
    }
    __syncthreads();
  }
  (void)t; (void)nvalid;
}

// reduce per-thread dY partials (CPH phases per edge) through region D and add to global dY
template <int L, bool ASSIGN>
__device__ __forceinline__ void dy_reduce_store(const ChunkArgs& a, int tile, float* sm, const float* dYp, const float* extra_s /*or nullptr*/) {
  using D = Dims<L>; using SM = Smem<L>; constexpr int TM = D::TM;
  const int t = threadIdx.x;
  float* P = sm + SM::oD;
  const int e = t % TM, uh = t / TM;
#pragma unroll
  for (int lm = 0; lm < D::NSH; ++lm) P[(uh * D::NSH + lm) * TM + e] = dYp[lm];
  __syncthreads();
  float* dYg = a.dY + (size_t)tile * D::NSH * TM;
  for (int i = t; i < D::NSH * TM; i += NT) {
    const int lm = i / TM, ee = i % TM;
    float v = 0.f;
#pragma unroll
    for (int h = 0; h < D::CPH; ++h) v += P[(h * D::NSH + lm) * TM + ee];
    if (extra_s) v += extra_s[i];
    if (ASSIGN) dYg[i] = v; else dYg[i] += v;
  }
  __syncthreads();
}

// phase 2 of layer kk for this tile: needs dGamma_kk complete.  Input: x^kk in region C.
//   w^kk = env_kk(x^kk) -> W_s ; dw rows -> IN[S..] ; dY partial -> DY_s (smem) ; then
//   dX(global) += env_kk^T dw
template <int L>
__device__ __forceinline__ void phase2(const ChunkArgs& a, const ModelW& w, int kk, int tile, float* sm) {
  using D = Dims<L>; using SM = Smem<L>; constexpr int TM = D::TM;
  const int t = threadIdx.x;
  float* IN = sm + SM::oIN; float* W_s = sm + SM::oA; float* C = sm + SM::oC; float* P = sm + SM::oD; float* DY_s = sm + SM::oDY;
  const float* Y_s = sm + SM::oY;
  const int* c_s = reinterpret_cast<const int*>(sm + SM::oMisc);
  gemm_tile_n<TM, D::ENVW>(C, S, w.layer[kk].env, D::ENVW, [&](int m, int n, float v) { W_s[m * D::WS + n] = v; });
  __syncthreads();
  {
    const int e = t % TM, uh = t / TM;
    const float* dgam = a.dgamma[kk] + (size_t)(c_s[e] - chunk_bounds(a).c0) * D::F;
    float dYp[D::NSH];
#pragma unroll
    for (int lm = 0; lm < D::NSH; ++lm) dYp[lm] = 0.f;
#pragma unroll 1
    for (int i = 0; i < D::CPT; ++i) {
      const int u = uh + D::CPH * i;
#pragma unroll
      for (int l = 0; l <= L; ++l) {
        const float wv = W_s[e * D::WS + l * U + u];
        float dw = 0.f;
#pragma unroll
        for (int lm = l * l; lm < (l + 1) * (l + 1); ++lm) {
          const float dg = dgam[lm * U + u] * w.inv_sqrt_n;
          dw += dg * Y_s[lm * TM + e];
          dYp[lm] += dg * wv;
        }
        IN[(S + l * U + u) * TM + e] = dw;
      }
    }
#pragma unroll
    for (int lm = 0; lm < D::NSH; ++lm) P[(uh * D::NSH + lm) * TM + e] = dYp[lm];
  }
  __syncthreads();
  for (int i = t; i < D::NSH * TM; i += NT) {
    const int lm = i / TM, ee = i % TM;
    float v = 0.f;
#pragma unroll
    for (int h = 0; h < D::CPH; ++h) v += P[(h * D::NSH + lm) * TM + ee];
    DY_s[i] = v;
  }
  float* dXg = a.dX + (size_t)tile * S * TM;
  gemm_tile<TM, 64>(IN + S * TM, D::ENVW, w.layer[kk].env_t, S, 0, [&](int m, int n, float v) { dXg[n * TM + m] += v; });
  __syncthreads();
}

// ============================================================================================
// T: last layer (kind 'A') forward + readout + backward phase 1
// ============================================================================================
template <int L, bool FIRST>
__global__ void __launch_bounds__(NT, Dims<L>::MINB) k_t(const __grid_constant__ ChunkArgs a, const __grid_constant__ ModelW w, const int k) {
  using D = Dims<L>; using SM = Smem<L>; using TP = tpgen::TP<L, 'A'>; constexpr int TM = D::TM;
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, tile = blockIdx.x;
  const ChunkBounds cb = chunk_bounds(a);
  const int es = cb.e0 + tile * TM;
  if (es >= cb.e1) return;
  const int nvalid = min(TM, cb.e1 - es);
  float* IN = sm + SM::oIN; float* A = sm + SM::oA; float* B = sm + SM::oB; float* C = sm + SM::oC; float* Dd = sm + SM::oD;
  const float* u_s = sm + SM::oU; const float* Y_s = sm + SM::oY;
  const int* c_s = reinterpret_cast<const int*>(sm + SM::oMisc);
  const int* zz_s = c_s + TM;
  float* e_s = sm + SM::oMisc + 2 * TM;
  const LayerW& lw = w.layer[k];
  load_rows<TM>(IN, a.X[k] + (size_t)tile * S * TM, S);
  if (t < TM) edge_geom<L>(a, w, es, nvalid, sm);
  __syncthreads();
  seg_setup<L>(sm, nvalid);
  {  // scalar tensor-product outputs s -> IN rows S..
    const int e = t % TM, uh = t / TM;
    const float* gam = a.gamma[k] + (size_t)(c_s[e] - chunk_bounds(a).c0) * D::F;
#pragma unroll 1
    for (int i = 0; i < D::CPT; ++i) {
      const int u = uh + D::CPH * i;
      float Vin[TP::DIN], G[D::NSH], s[TP::N0];
      load_vin<L, FIRST, TP::DIN>(a, tile, k, e, u, Y_s, Vin);
#pragma unroll
      for (int lm = 0; lm < D::NSH; ++lm) G[lm] = gam[lm * U + u];
      TP::template fwd<U>(Vin, G, lw.omega + u, nullptr, s);
#pragma unroll
      for (int q = 0; q < TP::N0; ++q) IN[(S + q * U + u) * TM + e] = s[q];
    }
  }
  __syncthreads();
  mlp_fwd_keep<L>(lw.mlp, D::SIN, lw.mlp.w[0], sm, false, 0);
  // GEMM3: m (pre-envelope) -> A ; x^n -> IN rows S..S+63 (s rows are dead)
  float* XN = IN + S * TM;
  gemm_tile<TM, 64>(C, H, lw.mlp.w[2], S, 0, [&](int m, int n, float v) {
    A[n * TM + m] = v;
    XN[n * TM + m] = lw.a * IN[n * TM + m] + lw.b * v * u_s[m];
  });
  __syncthreads();
  // readout hidden: r1 -> C rows 0..31, act' -> C rows 32..63
  gemm_tile<TM, 32>(XN, S, w.ro0, R, 0, [&](int m, int n, float v) { float d; C[n * TM + m] = silu_act(v, d); C[(R + n) * TM + m] = d; });
  __syncthreads();
  if (t < TM) {
    float ee = 0.f;
    const float ge = w.gscale[zz_s[t] & 255];
#pragma unroll 4
    for (int q = 0; q < R; ++q) {
      const float wq = __ldg(w.ro1 + q);
      ee += wq * C[q * TM + t];
      C[(R + q) * TM + t] *= ge * wq;                 // dz_readout
    }
    e_s[t] = ee;
    if (t < nvalid && a.edge_energy) a.edge_energy[es + t] = ee;
  }
  __syncthreads();
  {  // E_i raw sums (double): one thread per centre run, edges in order (deterministic)
    const int* seg = reinterpret_cast<const int*>(sm + SM::oSeg);
    const int nseg = seg[TM + 1];
    const bool contin = a.rowptr[c_s[0]] < es;
    for (int sgm = t; sgm < nseg; sgm += NT) {
      double acc = 0.0;
      for (int e = seg[sgm]; e < seg[sgm + 1]; ++e) acc += (double)e_s[e];
      if (sgm == 0 && contin) a.ecarry[tile] = acc; else a.esum[c_s[seg[sgm]]] = acc;
    }
  }
  // dx^n = ro0^T dz ; then residual / envelope split (fused epilogue)
  float* dXg = a.dX + (size_t)tile * S * TM;
  gemm_tile<TM, 64>(C + R * TM, R, w.ro0_t, S, 0, [&](int m, int n, float v) {
    const float dxt = lw.b * v;
    dXg[n * TM + m] = lw.a * v;                       // partial dx^{n-1}
    const float mv = A[n * TM + m];
    A[n * TM + m] = dxt * mv;                         // -> du
    IN[n * TM + m] = dxt * u_s[m];                    // dm
  });
  __syncthreads();
  if (t < TM) {
    float s = 0.f;
#pragma unroll 8
    for (int q = 0; q < S; ++q) s += A[q * TM + t];
    a.du[(size_t)tile * TM + t] = s;
  }
  mlp_bwd_hidden<L>(lw.mlp, sm);
  // dIN = W0^T dz1 : columns < S add into dX, columns >= S are ds (-> IN rows S..)
  gemm_tile_n<TM, D::SIN>(B, H, lw.mlp.wt[0], D::SIN, [&](int m, int n, float v) {
    if (n < S) dXg[n * TM + m] += v; else IN[n * TM + m] = v;
  });
  __syncthreads();
  float dYp[D::NSH];
  tp_backward<L, 'A', FIRST, false>(a, lw, tile, k, es, nvalid, sm, nullptr, FIRST ? nullptr : a.dV[k & 1], a.dgamma[k], dYp);
  if (FIRST) {
    dy_reduce_store<L, true>(a, tile, sm, dYp, nullptr);
  } else {
    float* dYg = a.dY + (size_t)tile * D::NSH * TM;
    for (int i = t; i < D::NSH * TM; i += NT) dYg[i] = 0.f;
  }
}

// ============================================================================================
// BK: backward of layer k (k < nl-1): phase 2 of layer k+1, then phase 1 of layer k
// ============================================================================================
template <int L, char KIND, bool FIRST>
__global__ void __launch_bounds__(NT, Dims<L>::MINB) k_bk(const __grid_constant__ ChunkArgs a, const __grid_constant__ ModelW w, const int k) {
  using D = Dims<L>; using SM = Smem<L>; using TP = tpgen::TP<L, KIND>; constexpr int TM = D::TM;
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, tile = blockIdx.x;
  const ChunkBounds cb = chunk_bounds(a);
  const int es = cb.e0 + tile * TM;
  if (es >= cb.e1) return;
  const int nvalid = min(TM, cb.e1 - es);
  float* IN = sm + SM::oIN; float* A = sm + SM::oA; float* B = sm + SM::oB; float* C = sm + SM::oC;
  const float* u_s = sm + SM::oU; const float* Y_s = sm + SM::oY; float* DY_s = sm + SM::oDY;
  const int* c_s = reinterpret_cast<const int*>(sm + SM::oMisc);
  const LayerW& lw = w.layer[k];
  load_rows<TM>(C, a.X[k + 1] + (size_t)tile * S * TM, S);
  if (t < TM) edge_geom<L>(a, w, es, nvalid, sm);
  __syncthreads();
  seg_setup<L>(sm, nvalid);
  phase2<L>(a, w, k + 1, tile, sm);                  // dX(global) now holds the complete dx^{k+1}
  // ---- phase 1 of layer k: recompute forward
  load_rows<TM>(IN, a.X[k] + (size_t)tile * S * TM, S);
  __syncthreads();
  {
    const int e = t % TM, uh = t / TM;
    const float* gam = a.gamma[k] + (size_t)(c_s[e] - chunk_bounds(a).c0) * D::F;
#pragma unroll 1
    for (int i = 0; i < D::CPT; ++i) {
      const int u = uh + D::CPH * i;
      float Vin[TP::DIN], G[D::NSH], Vout[TP::DOUT], s[TP::N0];
      load_vin<L, FIRST, TP::DIN>(a, tile, k, e, u, Y_s, Vin);
#pragma unroll
      for (int lm = 0; lm < D::NSH; ++lm) G[lm] = gam[lm * U + u];
      tpgen::TP<L, 'A'>::template fwd<U>(Vin, G, nullptr, nullptr, s);   // scalar paths only (same order as KIND's scalar paths)
      (void)Vout;
#pragma unroll
      for (int q = 0; q < TP::N0; ++q) IN[(S + q * U + u) * TM + e] = s[q];
    }
  }
  __syncthreads();
  mlp_fwd_keep<L>(lw.mlp, D::SIN, lw.mlp.w[0], sm, false, 0);
  float* dXg = a.dX + (size_t)tile * S * TM;
  gemm_tile<TM, 64>(C, H, lw.mlp.w[2], S, 0, [&](int m, int n, float v) {
    const float dxn = dXg[n * TM + m];
    const float dxt = lw.b * dxn;
    dXg[n * TM + m] = lw.a * dxn;
    A[n * TM + m] = dxt * v;
    IN[n * TM + m] = dxt * u_s[m];
  });
  __syncthreads();
  if (t < TM) {
    float s = 0.f;
#pragma unroll 8
    for (int q = 0; q < S; ++q) s += A[q * TM + t];
    a.du[(size_t)tile * TM + t] += s;
  }
  mlp_bwd_hidden<L>(lw.mlp, sm);
  gemm_tile_n<TM, D::SIN>(B, H, lw.mlp.wt[0], D::SIN, [&](int m, int n, float v) {
    if (n < S) dXg[n * TM + m] += v; else IN[n * TM + m] = v;
  });
  __syncthreads();
  float dYp[D::NSH];
  tp_backward<L, KIND, FIRST, true>(a, lw, tile, k, es, nvalid, sm, a.dV[(k + 1) & 1], FIRST ? nullptr : a.dV[k & 1], a.dgamma[k], dYp);
  if (FIRST) {
    dy_reduce_store<L, false>(a, tile, sm, dYp, DY_s);
  } else {
    float* dYg = a.dY + (size_t)tile * D::NSH * TM;
    for (int i = t; i < D::NSH * TM; i += NT) dYg[i] += DY_s[i];
  }
}

// ============================================================================================
// B0: phase 2 of layer 0, embed / two-body / geometry backward, force + virial accumulation
// ============================================================================================
template <int L>
__global__ void __launch_bounds__(NT, Dims<L>::MINB) k_b0(const __grid_constant__ ChunkArgs a, const __grid_constant__ ModelW w) {
  using D = Dims<L>; using SM = Smem<L>; constexpr int TM = D::TM;
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, tile = blockIdx.x;
  const ChunkBounds cb = chunk_bounds(a);
  const int es = cb.e0 + tile * TM;
  if (es >= cb.e1) return;
  const int nvalid = min(TM, cb.e1 - es);
  float* IN = sm + SM::oIN; float* A = sm + SM::oA; float* B = sm + SM::oB; float* C = sm + SM::oC; float* Dd = sm + SM::oD;
  const float* u_s = sm + SM::oU; float* DY_s = sm + SM::oDY;
  const int* c_s = reinterpret_cast<const int*>(sm + SM::oMisc);
  float* du_s = sm + SM::oMisc + 2 * TM;
  load_rows<TM>(C, a.X[0] + (size_t)tile * S * TM, S);
  Geom g;
  if (t < TM) g = edge_geom<L>(a, w, es, nvalid, sm);
  __syncthreads();
  seg_setup<L>(sm, nvalid);
  phase2<L>(a, w, 0, tile, sm);
  // dx0 += emb^T dw0
  float* dXg = a.dX + (size_t)tile * S * TM;
  load_rows<TM>(IN + S * TM, a.W0 + (size_t)tile * D::ENVW * TM, D::ENVW);
  __syncthreads();
  gemm_tile<TM, 64>(IN + S * TM, D::ENVW, w.emb_t, S, 0, [&](int m, int n, float v) { dXg[n * TM + m] += v; });
  __syncthreads();
  // recompute the two-body MLP: Bessel rows -> IN rows 0..B-1
  float bes[MAXB], dbes[MAXB];
  if (t < TM) {
    const float pref = sqrtf(2.0f / g.rc);
    const float xr = g.r / g.rc;
    for (int n = 0; n < w.B; ++n) {
      const float kn = (float)(n + 1) * 3.14159265358979323846f;
      float sn, cs;
      sincosf(kn * xr, &sn, &cs);
      bes[n] = pref * sn / g.r;
      dbes[n] = pref * (kn / g.rc * cs / g.r - sn / (g.r * g.r));
      IN[n * TM + t] = bes[n] * g.u;
    }
  }
  __syncthreads();
  mlp_fwd_keep<L>(w.two, w.B, w.two.w[0] + 2 * w.T * H, sm, true, w.T);
  gemm_tile<TM, 64>(C, H, w.two.w[2], S, 0, [&](int m, int n, float v) {
    const float dx0 = dXg[n * TM + m];
    A[n * TM + m] = dx0 * v;                          // -> du (x0 = m0*u)
    IN[n * TM + m] = dx0 * u_s[m];                    // dm0
  });
  __syncthreads();
  if (t < TM) {
    float s = 0.f;
#pragma unroll 8
    for (int q = 0; q < S; ++q) s += A[q * TM + t];
    du_s[t] = s + a.du[(size_t)tile * TM + t];
  }
  mlp_bwd_hidden<L>(w.two, sm);
  // per edge: d(bessel*u) = W0[2T+n,:] . dz1 ; assemble dE/dr, angular part, scatter
  float gx = 0.f, gy = 0.f, gz = 0.f;
  if (t < TM) {
    const float* w0b = w.two.w[0] + 2 * w.T * H;
    float dr = du_s[t] * g.dudr;
    for (int n = 0; n < w.B; ++n) {
      float acc = 0.f;
#pragma unroll 8
      for (int h = 0; h < H; ++h) acc += __ldg(w0b + n * H + h) * B[h * TM + t];
      dr += acc * (dbes[n] * g.u + bes[n] * g.dudr);
    }
    float dYt[D::NSH];
    const float* dYg = a.dY + (size_t)tile * D::NSH * TM;
#pragma unroll
    for (int lm = 0; lm < D::NSH; ++lm) dYt[lm] = dYg[lm * TM + t] + DY_s[lm * TM + t];
    float qx, qy, qz;
    sph_harm_vjp<L>(g.x, g.y, g.z, dYt, qx, qy, qz);
    const float nq = g.x * qx + g.y * qy + g.z * qz;
    const float ir = 1.0f / g.r;
    gx = dr * g.x + (qx - g.x * nq) * ir;
    gy = dr * g.y + (qy - g.y * nq) * ir;
    gz = dr * g.z + (qz - g.z * nq) * ir;
    if (t >= nvalid) { gx = gy = gz = 0.f; }
    // stage for the centre-side segmented sum and the virial
    A[0 * TM + t] = gx; A[1 * TM + t] = gy; A[2 * TM + t] = gz;
    if (t < nvalid) {
      const int e = es + t;
      if (a.edge_grad) { a.edge_grad[3 * (size_t)e + 0] = gx; a.edge_grad[3 * (size_t)e + 1] = gy; a.edge_grad[3 * (size_t)e + 2] = gz; }
      // F_j -= g_e  (ghost neighbours included: newton on)
      const int j = a.edge_j[e];
      atomicAdd(a.facc + 3 * (size_t)j + 0, (unsigned long long)__double2ll_rn(-(double)gx * FIX_SCALE));
      atomicAdd(a.facc + 3 * (size_t)j + 1, (unsigned long long)__double2ll_rn(-(double)gy * FIX_SCALE));
      atomicAdd(a.facc + 3 * (size_t)j + 2, (unsigned long long)__double2ll_rn(-(double)gz * FIX_SCALE));
      // virial W = -sum r (x) g, symmetrised: xx yy zz xy xz yz
      const float rx = g.x * g.r, ry = g.y * g.r, rz = g.z * g.r;
      Dd[0 * TM + t] = -rx * gx; Dd[1 * TM + t] = -ry * gy; Dd[2 * TM + t] = -rz * gz;
      Dd[3 * TM + t] = -0.5f * (rx * gy + ry * gx); Dd[4 * TM + t] = -0.5f * (rx * gz + rz * gx); Dd[5 * TM + t] = -0.5f * (ry * gz + rz * gy);
    } else {
#pragma unroll
      for (int q = 0; q < 6; ++q) Dd[q * TM + t] = 0.f;
    }
  }
  __syncthreads();
  {
    const int* seg = reinterpret_cast<const int*>(sm + SM::oSeg);
    const int nseg = seg[TM + 1];
    // F_i += sum of g_e over each centre run (fixed order), one atomic per (centre, component)
    for (int w2 = t; w2 < 3 * nseg; w2 += NT) {
      const int q = w2 % 3, sgm = w2 / 3;
      double acc = 0.0;
      for (int e = seg[sgm]; e < seg[sgm + 1]; ++e) acc += (double)A[q * TM + e];
      atomicAdd(a.facc + 3 * (size_t)a.ilist[c_s[seg[sgm]]] + q, (unsigned long long)__double2ll_rn(acc * FIX_SCALE));
    }
    // virial: exact integer (fixed-point) warp reduction, one atomic per warp and component
    if (a.vacc && t >= NT - TM) {
      const int e = t - (NT - TM);
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        long long v = __double2ll_rn((double)Dd[q * TM + e] * VIR_SCALE);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((t & 31) == 0) atomicAdd(a.vacc + q, (unsigned long long)v);
      }
    }
  }
}

// ============================================================================================
// carry fix-up: centres whose CSR row spans several tiles.  One block per tile; the tile in
// which the row STARTS adds the carries of the following tiles in tile order (deterministic).
// ============================================================================================
template <int TM>
__global__ void k_fixup(const int* __restrict__ edge_c, const int* __restrict__ rowptr, int e0, int e1, int c0, int ntiles, int NF,
                        float* __restrict__ out, const float* __restrict__ carry, const int* __restrict__ plan, int ci) {
  if (plan) { e0 = plan[3 * ci]; e1 = plan[3 * ci + 1]; c0 = plan[3 * ci + 2]; ntiles = (e1 - e0 + TM - 1) / TM; }
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {     // grid-stride: the grid need not know the tile count
    const int es = e0 + tile * TM;
    const int ee = min(es + TM, e1);
    const int c = edge_c[ee - 1];
    const int rb = rowptr[c], re = rowptr[c + 1];
    if (rb < es || re <= ee) continue;              // row does not start here, or ends here
    for (int f = threadIdx.x; f < NF; f += blockDim.x) {
      float acc = out[(size_t)(c - c0) * NF + f];
      for (int t2 = tile + 1; t2 < ntiles && e0 + t2 * TM < re; ++t2) acc += carry[(size_t)t2 * NF + f];
      out[(size_t)(c - c0) * NF + f] = acc;
    }
  }
}
template <int TM>
__global__ void k_fixup_e(const int* __restrict__ edge_c, const int* __restrict__ rowptr, int e0, int e1, int ntiles,
                          double* __restrict__ esum, const double* __restrict__ ecarry, const int* __restrict__ plan, int ci) {
  if (plan) { e0 = plan[3 * ci]; e1 = plan[3 * ci + 1]; ntiles = (e1 - e0 + TM - 1) / TM; }
  for (int tile = blockIdx.x * blockDim.x + threadIdx.x; tile < ntiles; tile += gridDim.x * blockDim.x) {
    const int es = e0 + tile * TM;
    const int ee = min(es + TM, e1);
    const int c = edge_c[ee - 1];
    const int rb = rowptr[c], re = rowptr[c + 1];
    if (rb < es || re <= ee) continue;
    double acc = esum[c];
    for (int t2 = tile + 1; t2 < ntiles && e0 + t2 * TM < re; ++t2) acc += ecarry[t2];
    esum[c] = acc;
  }
}

}  // namespace alg
