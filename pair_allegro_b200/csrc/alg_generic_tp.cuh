// Tensor-product kernels of the width-generic pipeline (alg_generic.cuh): one thread per (edge, channel), the generated
// fully unrolled Clebsch-Gordan code of tp_gen.cuh with a channel-major omega table (stride 1).
#pragma once
#include "alg_generic.cuh"
#include "tp_gen.cuh"

namespace alg {

template <int L, char KIND, bool FIRST>
__global__ void __launch_bounds__(128) k_gen_tp_fwd(const GenTp a) {
  using TP = tpgen::TP<L, KIND>;
  constexpr int NSH = (L + 1) * (L + 1);
  static_assert(!FIRST || TP::DIN == NSH, "first layer: V^0 = w0 (x) Y");
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)a.n * a.U) return;
  const int q = (int)(idx / a.U), u = (int)(idx % a.U);
  const size_t NU = (size_t)a.n * a.U;                  // V / dV / dG_e are component-major: [component][edge][channel] (coalesced)
  float Vin[TP::DIN], G[NSH], Vout[TP::DOUT], s[TP::N0];
  if (FIRST) {
#pragma unroll
    for (int lm = 0; lm < NSH; ++lm) Vin[lm < TP::DIN ? lm : 0] = a.vin[(size_t)q * a.envw + lsel(lm) * a.U + u] * a.Y[(size_t)q * NSH + lm];
  } else {
#pragma unroll
    for (int c = 0; c < TP::DIN; ++c) Vin[c] = a.vin[(size_t)c * NU + idx];
  }
  const float* gam = a.gamma + (size_t)(a.edge_c[a.e0 + q] - a.c0) * NSH * a.U;
#pragma unroll
  for (int lm = 0; lm < NSH; ++lm) G[lm] = gam[lm * a.U + u];
#pragma unroll
  for (int c = 0; c < TP::DOUT; ++c) Vout[c] = 0.f;
  TP::template fwd<1>(Vin, G, a.omega_t + (size_t)u * TP::NPATH, Vout, s);
  if (a.vout) {
#pragma unroll
    for (int c = 0; c < TP::DOUT; ++c) a.vout[(size_t)c * NU + idx] = Vout[c];
  }
#pragma unroll
  for (int i = 0; i < TP::N0; ++i) a.IN[(size_t)q * a.ldin + a.S + i * a.U + u] = s[i];
}

template <int L, char KIND, bool FIRST>
__global__ void __launch_bounds__(128) k_gen_tp_bwd(const GenTp a) {
  using TP = tpgen::TP<L, KIND>;
  constexpr int NSH = (L + 1) * (L + 1);
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)a.n * a.U) return;
  const int q = (int)(idx / a.U), u = (int)(idx % a.U);
  const size_t NU = (size_t)a.n * a.U;
  float Vin[TP::DIN], G[NSH], dVout[TP::DOUT], ds[TP::N0], dVin[TP::DIN], dG[NSH];
  if (FIRST) {
#pragma unroll
    for (int lm = 0; lm < NSH; ++lm) Vin[lm < TP::DIN ? lm : 0] = a.vin[(size_t)q * a.envw + lsel(lm) * a.U + u] * a.Y[(size_t)q * NSH + lm];
  } else {
#pragma unroll
    for (int c = 0; c < TP::DIN; ++c) Vin[c] = a.vin[(size_t)c * NU + idx];
  }
  const float* gam = a.gamma + (size_t)(a.edge_c[a.e0 + q] - a.c0) * NSH * a.U;
#pragma unroll
  for (int lm = 0; lm < NSH; ++lm) G[lm] = gam[lm * a.U + u];
#pragma unroll
  for (int c = 0; c < TP::DOUT; ++c) dVout[c] = a.dvout ? a.dvout[(size_t)c * NU + idx] : 0.f;
#pragma unroll
  for (int i = 0; i < TP::N0; ++i) ds[i] = a.IN[(size_t)q * a.ldin + a.S + i * a.U + u];
  TP::template bwd<1>(Vin, G, a.omega_t + (size_t)u * TP::NPATH, dVout, ds, dVin, dG);
#pragma unroll
  for (int c = 0; c < TP::DIN; ++c) a.dvin[(size_t)c * NU + idx] = dVin[c];
#pragma unroll
  for (int lm = 0; lm < NSH; ++lm) a.dge[(size_t)lm * NU + idx] = dG[lm];
}

template <int L, char KIND, bool FIRST>
cudaError_t gen_tp_run(bool backward, const GenTp& a, cudaStream_t st) {
  const long total = (long)a.n * a.U;
  if (total == 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((total + 127) / 128);
  if (backward) k_gen_tp_bwd<L, KIND, FIRST><<<blocks, 128, 0, st>>>(a);
  else k_gen_tp_fwd<L, KIND, FIRST><<<blocks, 128, 0, st>>>(a);
  return cudaGetLastError();
}

template <int L> cudaError_t gen_tp_launch_impl(char kind, bool first, bool backward, const GenTp& a, cudaStream_t st) {
  if (first) {
    if (kind == 'A') return gen_tp_run<L, 'A', true>(backward, a, st);
    if (kind == 'B') return gen_tp_run<L, 'B', true>(backward, a, st);
    if (kind == 'C') return gen_tp_run<L, 'C', true>(backward, a, st);
  } else {
    if (kind == 'A') return gen_tp_run<L, 'A', false>(backward, a, st);
    if (kind == 'D') return gen_tp_run<L, 'D', false>(backward, a, st);
  }
  return cudaErrorInvalidValue;
}
template <int L> GenTpDims gen_tp_dims_impl(char kind) {
  switch (kind) {
    case 'A': return {tpgen::TP<L, 'A'>::DIN, tpgen::TP<L, 'A'>::DOUT, tpgen::TP<L, 'A'>::NPATH, tpgen::TP<L, 'A'>::N0};
    case 'B': return {tpgen::TP<L, 'B'>::DIN, tpgen::TP<L, 'B'>::DOUT, tpgen::TP<L, 'B'>::NPATH, tpgen::TP<L, 'B'>::N0};
    case 'C': return {tpgen::TP<L, 'C'>::DIN, tpgen::TP<L, 'C'>::DOUT, tpgen::TP<L, 'C'>::NPATH, tpgen::TP<L, 'C'>::N0};
    case 'D': return {tpgen::TP<L, 'D'>::DIN, tpgen::TP<L, 'D'>::DOUT, tpgen::TP<L, 'D'>::NPATH, tpgen::TP<L, 'D'>::N0};
  }
  return {0, 0, 0, 0};
}

#define ALG_DEFINE_GENERIC_TP(LV)                                                                                             \
  cudaError_t gen_tp_launch_L##LV(char kind, bool first, bool backward, const GenTp& a, cudaStream_t st) {                    \
    return gen_tp_launch_impl<LV>(kind, first, backward, a, st);                                                              \
  }                                                                                                                           \
  GenTpDims gen_tp_dims_L##LV(char kind) { return gen_tp_dims_impl<LV>(kind); }

}  // namespace alg
