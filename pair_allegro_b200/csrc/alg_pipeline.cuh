// Per-l_max pipeline instantiation: one translation unit per L (alg_inst_L{1,2,3}.cu) so the
// unrolled tensor-product code compiles in parallel.
#pragma once
#include <vector>

#include "allegro_kernels.cuh"
#include "allegro_kernels_tc.cuh"

namespace alg {

struct PipelineInfo {
  int L, TM, NSH, ENVW, F;
  int vdim[3];        // components per channel of V^k (k = 1..nl-1), index by k
  int dvdim;          // max over vdim
  int vstride[3];     // storage components per channel of V^k (tensor-core pipeline: padded, see vpad)
  int dvstride;
  size_t smem_bytes;
};

// optional per-kernel CUDA-event timing (option profile=1) and launch counting
enum { KID_F0 = 0, KID_FK = 1, KID_T = 2, KID_BK = 3, KID_B0 = 4, KID_FIXUP = 5, KID_FUSED = 6, KID_COUNT = 7 };
struct Prof {
  bool on = false;
  std::vector<cudaEvent_t> ev;
  std::vector<int> ids;
  size_t used = 0;
  long launches = 0;
  void begin(int id, cudaStream_t st) {
    ++launches;
    if (!on) return;
    while (ev.size() < used + 2) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); }
    cudaEventRecord(ev[used], st);
    ids.push_back(id);
  }
  void end(cudaStream_t st) {
    if (!on) return;
    cudaEventRecord(ev[used + 1], st);
    used += 2;
  }
  void reset() { used = 0; ids.clear(); launches = 0; }
};

struct Pipeline {
  PipelineInfo (*info)(int nl);
  cudaError_t (*init)();                                                       // opt-in shared memory
  cudaError_t (*run_chunk)(const ChunkArgs& a, const ModelW& w, const TcW* tw, int ntiles, cudaStream_t st, Prof* prof);
  // fused persistent kernel over centre-aligned tiles (tensor-core pipeline only; nullptr otherwise)
  int (*fused_grid)(int nl);                                                   // resident CTAs of the whole device (0 = unsupported)
  cudaError_t (*run_fused)(const ChunkArgs& a, const ModelW& w, const TcW* tw, const FusedPlan& plan, int grid, cudaStream_t st, Prof* prof);
};

const Pipeline* get_pipeline(int L);

#ifdef ALG_PIPELINE_IMPL
template <int L> PipelineInfo info_impl(int nl) {
  using D = Dims<L>;
  PipelineInfo p{};
  p.L = L; p.TM = D::TM; p.NSH = D::NSH; p.ENVW = D::ENVW; p.F = D::F;
  p.vdim[0] = D::NSH; p.vdim[1] = p.vdim[2] = 0;
  if (nl == 2) p.vdim[1] = tpgen::TP<L, 'B'>::DOUT;
  if (nl == 3) { p.vdim[1] = tpgen::TP<L, 'C'>::DOUT; p.vdim[2] = tpgen::TP<L, 'D'>::DOUT; }
  p.dvdim = p.vdim[1] > p.vdim[2] ? p.vdim[1] : p.vdim[2];
  for (int q = 0; q < 3; ++q) p.vstride[q] = p.vdim[q];
  p.dvstride = p.dvdim;
  p.smem_bytes = Smem<L>::BYTES;
  return p;
}

template <int L> cudaError_t init_impl() {
  const int b = (int)Smem<L>::BYTES;
  cudaError_t e;
#define ALG_SET(kern) \
  if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, b)) != cudaSuccess) return e;
  ALG_SET((k_f0<L>));
  ALG_SET((k_fk<L, 'B', true>));
  ALG_SET((k_fk<L, 'C', true>));
  ALG_SET((k_fk<L, 'D', false>));
  ALG_SET((k_t<L, true>));
  ALG_SET((k_t<L, false>));
  ALG_SET((k_bk<L, 'B', true>));
  ALG_SET((k_bk<L, 'C', true>));
  ALG_SET((k_bk<L, 'D', false>));
  ALG_SET((k_b0<L>));
#undef ALG_SET
  return cudaSuccess;
}

template <int L> cudaError_t run_chunk_impl(const ChunkArgs& a, const ModelW& w, const TcW*, int ntiles, cudaStream_t st, Prof* pf) {
  using D = Dims<L>;
  constexpr int TM = D::TM;
  const size_t sm = Smem<L>::BYTES;
  const dim3 g(ntiles), b(NT);
#define ALG_RUN(kid, ...) do { pf->begin(kid, st); __VA_ARGS__; pf->end(st); } while (0)
  auto fix = [&](float* out) {
    ALG_RUN(KID_FIXUP, (k_fixup<TM><<<g, 128, 0, st>>>(a.edge_c, a.rowptr, a.e0, a.e1, a.c0, ntiles, D::F, out, a.carry, a.plan, a.ci)));
  };
  ALG_RUN(KID_F0, (k_f0<L><<<g, b, sm, st>>>(a, w)));
  fix(a.gamma[0]);
  const int nl = w.nl;
  if (nl == 1) {
    ALG_RUN(KID_T, (k_t<L, true><<<g, b, sm, st>>>(a, w, 0)));
  } else if (nl == 2) {
    ALG_RUN(KID_FK, (k_fk<L, 'B', true><<<g, b, sm, st>>>(a, w, 0)));
    fix(a.gamma[1]);
    ALG_RUN(KID_T, (k_t<L, false><<<g, b, sm, st>>>(a, w, 1)));
  } else {
    ALG_RUN(KID_FK, (k_fk<L, 'C', true><<<g, b, sm, st>>>(a, w, 0)));
    fix(a.gamma[1]);
    ALG_RUN(KID_FK, (k_fk<L, 'D', false><<<g, b, sm, st>>>(a, w, 1)));
    fix(a.gamma[2]);
    ALG_RUN(KID_T, (k_t<L, false><<<g, b, sm, st>>>(a, w, 2)));
  }
  fix(a.dgamma[nl - 1]);
  ALG_RUN(KID_FIXUP, (k_fixup_e<TM><<<(ntiles + 127) / 128, 128, 0, st>>>(a.edge_c, a.rowptr, a.e0, a.e1, ntiles, a.esum, a.ecarry, a.plan, a.ci)));
  if (nl == 2) {
    ALG_RUN(KID_BK, (k_bk<L, 'B', true><<<g, b, sm, st>>>(a, w, 0)));
    fix(a.dgamma[0]);
  } else if (nl == 3) {
    ALG_RUN(KID_BK, (k_bk<L, 'D', false><<<g, b, sm, st>>>(a, w, 1)));
    fix(a.dgamma[1]);
    ALG_RUN(KID_BK, (k_bk<L, 'C', true><<<g, b, sm, st>>>(a, w, 0)));
    fix(a.dgamma[0]);
  }
  ALG_RUN(KID_B0, (k_b0<L><<<g, b, sm, st>>>(a, w)));
#undef ALG_RUN
  return cudaGetLastError();
}

// ---- tensor-core pipeline (l_max = 1) ------------------------------------------------------
template <int L> PipelineInfo info_tc_impl(int nl) {
  PipelineInfo p = info_impl<L>(nl);
  p.TM = DimsTC<L>::TM;
  p.smem_bytes = SmemTC<L>::BYTES;
  for (int q = 0; q < 3; ++q) p.vstride[q] = vpad(p.vdim[q]);
  p.dvstride = vpad(p.dvdim);
  return p;
}
template <int L> cudaError_t init_tc_impl() {
  const int b = (int)SmemTC<L>::BYTES;
  cudaError_t e;
#define ALG_SET(kern) \
  if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, b)) != cudaSuccess) return e; \
  if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  ALG_SET((k_f0_tc<L>));
  ALG_SET((k_fk_tc<L, 'B', true>));
  ALG_SET((k_fk_tc<L, 'C', true>));
  ALG_SET((k_fk_tc<L, 'D', false>));
  ALG_SET((k_t_tc<L, true>));
  ALG_SET((k_t_tc<L, false>));
  ALG_SET((k_bk_tc<L, 'B', true>));
  ALG_SET((k_bk_tc<L, 'C', true>));
  ALG_SET((k_bk_tc<L, 'D', false>));
  ALG_SET((k_b0_tc<L>));
  ALG_SET((k_fused_tc<L, 1>));
  ALG_SET((k_fused_tc<L, 2>));
  ALG_SET((k_fused_tc<L, 3>));
#undef ALG_SET
  return cudaSuccess;
}
template <int L> int fused_grid_tc_impl(int nl) {
  int dev = 0, sms = 0, smem_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) != cudaSuccess) return 0;
  // resident CTAs per SM: shared memory (+1 KB the driver reserves per CTA), at most 2 (tensor memory: 256 of 512 columns per CTA;
  // registers: 128 x 256 threads).  The kernel pulls tiles from a queue, so a smaller residency at run time only costs balance.
  int per_sm = (int)(smem_sm / (SmemTC<L>::BYTES + 1024));
  if (per_sm > 2) per_sm = 2;
  (void)nl;
  return per_sm < 1 ? 0 : sms * per_sm;
}
template <int L> cudaError_t run_fused_tc_impl(const ChunkArgs& a, const ModelW& w, const TcW* twp, const FusedPlan& plan, int grid, cudaStream_t st, Prof* pf) {
  const size_t sm = SmemTC<L>::BYTES;
  pf->begin(KID_FUSED, st);
  if (w.nl == 1) k_fused_tc<L, 1><<<grid, NT, sm, st>>>(a, w, *twp, plan);
  else if (w.nl == 2) k_fused_tc<L, 2><<<grid, NT, sm, st>>>(a, w, *twp, plan);
  else k_fused_tc<L, 3><<<grid, NT, sm, st>>>(a, w, *twp, plan);
  pf->end(st);
  return cudaGetLastError();
}
template <int L> cudaError_t run_chunk_tc_impl(const ChunkArgs& a, const ModelW& w, const TcW* twp, int ntiles, cudaStream_t st, Prof* pf) {
  using D = DimsTC<L>;
  constexpr int TM = D::TM;
  const size_t sm = SmemTC<L>::BYTES;
  // the phase kernels are persistent over the tiles of the chunk: one wave of resident CTAs (2 per SM; 1 for l_max = 3)
  static const int resident = fused_grid_tc_impl<L>(0);
  const dim3 g(ntiles), b(NT), gp(resident > 0 && resident < ntiles ? resident : ntiles), gf(ntiles < 2368 ? ntiles : 2368);
  const TcW& tw = *twp;
#define ALG_RUN(kid, ...) do { pf->begin(kid, st); __VA_ARGS__; pf->end(st); } while (0)
  auto fix = [&](float* out) {
    ALG_RUN(KID_FIXUP, (k_fixup<TM><<<gf, 128, 0, st>>>(a.edge_c, a.rowptr, a.e0, a.e1, a.c0, ntiles, D::F, out, a.carry, a.plan, a.ci)));
  };
  ALG_RUN(KID_F0, (k_f0_tc<L><<<gp, b, sm, st>>>(a, w, tw)));
  fix(a.gamma[0]);
  const int nl = w.nl;
  if (nl == 1) {
    ALG_RUN(KID_T, (k_t_tc<L, true><<<gp, b, sm, st>>>(a, w, tw, 0)));
  } else if (nl == 2) {
    ALG_RUN(KID_FK, (k_fk_tc<L, 'B', true><<<gp, b, sm, st>>>(a, w, tw, 0)));
    fix(a.gamma[1]);
    ALG_RUN(KID_T, (k_t_tc<L, false><<<gp, b, sm, st>>>(a, w, tw, 1)));
  } else {
    ALG_RUN(KID_FK, (k_fk_tc<L, 'C', true><<<gp, b, sm, st>>>(a, w, tw, 0)));
    fix(a.gamma[1]);
    ALG_RUN(KID_FK, (k_fk_tc<L, 'D', false><<<gp, b, sm, st>>>(a, w, tw, 1)));
    fix(a.gamma[2]);
    ALG_RUN(KID_T, (k_t_tc<L, false><<<gp, b, sm, st>>>(a, w, tw, 2)));
  }
  fix(a.dgamma[nl - 1]);
  ALG_RUN(KID_FIXUP, (k_fixup_e<TM><<<(gf.x + 127) / 128, 128, 0, st>>>(a.edge_c, a.rowptr, a.e0, a.e1, ntiles, a.esum, a.ecarry, a.plan, a.ci)));
  if (nl == 2) {
    ALG_RUN(KID_BK, (k_bk_tc<L, 'B', true><<<gp, b, sm, st>>>(a, w, tw, 0)));
    fix(a.dgamma[0]);
  } else if (nl == 3) {
    ALG_RUN(KID_BK, (k_bk_tc<L, 'D', false><<<gp, b, sm, st>>>(a, w, tw, 1)));
    fix(a.dgamma[1]);
    ALG_RUN(KID_BK, (k_bk_tc<L, 'C', true><<<gp, b, sm, st>>>(a, w, tw, 0)));
    fix(a.dgamma[0]);
  }
  ALG_RUN(KID_B0, (k_b0_tc<L><<<gp, b, sm, st>>>(a, w, tw)));
#undef ALG_RUN
  return cudaGetLastError();
}
#define ALG_DEFINE_PIPELINE_TC(L)                                                                  \
  const Pipeline* get_pipeline_tc_L##L() {                                                         \
    static const Pipeline p = {&info_tc_impl<L>, &init_tc_impl<L>, &run_chunk_tc_impl<L>,          \
                               &fused_grid_tc_impl<L>, &run_fused_tc_impl<L>};                     \
    return &p;                                                                                     \
  }

#define ALG_DEFINE_PIPELINE(L)                                                             \
  const Pipeline* get_pipeline_L##L() {                                                    \
    static const Pipeline p = {&info_impl<L>, &init_impl<L>, &run_chunk_impl<L>, nullptr, nullptr}; \
    return &p;                                                                             \
  }
#endif

const Pipeline* get_pipeline_L1();
const Pipeline* get_pipeline_L2();
const Pipeline* get_pipeline_L3();
const Pipeline* get_pipeline_tc_L1();
const Pipeline* get_pipeline_tc_L2();
const Pipeline* get_pipeline_tc_L3();

}  // namespace alg
