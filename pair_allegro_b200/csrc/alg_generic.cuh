// Width-generic pipeline: the same network for ANY num_scalar_features / num_tensor_features /
// mlp width and depth / readout width, as a plain sequence of per-operation kernels over edge-major [edge][feature] arrays
// of one centre-aligned chunk.  It exists so that models outside the widths the tiled kernels are specialised for
// (S = 64, U = 32, MLP 2x64, readout 32) load and run instead of failing with ALG_EINVAL (e.g. the "high-capacity"
// S = 128 / U = 64 / H = 128 model of BASELINE.json configs[4]); it is also a third, independent implementation of the
// chain rule for the standard widths (option gemm=generic).  FP32 FFMA GEMMs, activations kept for the backward,
// per-centre environment sums as ordered loops (deterministic), forces / virial through the same fixed-point accumulators
// as the tiled kernels.  Not tuned: every intermediate round-trips HBM.
// Replaces, like the other pipelines, /root/reference/pair_nequip_allegro.cpp:425 (model.forward incl. autograd).
#pragma once
#include <cuda_runtime.h>

#include "alg_common.cuh"

namespace alg {

constexpr int GEN_MAXLIN = 5;        // linear layers of one MLP (mlp_depth + 1)

struct GenMLP {
  const float* w[GEN_MAXLIN];        // w[i]  : [dims[i]][dims[i+1]] row-major
  const float* wt[GEN_MAXLIN];       // wt[i] : [dims[i+1]][dims[i]]
  int dims[GEN_MAXLIN + 1];
  int nlin;
};
struct GenLayer {
  const float* env; const float* env_t;      // [S][ENVW], [ENVW][S]
  const float* omega_t;                      // [U][npaths]  (channel-major: the generated tensor products read omega[path])
  GenMLP mlp;                                // S + n0*U -> H.. -> S
  float a, b;
  char kind;                                 // 'A'..'D' (tp_gen.cuh)
};
struct GenModel {
  int S, H, U, R, L, nl, T, B, depth;
  GenMLP two;                                // 2T+B -> H.. -> S
  const float* emb; const float* emb_t;      // [S][ENVW], [ENVW][S]
  GenLayer layer[3];
  const float* ro0; const float* ro0_t; const float* ro1;
  float p, inv_sqrt_n;
};
// small tables passed to kernels by value
struct GenTables { float rc[MAXT * MAXT]; float gscale[MAXT]; };

struct GenEdges {
  const float4* rvec; const int* edge_j; const int* edge_c; const int* rowptr; const int* ilist;
  double* esum; float* edge_energy; float* edge_grad;
  unsigned long long* facc; unsigned long long* vacc;
};

// floats of workspace one chunk of n edges / nc centres needs
size_t gen_work_floats(const GenModel& m, long n, long nc);
// one centre-aligned chunk: centres [c0, c1), edges [e0, e1) = [rowptr[c0], rowptr[c1])
cudaError_t gen_run_chunk(const GenModel& m, const GenTables& tb, const GenEdges& g, int c0, int c1, int e0, int e1, float* work,
                          cudaStream_t st, long* launches);

// tensor-product kernels, one translation unit per l_max (alg_generic_tp_L{1,2,3}.cu)
struct GenTp {
  int n, U, S, e0, c0, ldin, envw;
  const int* edge_c;
  const float* vin;        // first layer: w0 [n][ENVW]; else V^k [DIN][n][U] (component-major)
  const float* Y;          // [n][NSH]
  const float* gamma;      // [nc][NSH*U]
  const float* omega_t;
  float* vout;             // forward: V^{k+1} [DOUT][n][U] or nullptr
  float* IN;               // forward: s written to columns S.. of [n][ldin]; backward: ds read from there
  const float* dvout;      // backward: dV^{k+1} or nullptr
  float* dvin;             // backward: [DIN][n][U]
  float* dge;              // backward: per-edge dGamma [NSH][n][U]
};
struct GenTpDims { int din, dout, npath, n0; };
GenTpDims gen_tp_dims(int L, char kind);
cudaError_t gen_tp_launch(int L, char kind, bool first, bool backward, const GenTp& a, cudaStream_t st);
cudaError_t gen_tp_launch_L1(char kind, bool first, bool backward, const GenTp& a, cudaStream_t st);
cudaError_t gen_tp_launch_L2(char kind, bool first, bool backward, const GenTp& a, cudaStream_t st);
cudaError_t gen_tp_launch_L3(char kind, bool first, bool backward, const GenTp& a, cudaStream_t st);
GenTpDims gen_tp_dims_L1(char kind);
GenTpDims gen_tp_dims_L2(char kind);
GenTpDims gen_tp_dims_L3(char kind);

}  // namespace alg
