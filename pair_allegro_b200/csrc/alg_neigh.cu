// Binned FULL neighbour-list build on the device (the step on the caller's side of the
// path -- what LAMMPS' Neighbor class (NBinStandard + NPairFullBin [not in /root/reference]) does before
// PairNequIPAllegro::compute reads list->ilist / numneigh / firstneigh, pair_nequip_allegro.cpp:340-350, 469-480; the
// pair style asks for it with neighbor->add_request(this, REQ_FULL), :142-147).
//
// For every LOCAL atom i: all atoms j != i among locals + ghosts with |x_i - x_j|^2 <= rneigh^2 (f64), written as the
// KOKKOS-style 2-D view d_neighbors(i, jj) that alg_compute_device consumes, plus d_numneigh.  Deterministic: atoms are
// sorted by (cell, index) with a stable radix sort, a warp walks the 27 cells of its atom in a fixed order and compacts
// with ballots, so the same positions always give the same list in the same order (the edge order, and with it the
// summation order of the force evaluation, is reproducible).  A Verlet-skin check (largest displacement since the last
// build) decides whether the caller has to rebuild.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>

#include "../../include/allegro_b200.h"

namespace {

struct NBuf {
  void* p = nullptr; size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    const size_t want = std::max(bytes, cap + cap / 2);
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(want, 256));
    if (e == cudaSuccess) cap = std::max<size_t>(want, 256);
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Grid { double lo[3]; double inv; int n[3]; };

__device__ __forceinline__ int cell_of(const Grid& g, const double* x, int i, int* c3) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int c = (int)floor((x[3 * (size_t)i + k] - g.lo[k]) * g.inv);
    c3[k] = min(max(c, 0), g.n[k] - 1);
  }
  return (c3[0] * g.n[1] + c3[1]) * g.n[2] + c3[2];
}
__global__ void k_cell_keys(int ntot, const double* __restrict__ x, Grid g, int* __restrict__ key, int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntot) return;
  int c3[3];
  key[i] = cell_of(g, x, i, c3);
  idx[i] = i;
}
// first sorted position of every cell: cell_start[c] = lower bound of c in the sorted keys (ncell + 1 entries)
__global__ void k_cell_start(int ntot, int ncell, const int* __restrict__ skey, int* __restrict__ cell_start) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > ntot) return;
  const int cur = i < ntot ? skey[i] : ncell;
  const int prev = i > 0 ? skey[i - 1] : -1;
  for (int c = prev + 1; c <= cur; ++c) cell_start[c] = i;
}
// one warp per local atom; neighbours in (cell offset, sorted position) order
__global__ void k_full_list(int nlocal, const double* __restrict__ x, Grid g, const int* __restrict__ cell_start, const int* __restrict__ sidx,
                            double r2, int max_neigh, long long stride_i, long long stride_jj, int* __restrict__ neighbors,
                            int* __restrict__ numneigh, int* __restrict__ max_count) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= nlocal) return;
  const int i = warp;
  int c3[3];
  cell_of(g, x, i, c3);
  const double xi = x[3 * (size_t)i], yi = x[3 * (size_t)i + 1], zi = x[3 * (size_t)i + 2];
  int count = 0;
  for (int da = -1; da <= 1; ++da)
    for (int db = -1; db <= 1; ++db)
      for (int dc = -1; dc <= 1; ++dc) {
        const int a = c3[0] + da, b = c3[1] + db, c = c3[2] + dc;
        if (a < 0 || b < 0 || c < 0 || a >= g.n[0] || b >= g.n[1] || c >= g.n[2]) continue;
        const int cell = (a * g.n[1] + b) * g.n[2] + c;
        const int s0 = cell_start[cell], s1 = cell_start[cell + 1];
        for (int s = s0; s < s1; s += 32) {
          const int p = s + lane;
          bool keep = false;
          int j = 0;
          if (p < s1) {
            j = sidx[p];
            const double dx = xi - x[3 * (size_t)j], dy = yi - x[3 * (size_t)j + 1], dz = zi - x[3 * (size_t)j + 2];
            keep = j != i && __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)) <= r2;
          }
          const unsigned m = __ballot_sync(0xffffffffu, keep);
          if (keep) {
            const int k = count + __popc(m & ((1u << lane) - 1u));
            if (k < max_neigh) neighbors[(long long)i * stride_i + (long long)k * stride_jj] = j;
          }
          count += __popc(m);
        }
      }
  if (lane == 0) {
    numneigh[i] = min(count, max_neigh);
    atomicMax(max_count, count);
  }
}
// largest squared displacement since the positions of the last build (Verlet-skin criterion, LAMMPS Neighbor::check_distance)
__global__ void k_max_disp2(int n, const double* __restrict__ x, const double* __restrict__ xold, unsigned long long* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double d2 = 0.0;
  if (i < n) {
    const double dx = x[3 * (size_t)i] - xold[3 * (size_t)i], dy = x[3 * (size_t)i + 1] - xold[3 * (size_t)i + 1], dz = x[3 * (size_t)i + 2] - xold[3 * (size_t)i + 2];
    d2 = dx * dx + dy * dy + dz * dz;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d2 = fmax(d2, __shfl_xor_sync(0xffffffffu, d2, o));
  if ((threadIdx.x & 31) == 0 && d2 > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(d2));   // non-negative doubles order like integers
}

}  // namespace

struct alg_neigh {
  int device = 0;
  std::string err;
  NBuf key, idx, skey, sidx, cell_start, tmp, scalars, xold;
  int nold = 0;
  int* h_scalars = nullptr;       // pinned: [0] max neighbour count of the last build, [2..3] max squared displacement (bits)
  int last_max = 0;
};

#define NCKH(call)                                                                                          \
  do {                                                                                                      \
    cudaError_t _e = (call);                                                                                \
    if (_e != cudaSuccess) { n->err = std::string("CUDA error: ") + cudaGetErrorString(_e) + " (" #call ")"; return ALG_ECUDA; } \
  } while (0)

extern "C" int alg_neigh_create(int cuda_device, alg_neigh** out) {
  if (!out) return ALG_EINVAL;
  *out = nullptr;
  if (cudaSetDevice(cuda_device) != cudaSuccess) { cudaGetLastError(); return ALG_ECUDA; }
  alg_neigh* n = new alg_neigh();
  n->device = cuda_device;
  if (cudaMallocHost(&n->h_scalars, 64) != cudaSuccess) { delete n; return ALG_ECUDA; }
  *out = n;
  return ALG_OK;
}
extern "C" void alg_neigh_destroy(alg_neigh* n) {
  if (!n) return;
  cudaSetDevice(n->device);
  for (NBuf* b : {&n->key, &n->idx, &n->skey, &n->sidx, &n->cell_start, &n->tmp, &n->scalars, &n->xold}) b->release();
  if (n->h_scalars) cudaFreeHost(n->h_scalars);
  delete n;
}
extern "C" const char* alg_neigh_last_error(const alg_neigh* n) { return n ? n->err.c_str() : ""; }

extern "C" int alg_neigh_build(alg_neigh* n, int nlocal, int nghost, const double* d_x, const double* lo, const double* hi, double rneigh,
                               int max_neigh, int64_t stride_i, int64_t stride_jj, int* d_neighbors, int* d_numneigh, int* max_count, void* stream) {
  if (!n) return ALG_EINVAL;
  if (nlocal < 0 || nghost < 0 || !lo || !hi || rneigh <= 0.0 || max_neigh < 1 || (nlocal > 0 && (!d_x || !d_neighbors || !d_numneigh))) {
    n->err = "alg_neigh_build: bad arguments"; return ALG_EINVAL;
  }
  if (max_count) *max_count = 0;
  if (nlocal == 0) return ALG_OK;
  NCKH(cudaSetDevice(n->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int ntot = nlocal + nghost;
  Grid g;
  long ncell = 1;
  g.inv = 1.0 / rneigh;
  for (int k = 0; k < 3; ++k) {
    g.lo[k] = lo[k];
    g.n[k] = std::max(1, (int)std::floor((hi[k] - lo[k]) / rneigh) + 1);
    ncell *= g.n[k];
  }
  if (ncell > 0x3fffffff) { n->err = "alg_neigh_build: too many cells (box / rneigh)"; return ALG_EINVAL; }
  NCKH(n->key.ensure(sizeof(int) * ntot)); NCKH(n->idx.ensure(sizeof(int) * ntot));
  NCKH(n->skey.ensure(sizeof(int) * ntot)); NCKH(n->sidx.ensure(sizeof(int) * ntot));
  NCKH(n->cell_start.ensure(sizeof(int) * (ncell + 2))); NCKH(n->scalars.ensure(64));
  k_cell_keys<<<(ntot + 255) / 256, 256, 0, st>>>(ntot, d_x, g, n->key.as<int>(), n->idx.as<int>());
  int bits = 1;
  while ((1L << bits) < ncell) ++bits;
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, n->key.as<int>(), n->skey.as<int>(), n->idx.as<int>(), n->sidx.as<int>(), ntot, 0, bits, st);
  NCKH(n->tmp.ensure(tb));
  NCKH(cub::DeviceRadixSort::SortPairs(n->tmp.p, tb, n->key.as<int>(), n->skey.as<int>(), n->idx.as<int>(), n->sidx.as<int>(), ntot, 0, bits, st));
  k_cell_start<<<(ntot + 256) / 256, 256, 0, st>>>(ntot, (int)ncell, n->skey.as<int>(), n->cell_start.as<int>());
  NCKH(cudaMemsetAsync(n->scalars.p, 0, 64, st));
  const long nthreads = (long)nlocal * 32;
  k_full_list<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(nlocal, d_x, g, n->cell_start.as<int>(), n->sidx.as<int>(), rneigh * rneigh, max_neigh,
                                                                 (long long)stride_i, (long long)stride_jj, d_neighbors, d_numneigh, n->scalars.as<int>());
  NCKH(cudaGetLastError());
  // remember the positions of this build for the Verlet-skin check
  NCKH(n->xold.ensure(sizeof(double) * 3 * ntot));
  NCKH(cudaMemcpyAsync(n->xold.p, d_x, sizeof(double) * 3 * ntot, cudaMemcpyDeviceToDevice, st));
  n->nold = ntot;
  if (max_count) {       // the caller wants to know whether max_neigh was enough: one small synchronisation (builds are rare)
    NCKH(cudaMemcpyAsync(n->h_scalars, n->scalars.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    NCKH(cudaStreamSynchronize(st));
    *max_count = n->h_scalars[0];
    n->last_max = *max_count;
    if (*max_count > max_neigh) { n->err = "alg_neigh_build: an atom has more neighbours than max_neigh (rows truncated): rebuild with a larger view"; return ALG_ESTATE; }
  }
  return ALG_OK;
}

// 1 if some atom moved further than skin/2 since the last build (or the atom count changed): the list must be rebuilt
extern "C" int alg_neigh_check(alg_neigh* n, int ntot, const double* d_x, double skin, int* rebuild, void* stream) {
  if (!n || !rebuild || (ntot > 0 && !d_x)) return ALG_EINVAL;
  *rebuild = 1;
  if (ntot != n->nold || ntot == 0) return ALG_OK;
  NCKH(cudaSetDevice(n->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned long long* d_out = reinterpret_cast<unsigned long long*>(n->scalars.as<char>() + 16);
  NCKH(cudaMemsetAsync(d_out, 0, sizeof(unsigned long long), st));
  k_max_disp2<<<(ntot + 255) / 256, 256, 0, st>>>(ntot, d_x, n->xold.as<double>(), d_out);
  NCKH(cudaMemcpyAsync(n->h_scalars + 2, d_out, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  NCKH(cudaStreamSynchronize(st));
  double d2;
  memcpy(&d2, n->h_scalars + 2, sizeof(double));
  *rebuild = d2 > 0.25 * skin * skin ? 1 : 0;
  return ALG_OK;
}
