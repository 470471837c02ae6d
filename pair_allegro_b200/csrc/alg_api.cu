// C-ABI of the Allegro B200 engine (include/allegro_b200.h): weight loader, device scratch
// management, edge-list build from the LAMMPS full neighbour list, chunked pipeline
// orchestration, output store.  Replaces /root/reference/pair_nequip_allegro.cpp:333-650
// (compute/preprocess/call) and pair_nequip_allegro_kokkos.cpp:87-353.  No libtorch, no CPU
// fallback: every failure surfaces as an error code.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif
#include <cub/device/device_scan.cuh>

#include "../../include/allegro_b200.h"
#include "alg_pipeline.cuh"
#include "alg_generic.cuh"

namespace alg {
const Pipeline* get_pipeline(int L) {
  switch (L) {
    case 1: return get_pipeline_L1();
    case 2: return get_pipeline_L2();
    case 3: return get_pipeline_L3();
  }
  return nullptr;
}
const Pipeline* get_pipeline_tc(int L) { return L == 1 ? get_pipeline_tc_L1() : (L == 2 ? get_pipeline_tc_L2() : (L == 3 ? get_pipeline_tc_L3() : nullptr)); }
}  // namespace alg

using namespace alg;

#define NEIGHMASK 0x1FFFFFFF

static std::string g_create_error;

// ------------------------------------------------------------------------------------------
// device buffer that only grows (geometric), never shrinks
// ------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    size_t want = std::max(bytes, cap + cap / 2);
    want = (want + 255) / 256 * 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { cap = 0; p = nullptr; return e; }
    cap = want;
    return cudaSuccess;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    size_t want = std::max(bytes, cap + cap / 2);
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) { cap = 0; p = nullptr; return e; }
    cap = want;
    return cudaSuccess;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct HostTensor { std::string dtype; std::vector<long> dims; std::vector<char> data; size_t count() const { size_t n = 1; for (long d : dims) n *= d; return n; } };

struct alg_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  // model
  std::map<std::string, std::string> header;
  std::map<std::string, HostTensor> tensors;
  int L = 0, nl = 0, T = 0, B = 0;
  double r_max = 0, avg_n = 1, p = 6;
  std::string type_names;
  std::vector<double> per_edge_cut;   // T*T or empty
  std::vector<double> cut_table;      // T*T (model types)
  std::vector<double> scales, shifts;
  int allow_tf32 = 0;
  DevBuf weights, tc_weights;
  ModelW mw{};
  TcW tcw{};
  const Pipeline* pipe = nullptr;        // active pipeline
  PipelineInfo pinfo{};
  const Pipeline* pipe_ffma = nullptr;
  const Pipeline* pipe_tc = nullptr;
  bool use_tc = false;
  // width-generic pipeline (alg_generic.cuh): the only one for models outside the specialised widths; option gemm=generic otherwise
  bool generic = false, std_widths = true;
  GenModel gm{};
  GenTables gtb{};
  DevBuf gen_weights, gen_work;
  long gen_chunk_edges = 65536;
  // fused persistent pipeline (centre-aligned tiles, allegro_kernels_tc.cuh: k_fused_tc)
  int pipeline_mode = 0;                   // option pipeline: 0 = auto, 1 = fused, 2 = tiled
  bool force_tiled = false;                // sticky: a centre with more than 128 edges was seen -> chunked edge-tile pipeline
  int fused_grid = 0;                      // resident CTAs of the fused kernel on this device
  int fused_batch = 8;                     // option fused_batch: tiles a CTA runs phase by phase before it takes the next batch
  long edge_cap_hint = 0;                  // upper bound on the edge count known to the caller path (0 = unknown)
  long max_neighbors = 0;                  // option max_neighbors: extent(1) of the caller's 2-D neighbour view
  long last_E_known = -1; int last_E_nlocal = -1;
  DevBuf d_blk, d_blk_base, d_tile_c0, d_info, d_sm_phase, d_cplan;
  bool tiled_plan_device = true;           // option chunk_plan: device (default) | host
  bool tiled_sync = false;                 // sticky: the device-built chunk plan did not fit the buffers -> host-built plan (one sync per step)
  int num_sms = 148;
  bool phase_align = true;                 // option phase_align
  // deferred verification of asynchronous steps: a ring of pinned info records, one event each, so that the host may run
  // up to INFO_RING steps ahead of the device (it only blocks when a slot is about to be reused)
  static constexpr int INFO_RING = 4;
  PinBuf h_info;                           // INFO_RING x 16 ints; slot 0 doubles as the record of synchronous steps
  cudaEvent_t ev_info[INFO_RING] = {nullptr, nullptr, nullptr, nullptr};
  bool info_used[INFO_RING] = {false, false, false, false};
  int info_mode[INFO_RING] = {0, 0, 0, 0};  // 0 = host-planned chunked step, 1 = fused, 2 = device-planned chunked step
  unsigned info_head = 0, info_tail = 0;   // pending slots: [tail, head)
  bool last_fused = false;
  int last_mode = 0;                       // 0 host-planned chunked, 1 fused, 2 device-planned chunked
  std::string deferred_err;                // failure of an asynchronous step, reported by the next call
  std::pair<const void*, size_t> reg[3] = {{nullptr, 0}, {nullptr, 0}, {nullptr, 0}};   // caller arrays (x, f, type) pinned with cudaHostRegister
  bool host_register = true;               // option host_register
  DevBuf d_f_stage;
  PinBuf h_eatom;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copy = nullptr, ev_order = nullptr;
  int neigh_ago = 0;                       // option neigh_ago: steps since the caller's last neighbour-list rebuild
  int list_nlocal = -1, list_ntot = -1, list_reused = 0;
  long long list_tot = -1;                 // device-resident copy of the host neighbour list (alg_compute_host)
  // type map
  int ntypes = 0;
  DevBuf d_tmap, d_cutsq, d_scale, d_shift;
  bool have_map = false;
  // options
  bool filter_le = true, keep_edges = false, debug = false;
  long chunk_edges = 1 << 21;
  // per-step scratch
  DevBuf d_x, d_type, d_ilist, d_numneigh, d_cand, d_first, d_cnt, d_rowptr, d_scan_tmp;
  DevBuf d_mtype, d_edge_j, d_edge_c, d_rvec, d_esum, d_facc, d_vacc, d_forces, d_eall, d_red, d_edge_index, d_edge_energy, d_edge_grad, d_eatom_out;
  DevBuf c_ZD[4], c_X[3], c_W0, c_V[3], c_dX, c_dV[2], c_dY, c_du, c_gamma[3], c_dgamma[3], c_carry, c_ecarry;
  PinBuf h_stage, h_rowptr, h_out, h_first;
  // results
  int last_nlocal = 0, last_ntot = 0;
  long last_E = 0;
  std::vector<int64_t> edges_host;
  std::map<std::string, std::vector<double>> outputs;
  double timings[3] = {0, 0, 0};
  Prof prof;
  double kernel_ms[KID_COUNT] = {0, 0, 0, 0, 0, 0, 0};
  double kernel_n[KID_COUNT] = {0, 0, 0, 0, 0, 0, 0};
  double step_stats[4] = {0, 0, 0, 0};   // launches of own kernels, edges, chunks, tiles
  double host_ms[3] = {0, 0, 0};          // alg_compute_host wall times of the last call: list flatten + upload issue, whole call, pinning
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  // debug bookkeeping of the last single-chunk run
  int dbg_ntiles = 0, dbg_c0 = 0, dbg_ncent = 0;
};

#define CK(call)                                                                                      \
  do {                                                                                                \
    cudaError_t _e = (call);                                                                          \
    if (_e != cudaSuccess) {                                                                          \
      h->err = std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + std::to_string(__LINE__) + " (" #call ")"; \
      return ALG_ECUDA;                                                                               \
    }                                                                                                 \
  } while (0)

static int fail(alg_handle* h, int code, const std::string& msg) { h->err = msg; return code; }

// ------------------------------------------------------------------------------------------
// .alg reader (format: pair_allegro_b200/export.py)
// ------------------------------------------------------------------------------------------
static bool read_alg(const char* path, alg_handle* h, std::string& err) {
  FILE* f = fopen(path, "rb");
  if (!f) { err = std::string("cannot open weight file ") + path; return false; }
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<char> raw(sz);
  if (sz <= 0 || fread(raw.data(), 1, sz, f) != (size_t)sz) { fclose(f); err = "short read on weight file"; return false; }
  fclose(f);
  std::string all(raw.data(), std::min<long>(sz, 1 << 20));
  size_t endp = all.find("\nend\n");
  if (all.compare(0, 9, "ALGB200 1") != 0 || endp == std::string::npos) { err = std::string(path) + " is not an ALGB200 version-1 weight file"; return false; }
  std::istringstream in(all.substr(0, endp + 1));
  std::string line;
  long data_offset = -1;
  std::getline(in, line);
  while (std::getline(in, line)) {
    if (line.empty()) continue;
    size_t sp = line.find(' ');
    std::string key = line.substr(0, sp), val = sp == std::string::npos ? "" : line.substr(sp + 1);
    if (key == "data_offset") { data_offset = atol(val.c_str()); continue; }
    if (key == "tensor") {
      std::istringstream ts(val);
      std::string name, dt; int nd;
      ts >> name >> dt >> nd;
      HostTensor t; t.dtype = dt;
      for (int i = 0; i < nd; ++i) { long d; ts >> d; t.dims.push_back(d); }
      long off = -1, nb = -1; ts >> off >> nb;
      long cnt = 1;
      bool dims_ok = nd >= 0 && nd <= 8 && (long)t.dims.size() == nd;
      for (long d : t.dims) { if (d < 0 || (d > 0 && cnt > (1L << 40) / d)) dims_ok = false; else cnt *= d; }
      const long esz = dt == "f32" ? 4 : (dt == "f64" ? 8 : (dt == "i32" ? 4 : (dt == "i64" ? 8 : 0)));
      if (ts.fail() || !dims_ok || esz == 0) { err = "malformed tensor record '" + name + "' in weight file header"; return false; }
      if (data_offset < 0 || off < 0 || nb < 0 || nb != cnt * esz || data_offset > sz || off > sz - data_offset || nb > sz - data_offset - off) {
        err = "tensor " + name + " out of file bounds"; return false;
      }
      t.data.assign(raw.begin() + data_offset + off, raw.begin() + data_offset + off + nb);
      h->tensors[name] = std::move(t);
    } else {
      h->header[key] = val;
    }
  }
  return true;
}

static const float* tf32(alg_handle* h, const std::string& name, long d0, long d1, std::string& err) {
  auto it = h->tensors.find(name);
  if (it == h->tensors.end()) { err = "weight file misses tensor " + name; return nullptr; }
  const HostTensor& t = it->second;
  long n = 1; for (long d : t.dims) n *= d;
  if (t.dtype != "f32" || n != d0 * d1 || t.data.size() != (size_t)n * sizeof(float)) {
    err = "tensor " + name + " has unexpected shape/dtype (expected " + std::to_string(d0) + "x" + std::to_string(d1) + " f32)";
    return nullptr;
  }
  return reinterpret_cast<const float*>(t.data.data());
}

// ------------------------------------------------------------------------------------------
// edge build kernels (K1): LAMMPS full list -> centre-sorted CSR, same filter and same
// (ilist, jlist) order as the reference host loop (pair_nequip_allegro.cpp:488-512, 566-629)
// ------------------------------------------------------------------------------------------
struct NeighAcc {
  const int* base;
  const long long* first;   // flat mode: offset of centre slot ii
  const int* cnt;           // flat: per slot ii ; 2-D: per atom i (numneigh)
  long long stride_i, stride_jj;
  int flat;
};

__device__ __forceinline__ double rsq_nofma(const double* __restrict__ x, int i, int j, double& dx, double& dy, double& dz) {
  dx = x[3 * (size_t)i + 0] - x[3 * (size_t)j + 0];
  dy = x[3 * (size_t)i + 1] - x[3 * (size_t)j + 1];
  dz = x[3 * (size_t)i + 2] - x[3 * (size_t)j + 2];
  // same rounding as the reference's `dx*dx + dy*dy + dz*dz` compiled without contraction
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

template <bool FILL>
__global__ void k_edges(int nlocal, const double* __restrict__ x, const int* __restrict__ type, const int* __restrict__ ilist,
                        NeighAcc acc, const double* __restrict__ cutsq, int ntypes, int filter_le,
                        int* __restrict__ cnt_out, const int* __restrict__ rowptr, const int* __restrict__ mtype,
                        int* __restrict__ edge_j, int* __restrict__ edge_c, float4* __restrict__ rvec, long cap) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= nlocal) return;
  const int ii = warp;
  const int i = ilist[ii];
  const int ti = type[i] - 1;
  int n; const int* ptr; long long st;
  if (acc.flat) { n = acc.cnt[ii]; ptr = acc.base + acc.first[ii]; st = 1; }
  else { n = acc.cnt[i]; ptr = acc.base + (long long)i * acc.stride_i; st = acc.stride_jj; }
  int total = 0;
  int base = FILL ? rowptr[ii] : 0;
  const int zi = mtype[i];
  if (zi < 0) n = 0;                                  // unmapped centre: no edges (the step reports the error, see k_mtype)
  for (int j0 = 0; j0 < n; j0 += 32) {
    const int jj = j0 + lane;
    bool keep = false;
    int j = 0; double dx = 0, dy = 0, dz = 0;
    if (jj < n) {
      j = ptr[(long long)jj * st] & NEIGHMASK;
      const double rsq = rsq_nofma(x, i, j, dx, dy, dz);
      const double c2 = cutsq[ti * ntypes + (type[j] - 1)];
      keep = (filter_le ? (rsq <= c2) : (rsq < c2)) && mtype[j] >= 0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (FILL) {
      if (keep) {
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        if (pos >= cap) continue;                       // edge arrays too small (flagged by k_plan; the host retries)
        edge_j[pos] = j;
        edge_c[pos] = ii;
        const int zj = mtype[j];
        rvec[pos] = make_float4((float)(-dx), (float)(-dy), (float)(-dz), __int_as_float(zi | (zj << 8)));
      }
      base += __popc(m);
    } else {
      total += __popc(m);
    }
  }
  if (!FILL && lane == 0) cnt_out[ii] = total;
}

__global__ void k_edge_index(const int* __restrict__ rowptr, int nlocal, long cap, const int* __restrict__ edge_j, const int* __restrict__ edge_c,
                             const int* __restrict__ ilist, long long* __restrict__ out) {
  const long E = rowptr[nlocal];
  if (E > cap) return;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < E; e += (long)gridDim.x * blockDim.x) {
    out[e] = ilist[edge_c[e]];
    out[E + e] = edge_j[e];
  }
}

// ------------------------------------------------------------------------------------------
// batch plan of the fused kernel: centre-aligned batches of <= `rows` edges and <= `rows` centres (rows = B*128).
// Centres are split into blocks of PLAN_CB; one thread packs its block greedily (batches never span blocks), pass 0
// counts the batches of a block, pass 1 (after an exclusive scan) writes the first centre of every batch.
// info: [0] nbatch, [1] max degree, [2] E, [3] 1 if E exceeds the capacity of the edge arrays, [6] number of 128-edge tiles
// ------------------------------------------------------------------------------------------
constexpr int PLAN_CB = 1024;      // centres per planning block (one thread packs them greedily from shared memory)
constexpr int PLAN_TPB = 8;        // planning blocks per CUDA block (8 x 1024 row pointers staged in 32 KB of shared memory)
template <bool FILL>
__global__ void __launch_bounds__(256) k_plan(int nlocal, const int* __restrict__ rowptr, int* __restrict__ blk_tiles, const int* __restrict__ blk_base,
                                              int* __restrict__ tile_c0, int* __restrict__ info, long cap, int PLAN_ROWS) {
  __shared__ int rp[PLAN_TPB * PLAN_CB + 1];
  const int nblk = (nlocal + PLAN_CB - 1) / PLAN_CB;
  const int c_base = blockIdx.x * PLAN_TPB * PLAN_CB;
  const int c_end = min(nlocal, c_base + PLAN_TPB * PLAN_CB);
  for (int i = threadIdx.x; i <= c_end - c_base; i += blockDim.x) rp[i] = rowptr[c_base + i];
  __syncthreads();
  const int b = blockIdx.x * PLAN_TPB + threadIdx.x;
  if (threadIdx.x >= PLAN_TPB || b >= nblk) return;
  const int cb = b * PLAN_CB, ce = min(nlocal, cb + PLAN_CB);
  // the last ~4 % of the centres are cut into quarter-size batches: the batch queue then drains evenly (a full batch is
  // ~1/24 of a CTA's share at 1 M atoms; without this the last round leaves most SMs idle)
  if (PLAN_ROWS >= 512 && (long)cb * 25 >= (long)nlocal * 24) PLAN_ROWS >>= 2;
  int rows = 0, span = 0, nt = 0, maxdeg = 0, tiles = 0;
  int base = FILL ? blk_base[b] : 0;
  int prev = rp[cb - c_base];
  for (int c = cb; c < ce; ++c) {
    const int nxt = rp[c + 1 - c_base];
    const int deg = nxt - prev;
    prev = nxt;
    maxdeg = max(maxdeg, deg);
    if (span > 0 && (rows + deg > PLAN_ROWS || span == PLAN_ROWS)) { ++nt; tiles += (rows + 127) >> 7; rows = 0; span = 0; }
    if (FILL && span == 0) tile_c0[base + nt] = c;
    rows += deg; ++span;
  }
  if (span > 0) { ++nt; tiles += (rows + 127) >> 7; }
  if (FILL) atomicAdd(info + 6, tiles);
  if (!FILL) {
    blk_tiles[b] = nt;
    if (b == 0) blk_tiles[nblk] = 0;
    atomicMax(info + 1, maxdeg);
  } else if (b == nblk - 1) {
    const int ntiles = base + nt;
    tile_c0[ntiles] = nlocal;
    info[0] = ntiles;
    const int E = rowptr[nlocal];
    info[2] = E;
    info[3] = E > cap ? 1 : 0;
  }
}

// model type per atom; an atom whose LAMMPS type has no model type raises flags[0] (the step then fails with ALG_EINVAL)
__global__ void k_mtype(int ntot, const int* __restrict__ type, const int* __restrict__ tmap, int ntypes, int* __restrict__ mtype, int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntot) return;
  const int t = type[i] - 1;
  const int z = (t >= 0 && t < ntypes) ? tmap[t] : -1;
  mtype[i] = z;
  if (z < 0) flags[0] = 1 + i;
}

// finalize: fixed-point accumulators -> model forces (double), optionally f += ; per-atom energies
// Chunk plan of the chunked pipeline, built on the device (the round-1 version copied the row pointer to the host and
// blocked on it every step -- the same mid-step synchronisation as the reference's edge count, kokkos.cpp:203-206).
// Chunk i = the centres whose row STARTS in [i*Q, (i+1)*Q): boundaries by binary search, one thread per chunk.
// plan[3i..3i+2] = {e0, e1, c0}; chunks behind the last edge are empty (e0 == e1).  info[2] = E, info[3] |= capacity
// overflow, info[9] = 1 when a chunk exceeds the buffers (max_edges edges or max_cent centres).
__global__ void k_chunk_plan(int nlocal, const int* __restrict__ rowptr, long Q, int nch_max, int max_edges, int max_cent, long cap,
                             int* __restrict__ plan, int* __restrict__ info) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nch_max) return;
  const long E = rowptr[nlocal];
  auto lb = [&](long v) {                           // first centre c in [0, nlocal] with rowptr[c] >= v
    int lo = 0, hi = nlocal;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (rowptr[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
  };
  const int ca = (long)i * Q >= E ? nlocal : lb((long)i * Q);
  const int cb = (long)(i + 1) * Q >= E ? nlocal : lb((long)(i + 1) * Q);
  int e0 = rowptr[ca], e1 = rowptr[cb];
  if (e1 - e0 > max_edges || (e1 > e0 && cb - ca > max_cent)) { info[9] = 1; e1 = e0; }   // does not fit the buffers: no work, the host repeats the step
  if (E > cap) e1 = e0;                                                                  // edge arrays too small (info[3]): no work either
  plan[3 * i] = e0; plan[3 * i + 1] = e1; plan[3 * i + 2] = ca;
  if (i == 0) { info[2] = (int)E; info[0] = (int)((E + Q - 1) / Q); if (E > cap) info[3] = 1; }
}

// Thread-per-atom variant of K1 for the KOKKOS device layout (LayoutLeft: d_neighbors(i,jj) at base[i + jj*stride_jj]): the
// lanes of a warp are consecutive atoms, so every step of the jj loop is one coalesced read (the warp-per-atom kernel
// above would touch 32 different lines per step).  Same filter, same (ilist, jlist) order.
template <bool FILL>
__global__ void k_edges_tpa(int nlocal, const double* __restrict__ x, const int* __restrict__ type, const int* __restrict__ ilist,
                            NeighAcc acc, const double* __restrict__ cutsq, int ntypes, int filter_le,
                            int* __restrict__ cnt_out, const int* __restrict__ rowptr, const int* __restrict__ mtype,
                            int* __restrict__ edge_j, int* __restrict__ edge_c, float4* __restrict__ rvec, long cap) {
  const int ii = blockIdx.x * blockDim.x + threadIdx.x;
  if (ii >= nlocal) return;
  const int i = ilist[ii];
  const int ti = type[i] - 1;
  const int zi = mtype[i];
  const int n = zi < 0 ? 0 : acc.cnt[i];
  const int* ptr = acc.base + (long long)i * acc.stride_i;
  int pos = FILL ? rowptr[ii] : 0, total = 0;
  for (int jj = 0; jj < n; ++jj) {
    const int j = ptr[(long long)jj * acc.stride_jj] & NEIGHMASK;
    double dx, dy, dz;
    const double rsq = rsq_nofma(x, i, j, dx, dy, dz);
    const double c2 = cutsq[ti * ntypes + (type[j] - 1)];
    const bool keep = (filter_le ? (rsq <= c2) : (rsq < c2)) && mtype[j] >= 0;
    if (!keep) continue;
    if (FILL) {
      if (pos < cap) {
        edge_j[pos] = j;
        edge_c[pos] = ii;
        rvec[pos] = make_float4((float)(-dx), (float)(-dy), (float)(-dz), __int_as_float(zi | (mtype[j] << 8)));
      }
      ++pos;
    } else {
      ++total;
    }
  }
  if (!FILL) cnt_out[ii] = total;
}

// info != nullptr (device-built plans): a step the pipeline refused (info[1] > rows, info[3] or info[9]) must not touch f; the host repeats it
__global__ void k_forces(int ntot, const unsigned long long* __restrict__ facc, double* __restrict__ forces, double* __restrict__ f_inout,
                         const int* __restrict__ info, int rows) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= 3L * ntot) return;
  if (info && (info[1] > rows || info[3] != 0 || info[9] != 0)) return;
  const double v = (double)(long long)facc[i] * FIX_INV;
  forces[i] = v;
  if (f_inout) f_inout[i] += v;
}
__global__ void k_eall_ghost(int ntot, const int* __restrict__ mtype, const double* __restrict__ shift, double* __restrict__ eall) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ntot) { const int z = mtype[i]; eall[i] = z >= 0 ? shift[z] : 0.0; }
}
// one block-partial per 1024 centres, fixed reduction tree -> deterministic
__global__ void k_eall_local(int nlocal, const int* __restrict__ ilist, const int* __restrict__ mtype, const double* __restrict__ esum,
                             const double* __restrict__ scale, const double* __restrict__ shift, double inv_sqrt_n,
                             double* __restrict__ eall, double* __restrict__ eatom, double* __restrict__ partial) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int q = 0; q < 4; ++q) {
    const int ii = blockIdx.x * 1024 + q * 256 + threadIdx.x;
    if (ii < nlocal) {
      const int i = ilist[ii];
      const int z = mtype[i];
      const double e = z >= 0 ? scale[z] * (inv_sqrt_n * esum[ii]) + shift[z] : 0.0;
      eall[i] = e;
      if (eatom) eatom[i] = e;
      acc += e;
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
__global__ void k_final_scalars(int nblocks, const double* __restrict__ partial, const unsigned long long* __restrict__ vacc, double* __restrict__ out7) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double e = 0.0;
    for (int b = 0; b < nblocks; ++b) e += partial[b];
    out7[0] = e;
    for (int q = 0; q < 6; ++q) out7[1 + q] = (double)(long long)vacc[q] / VIR_SCALE;
  }
}

__global__ void k_halo_pack(const double* __restrict__ x, const int* __restrict__ list, int n, const double* __restrict__ shift, double* __restrict__ buf) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int i = list[k];
  buf[3 * (size_t)k + 0] = x[3 * (size_t)i + 0] + (shift ? shift[3 * (size_t)k + 0] : 0.0);
  buf[3 * (size_t)k + 1] = x[3 * (size_t)i + 1] + (shift ? shift[3 * (size_t)k + 1] : 0.0);
  buf[3 * (size_t)k + 2] = x[3 * (size_t)i + 2] + (shift ? shift[3 * (size_t)k + 2] : 0.0);
}
__global__ void k_halo_unpack_add(double* __restrict__ f, const int* __restrict__ list, int n, const double* __restrict__ buf) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int i = list[k];
  atomicAdd(f + 3 * (size_t)i + 0, buf[3 * (size_t)k + 0]);   // list entries may repeat (several images of one atom)
  atomicAdd(f + 3 * (size_t)i + 1, buf[3 * (size_t)k + 1]);
  atomicAdd(f + 3 * (size_t)i + 2, buf[3 * (size_t)k + 2]);
}

// width-generic pipeline: plain row-major fp32 matrices (+ transposes) of any width, omega channel-major
static int setup_generic(alg_handle* h, int gS, int gH, int gU, int gR, int gD) {
  const int L = h->L, T = h->T, B = h->B, nl = h->nl, ENVW = (L + 1) * gU;
  std::string err;
  struct Item { const float* src; long k, n; bool transpose; size_t off; };
  std::vector<Item> items;
  std::vector<std::pair<const float**, long>> fix;
  size_t total = 0;
  auto add = [&](const std::string& name, long k, long n, bool tr, const float** dst) -> bool {
    const float* p = tf32(h, name, k, n, err);
    if (!p) return false;
    items.push_back({p, k, n, tr, total});
    fix.push_back({dst, (long)items.size() - 1});
    total += ((size_t)k * n + 63) / 64 * 64;
    return true;
  };
  auto both = [&](const std::string& name, long k, long n, const float** w, const float** wt) { return add(name, k, n, false, w) && add(name, k, n, true, wt); };
  GenModel& gm = h->gm;
  gm = GenModel{};
  gm.S = gS; gm.H = gH; gm.U = gU; gm.R = gR; gm.L = L; gm.nl = nl; gm.T = T; gm.B = B; gm.depth = gD;
  auto mlp = [&](const std::string& pre, int din, GenMLP& m) -> bool {
    m.nlin = gD + 1;
    m.dims[0] = din;
    for (int i = 1; i <= gD; ++i) m.dims[i] = gH;
    m.dims[gD + 1] = gS;
    for (int i = 0; i < m.nlin; ++i)
      if (!both(pre + std::to_string(i), m.dims[i], m.dims[i + 1], &m.w[i], &m.wt[i])) return false;
    return true;
  };
  bool ok = mlp("twobody.w", 2 * T + B, gm.two) && both("embed_linear", gS, ENVW, &gm.emb, &gm.emb_t);
  static const char kinds[4][4] = {"", "A", "BA", "CDA"};
  for (int k = 0; k < nl && ok; ++k) {
    const std::string pre = "layer" + std::to_string(k) + ".";
    GenLayer& gl = gm.layer[k];
    gl.kind = kinds[nl][k];
    const GenTpDims td = gen_tp_dims(L, gl.kind);
    ok = ok && both(pre + "env_linear", gS, ENVW, &gl.env, &gl.env_t);
    ok = ok && add(pre + "omega", td.npath, gU, true, &gl.omega_t);          // [npath][U] -> [U][npath]
    ok = ok && mlp(pre + "mlp.w", gS + td.n0 * gU, gl.mlp);
    const float* al = ok ? tf32(h, pre + "alpha", 1, 1, err) : nullptr;
    if (!al) { ok = false; break; }
    const double alpha = al[0];
    gl.a = (float)(1.0 / std::sqrt(1.0 + alpha * alpha));
    gl.b = (float)(alpha / std::sqrt(1.0 + alpha * alpha));
  }
  ok = ok && both("readout.w0", gS, gR, &gm.ro0, &gm.ro0_t) && add("readout.w1", gR, 1, false, &gm.ro1);
  if (!ok) return fail(h, ALG_EIO, err);
  std::vector<float> blob(total, 0.f);
  for (const Item& it : items) {
    float* dst = blob.data() + it.off;
    if (!it.transpose) memcpy(dst, it.src, sizeof(float) * it.k * it.n);
    else for (long r = 0; r < it.k; ++r) for (long c = 0; c < it.n; ++c) dst[c * it.k + r] = it.src[r * it.n + c];
  }
  CK(h->gen_weights.ensure(total * sizeof(float)));
  CK(cudaMemcpy(h->gen_weights.p, blob.data(), total * sizeof(float), cudaMemcpyHostToDevice));
  for (auto& fx : fix) *fx.first = h->gen_weights.as<float>() + items[fx.second].off;
  // per-type tables and scalars (also needed when the tiled pipelines are not set up)
  auto getf64 = [&](const char* name, long n, std::vector<double>& out) -> bool {
    auto it = h->tensors.find(name);
    if (it == h->tensors.end() || it->second.dtype != "f64" || (long)it->second.count() != n) { err = std::string("bad tensor ") + name; return false; }
    out.assign(reinterpret_cast<const double*>(it->second.data.data()), reinterpret_cast<const double*>(it->second.data.data()) + n);
    return true;
  };
  if (!getf64("scales", T, h->scales) || !getf64("shifts", T, h->shifts) || !getf64("cutoff_table", (long)T * T, h->cut_table))
    return fail(h, ALG_EIO, err);
  const double inv = 1.0 / std::sqrt(h->avg_n);
  for (int i = 0; i < MAXT * MAXT; ++i) h->gtb.rc[i] = (float)h->r_max;
  for (int i = 0; i < T; ++i) for (int j = 0; j < T; ++j) h->gtb.rc[i * MAXT + j] = (float)h->cut_table[i * T + j];
  for (int i = 0; i < MAXT; ++i) h->gtb.gscale[i] = i < T ? (float)(inv * h->scales[i]) : 0.f;
  gm.p = (float)h->p; gm.inv_sqrt_n = (float)inv;
  {  // chunk size of the width-generic pipeline: about 8 GB of workspace, 16 k .. 1 M edges
    const double bytes_per_edge = 4.0 * (double)gen_work_floats(gm, 65536, 2048) / 65536.0;
    h->gen_chunk_edges = std::min<long>(1048576, std::max<long>(16384, (long)(8.0e9 / bytes_per_edge)));
  }
  CK(h->d_scale.ensure(sizeof(double) * MAXT));
  CK(h->d_shift.ensure(sizeof(double) * MAXT));
  CK(cudaMemcpy(h->d_scale.p, h->scales.data(), sizeof(double) * T, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_shift.p, h->shifts.data(), sizeof(double) * T, cudaMemcpyHostToDevice));
  return ALG_OK;
}

// ------------------------------------------------------------------------------------------
// model setup
// ------------------------------------------------------------------------------------------
static int setup_model(alg_handle* h) {
  auto& hd = h->header;
  auto geti = [&](const char* k, int def) { auto it = hd.find(k); return it == hd.end() ? def : atoi(it->second.c_str()); };
  auto getd = [&](const char* k, double def) { auto it = hd.find(k); return it == hd.end() ? def : atof(it->second.c_str()); };
  for (const char* k : {"r_max", "type_names", "num_types", "allow_tf32", "l_max", "num_layers", "num_bessels"})
    if (!hd.count(k)) return fail(h, ALG_EIO, std::string("weight file header misses key ") + k);
  h->L = geti("l_max", 0); h->nl = geti("num_layers", 0); h->T = geti("num_types", 0); h->B = geti("num_bessels", 0);
  h->r_max = getd("r_max", 0); h->avg_n = getd("avg_num_neighbors", 1.0); h->p = getd("polynomial_cutoff_p", 6.0);
  h->allow_tf32 = geti("allow_tf32", 0);
  h->type_names = hd["type_names"];
  if (h->L < 1 || h->L > 3) return fail(h, ALG_EINVAL, "unsupported l_max (supported: 1..3)");
  if (h->nl < 1 || h->nl > 3) return fail(h, ALG_EINVAL, "unsupported num_layers (supported: 1..3)");
  if (h->T < 1 || h->T > MAXT) return fail(h, ALG_EINVAL, "unsupported num_types (supported: 1..8)");
  if (h->B < 1 || h->B > MAXB) return fail(h, ALG_EINVAL, "unsupported num_bessels (supported: 1..16)");
  const int gS = geti("num_scalar_features", 0), gU = geti("num_tensor_features", 0), gH = geti("mlp_width", 0), gD = geti("mlp_depth", 0),
            gR = geti("readout_width", 0);
  if (gS < 1 || gU < 1 || gH < 1 || gR < 1 || gD < 1 || gD + 1 > GEN_MAXLIN || gS > 4096 || gU > 1024 || gH > 4096 || gR > 4096)
    return fail(h, ALG_EINVAL, "unsupported widths: num_scalar_features / num_tensor_features / mlp_width / readout_width must be positive, "
                               "mlp_depth in 1..4");
  // the tiled tensor-core / FP32-pipe kernels are specialised for these widths; every other model runs on the width-generic pipeline
  h->std_widths = gS == S && gU == U && gH == H && gD == 2 && gR == R;
  {
    std::istringstream ss(hd.count("per_edge_type_cutoff") ? hd["per_edge_type_cutoff"] : "");
    double v;
    while (ss >> v) h->per_edge_cut.push_back(v);
    if (!h->per_edge_cut.empty() && (int)h->per_edge_cut.size() != h->T * h->T)
      return fail(h, ALG_EIO, "per_edge_type_cutoff must hold num_types^2 values");
  }
  {
    int rc = setup_generic(h, gS, gH, gU, gR, gD);
    if (rc != ALG_OK) return rc;
  }
  if (!h->std_widths) {
    h->generic = true; h->use_tc = false; h->pipe = nullptr; h->fused_grid = 0;
    cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device);
    h->tensors.clear();
    return ALG_OK;
  }
  h->pipe = get_pipeline(h->L);
  h->pinfo = h->pipe->info(h->nl);
  const int L = h->L, T = h->T, B = h->B, ENVW = (L + 1) * U, SIN = S + ENVW;
  std::string err;
  // gather all fp32 matrices (+ transposes) into one device blob
  struct Item { const float* src; long k, n; bool transpose; size_t off; };
  std::vector<Item> items;
  size_t total = 0;
  auto add = [&](const std::string& name, long k, long n, bool tr) -> long {
    const float* p = tf32(h, name, k, n, err);
    if (!p) return -1;
    items.push_back({p, k, n, tr, total});
    long idx = (long)items.size() - 1;
    total += ((size_t)k * n + 63) / 64 * 64;
    return idx;
  };
  std::vector<std::pair<const float**, long>> fix;   // pointers to patch after upload
  ModelW& mw = h->mw;
  auto both = [&](const std::string& name, long k, long n, const float** w, const float** wt) -> bool {
    long a = add(name, k, n, false); if (a < 0) return false;
    long b = add(name, k, n, true); if (b < 0) return false;
    fix.push_back({w, a}); fix.push_back({wt, b});
    return true;
  };
  bool ok = true;
  ok = ok && both("twobody.w0", 2 * T + B, H, &mw.two.w[0], &mw.two.wt[0]);
  ok = ok && both("twobody.w1", H, H, &mw.two.w[1], &mw.two.wt[1]);
  ok = ok && both("twobody.w2", H, S, &mw.two.w[2], &mw.two.wt[2]);
  ok = ok && both("embed_linear", S, ENVW, &mw.emb, &mw.emb_t);
  static const char* kinds[4][3] = {{"", "", ""}, {"A", "", ""}, {"B", "A", ""}, {"C", "D", "A"}};
  (void)kinds;
  for (int k = 0; k < h->nl && ok; ++k) {
    const std::string pre = "layer" + std::to_string(k) + ".";
    ok = ok && both(pre + "env_linear", S, ENVW, &mw.layer[k].env, &mw.layer[k].env_t);
    auto it = h->tensors.find(pre + "omega");
    if (it == h->tensors.end() || it->second.dims.size() != 2 || it->second.dims[1] != U) { err = "bad tensor " + pre + "omega"; ok = false; break; }
    long a = add(pre + "omega", it->second.dims[0], U, false); if (a < 0) { ok = false; break; }
    fix.push_back({&mw.layer[k].omega, a});
    ok = ok && both(pre + "mlp.w0", SIN, H, &mw.layer[k].mlp.w[0], &mw.layer[k].mlp.wt[0]);
    ok = ok && both(pre + "mlp.w1", H, H, &mw.layer[k].mlp.w[1], &mw.layer[k].mlp.wt[1]);
    ok = ok && both(pre + "mlp.w2", H, S, &mw.layer[k].mlp.w[2], &mw.layer[k].mlp.wt[2]);
    const float* al = tf32(h, pre + "alpha", 1, 1, err);
    if (!al) { ok = false; break; }
    const double alpha = al[0];
    mw.layer[k].a = (float)(1.0 / std::sqrt(1.0 + alpha * alpha));
    mw.layer[k].b = (float)(alpha / std::sqrt(1.0 + alpha * alpha));
  }
  if (ok) {
    ok = ok && both("readout.w0", S, R, &mw.ro0, &mw.ro0_t);
    long a = add("readout.w1", R, 1, false); if (a < 0) ok = false; else fix.push_back({&mw.ro1, a});
  }
  if (!ok) return fail(h, ALG_EIO, err);
  std::vector<float> blob(total, 0.f);
  for (const Item& it : items) {
    float* dst = blob.data() + it.off;
    if (!it.transpose) memcpy(dst, it.src, sizeof(float) * it.k * it.n);
    else for (long r = 0; r < it.k; ++r) for (long c = 0; c < it.n; ++c) dst[c * it.k + r] = it.src[r * it.n + c];
  }
  CK(h->weights.ensure(total * sizeof(float)));
  CK(cudaMemcpy(h->weights.p, blob.data(), total * sizeof(float), cudaMemcpyHostToDevice));
  for (auto& fx : fix) *fx.first = h->weights.as<float>() + items[fx.second].off;
  // scalars
  auto getf64 = [&](const char* name, long n, std::vector<double>& out) -> bool {
    auto it = h->tensors.find(name);
    if (it == h->tensors.end() || it->second.dtype != "f64" || (long)it->second.count() != n) { err = std::string("bad tensor ") + name; return false; }
    out.assign(reinterpret_cast<const double*>(it->second.data.data()), reinterpret_cast<const double*>(it->second.data.data()) + n);
    return true;
  };
  if (!getf64("scales", T, h->scales) || !getf64("shifts", T, h->shifts) || !getf64("cutoff_table", (long)T * T, h->cut_table))
    return fail(h, ALG_EIO, err);
  const double inv = 1.0 / std::sqrt(h->avg_n);
  for (int i = 0; i < MAXT * MAXT; ++i) mw.rc[i] = (float)h->r_max;
  for (int i = 0; i < T; ++i) for (int j = 0; j < T; ++j) mw.rc[i * MAXT + j] = (float)h->cut_table[i * T + j];
  for (int i = 0; i < MAXT; ++i) mw.gscale[i] = i < T ? (float)(inv * h->scales[i]) : 0.f;
  mw.T = T; mw.B = B; mw.nl = h->nl; mw.p = (float)h->p; mw.inv_sqrt_n = (float)inv;
  CK(h->d_scale.ensure(sizeof(double) * MAXT));
  CK(h->d_shift.ensure(sizeof(double) * MAXT));
  CK(cudaMemcpy(h->d_scale.p, h->scales.data(), sizeof(double) * T, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_shift.p, h->shifts.data(), sizeof(double) * T, cudaMemcpyHostToDevice));
  h->pipe_ffma = h->pipe;
  CK(h->pipe->init());
  // ---- tensor-core path: pre-swizzled K-major hi/lo weight images (umma.cuh), one per GEMM
  h->pipe_tc = get_pipeline_tc(h->L);
  if (h->pipe_tc) {
    struct Img { TcMat* dst; int N, K; std::vector<float> hi, lo; size_t off; };
    std::vector<Img> imgs;
    auto T_ = [&](const std::string& name) { return reinterpret_cast<const float*>(h->tensors.at(name).data.data()); };
    auto make = [&](TcMat* dst, int N, int K, auto&& f /*(n,k)->float*/) {
      Img im; im.dst = dst; im.N = N; im.K = K; im.hi.assign((size_t)N * K, 0.f); im.lo.assign((size_t)N * K, 0.f);
      for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
          const float v = f(n, k);
          const float vh = umma::tf32_hi(v);
          const int o = umma::opk_idx(n, k, N);
          im.hi[o] = vh; im.lo[o] = v - vh;
        }
      imgs.push_back(std::move(im));
    };
    const float* t0 = T_("twobody.w0"); const float* t1 = T_("twobody.w1"); const float* t2 = T_("twobody.w2");
    const float* emb = T_("embed_linear"); const float* r0 = T_("readout.w0");
    TcW& tw = h->tcw;
    make(&tw.two0, H, 32, [&](int n, int k) { return k < B ? t0[(size_t)(2 * T + k) * H + n] : 0.f; });
    make(&tw.two1, H, H, [&](int n, int k) { return t1[(size_t)k * H + n]; });
    make(&tw.two2, S, H, [&](int n, int k) { return t2[(size_t)k * S + n]; });
    const int NBK = (ENVW + 63) / 64;                    // blocks of the l-indexed width
    auto bwid = [&](int b) { return std::min(64, ENVW - 64 * b); };
    for (int b = 0; b < NBK; ++b)
      make(&tw.emb[b], bwid(b), S, [&](int n, int k) { return emb[(size_t)k * ENVW + 64 * b + n]; });
    make(&tw.two2_b, H, S, [&](int n, int k) { return t2[(size_t)n * S + k]; });
    make(&tw.two1_b, H, H, [&](int n, int k) { return t1[(size_t)n * H + k]; });
    make(&tw.two0_b, 32, H, [&](int n, int k) { return n < B ? t0[(size_t)(2 * T + n) * H + k] : 0.f; });
    for (int b = 0; b < NBK; ++b)
      make(&tw.emb_b[b], S, bwid(b), [&](int n, int k) { return emb[(size_t)n * ENVW + 64 * b + k]; });
    make(&tw.ro0, R, S, [&](int n, int k) { return r0[(size_t)k * R + n]; });
    make(&tw.ro0_b, S, R, [&](int n, int k) { return r0[(size_t)n * R + k]; });
    for (int kk = 0; kk < h->nl; ++kk) {
      const std::string pre = "layer" + std::to_string(kk) + ".";
      const float* m0 = T_(pre + "mlp.w0"); const float* m1 = T_(pre + "mlp.w1"); const float* m2 = T_(pre + "mlp.w2");
      const float* env = T_(pre + "env_linear");
      TcLayerW& tl = tw.layer[kk];
      make(&tl.m0x, H, S, [&](int n, int k) { return m0[(size_t)k * H + n]; });
      for (int b = 0; b < NBK; ++b)
        make(&tl.m0s[b], H, bwid(b), [&](int n, int k) { return m0[(size_t)(S + 64 * b + k) * H + n]; });
      make(&tl.m1, H, H, [&](int n, int k) { return m1[(size_t)k * H + n]; });
      make(&tl.m2, S, H, [&](int n, int k) { return m2[(size_t)k * S + n]; });
      for (int b = 0; b < NBK; ++b)
        make(&tl.env[b], bwid(b), S, [&](int n, int k) { return env[(size_t)k * ENVW + 64 * b + n]; });
      make(&tl.m2_b, H, S, [&](int n, int k) { return m2[(size_t)n * S + k]; });
      make(&tl.m1_b, H, H, [&](int n, int k) { return m1[(size_t)n * H + k]; });
      make(&tl.m0_bx, S, H, [&](int n, int k) { return m0[(size_t)n * H + k]; });
      for (int b = 0; b < NBK; ++b)
        make(&tl.m0_bs[b], bwid(b), H, [&](int n, int k) { return m0[(size_t)(S + 64 * b + n) * H + k]; });
      for (int b = 0; b < NBK; ++b)
        make(&tl.env_b[b], S, bwid(b), [&](int n, int k) { return env[(size_t)n * ENVW + 64 * b + k]; });
    }
    size_t tot = 0;
    for (Img& im : imgs) { im.off = tot; tot += 2 * (((size_t)im.N * im.K + 255) / 256 * 256); }
    std::vector<float> tb(tot, 0.f);
    for (Img& im : imgs) {
      memcpy(tb.data() + im.off, im.hi.data(), sizeof(float) * im.hi.size());
      memcpy(tb.data() + im.off + ((size_t)im.N * im.K + 255) / 256 * 256, im.lo.data(), sizeof(float) * im.lo.size());
    }
    CK(h->tc_weights.ensure(tot * sizeof(float)));
    CK(cudaMemcpy(h->tc_weights.p, tb.data(), tot * sizeof(float), cudaMemcpyHostToDevice));
    for (Img& im : imgs) {
      im.dst->hi = h->tc_weights.as<float>() + im.off;
      im.dst->lo = im.dst->hi + ((size_t)im.N * im.K + 255) / 256 * 256;
      im.dst->N = im.N; im.dst->K = im.K;
    }
    tw.passes = 3;
    CK(h->pipe_tc->init());
    // default: tensor-core pipeline in strict (3xTF32) mode whenever the model is supported
    h->use_tc = true; h->pipe = h->pipe_tc; h->pinfo = h->pipe->info(h->nl);
    h->fused_grid = h->pipe_tc->fused_grid ? h->pipe_tc->fused_grid(h->nl) : 0;
    cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device);
  }
  h->tensors.clear();
  return ALG_OK;
}

// ------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------
extern "C" const char* alg_version(void) { return "allegro_b200 0.2.0 sm_100a"; }
extern "C" int alg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" const char* alg_last_error(const alg_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int alg_create(const char* weight_path, int cuda_device, alg_handle** out) {
  if (!out || !weight_path) { g_create_error = "alg_create: null argument"; return ALG_EINVAL; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this library has no CPU fallback";
    return ALG_ECUDA;
  }
  if (cuda_device < 0 || cuda_device >= ndev) { g_create_error = "alg_create: cuda_device out of range"; return ALG_EINVAL; }
  alg_handle* h = new alg_handle();
  h->device = cuda_device;
  std::string err;
  if (!read_alg(weight_path, h, err)) { g_create_error = err; delete h; return ALG_EIO; }
  int rc = ALG_OK;
  if ((e = cudaSetDevice(cuda_device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_create_error = std::string("CUDA error: ") + cudaGetErrorString(e); delete h; return ALG_ECUDA;
  }
  for (int i = 0; i < 4; ++i) cudaEventCreate(&h->ev[i]);
  cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
  for (int i = 0; i < alg_handle::INFO_RING; ++i) cudaEventCreateWithFlags(&h->ev_info[i], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_order, cudaEventDisableTiming);
  rc = setup_model(h);
  if (rc != ALG_OK) { g_create_error = h->err; alg_destroy(h); return rc; }
  *out = h;
  return ALG_OK;
}

extern "C" void alg_destroy(alg_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  for (auto& r : h->reg) if (r.first) { cudaHostUnregister(const_cast<void*>(r.first)); cudaGetLastError(); }
  for (cudaEvent_t e : {h->ev_info[0], h->ev_info[1], h->ev_info[2], h->ev_info[3], h->ev_copy, h->ev_order}) if (e) cudaEventDestroy(e);
  h->d_cplan.release(); h->d_sm_phase.release(); h->d_blk.release(); h->d_blk_base.release(); h->d_tile_c0.release(); h->d_info.release(); h->d_f_stage.release();
  h->h_info.release(); h->h_eatom.release();
  DevBuf* bufs[] = {&h->weights, &h->d_tmap, &h->d_cutsq, &h->d_scale, &h->d_shift, &h->d_x, &h->d_type, &h->d_ilist, &h->d_numneigh,
                    &h->d_cand, &h->d_first, &h->d_cnt, &h->d_rowptr, &h->d_scan_tmp, &h->d_mtype, &h->d_edge_j, &h->d_edge_c, &h->d_rvec,
                    &h->d_esum, &h->d_facc, &h->d_vacc, &h->d_forces, &h->d_eall, &h->d_red, &h->d_edge_index, &h->d_edge_energy,
                    &h->d_edge_grad, &h->d_eatom_out, &h->tc_weights, &h->c_ZD[0], &h->c_ZD[1], &h->c_ZD[2], &h->c_ZD[3], &h->c_W0, &h->c_dX, &h->c_dY, &h->c_du, &h->c_carry, &h->c_ecarry,
                    &h->c_X[0], &h->c_X[1], &h->c_X[2], &h->c_V[0], &h->c_V[1], &h->c_V[2], &h->c_dV[0], &h->c_dV[1],
                    &h->c_gamma[0], &h->c_gamma[1], &h->c_gamma[2], &h->c_dgamma[0], &h->c_dgamma[1], &h->c_dgamma[2], &h->gen_weights, &h->gen_work};
  for (DevBuf* b : bufs) b->release();
  h->h_stage.release(); h->h_rowptr.release(); h->h_out.release(); h->h_first.release();
  for (int i = 0; i < 4; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

extern "C" int alg_metadata(const alg_handle* h, double* r_max, int* num_types, const char** type_names,
                            const double** per_edge_type_cutoff, int* allow_tf32) {
  if (!h) return ALG_EINVAL;
  if (r_max) *r_max = h->r_max;
  if (num_types) *num_types = h->T;
  if (type_names) *type_names = h->type_names.c_str();
  if (per_edge_type_cutoff) *per_edge_type_cutoff = h->per_edge_cut.empty() ? nullptr : h->per_edge_cut.data();
  if (allow_tf32) *allow_tf32 = h->allow_tf32;
  return ALG_OK;
}

extern "C" int alg_set_type_map(alg_handle* h, int ntypes, const int* map, const double* cutoff_matrix) {
  if (!h) return ALG_EINVAL;
  if (ntypes < 1 || !map || !cutoff_matrix) return fail(h, ALG_EINVAL, "alg_set_type_map: bad arguments");
  for (int i = 0; i < ntypes; ++i)
    if (map[i] < -1 || map[i] >= h->T) return fail(h, ALG_EINVAL, "alg_set_type_map: model type index out of range");
  CK(cudaSetDevice(h->device));
  std::vector<double> c2((size_t)ntypes * ntypes);
  for (size_t i = 0; i < c2.size(); ++i) c2[i] = cutoff_matrix[i] * cutoff_matrix[i];   // cutij*cutij, cpp:507
  CK(h->d_tmap.ensure(sizeof(int) * ntypes));
  CK(h->d_cutsq.ensure(sizeof(double) * c2.size()));
  CK(cudaMemcpy(h->d_tmap.p, map, sizeof(int) * ntypes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_cutsq.p, c2.data(), sizeof(double) * c2.size(), cudaMemcpyHostToDevice));
  h->ntypes = ntypes;
  h->have_map = true;
  return ALG_OK;
}

extern "C" int alg_set_option(alg_handle* h, const char* key, const char* value) {
  if (!h || !key || !value) return ALG_EINVAL;
  const std::string k(key), v(value);
  if (k == "filter") {
    if (v == "le") h->filter_le = true; else if (v == "lt") h->filter_le = false; else return fail(h, ALG_EINVAL, "filter must be le or lt");
  } else if (k == "chunk_edges") {
    long c = atol(value);
    if (c < 4096) return fail(h, ALG_EINVAL, "chunk_edges must be >= 4096");
    h->chunk_edges = c;
  } else if (k == "keep_edges") h->keep_edges = v == "1";
  else if (k == "debug") h->debug = v == "1";
  else if (k == "profile") h->prof.on = v == "1";
  else if (k == "gemm") {
    if (v == "generic") { h->generic = true; h->use_tc = false; }
    else if (!h->std_widths) return fail(h, ALG_EINVAL, "gemm=" + v + ": the tiled kernels are specialised for num_scalar_features=64, num_tensor_features=32, "
                                                          "mlp 2x64, readout 32; this model runs on gemm=generic");
    else if (v == "ffma") { h->generic = false; h->use_tc = false; h->pipe = h->pipe_ffma; }
    else if (v == "tc") {
      if (!h->pipe_tc) return fail(h, ALG_EINVAL, "gemm=tc: no tensor-core pipeline for this model");
      h->generic = false; h->use_tc = true; h->pipe = h->pipe_tc;
    } else return fail(h, ALG_EINVAL, "gemm must be tc, ffma or generic");
    if (h->pipe) h->pinfo = h->pipe->info(h->nl);
  } else if (k == "pipeline") {
    if (v == "auto") h->pipeline_mode = 0;
    else if (v == "fused") {
      if (h->generic || !h->pipe_tc || !h->pipe_tc->run_fused || h->fused_grid <= 0) return fail(h, ALG_EINVAL, "pipeline=fused needs the tensor-core pipeline of this model");
      h->pipeline_mode = 1;
    } else if (v == "tiled") h->pipeline_mode = 2;
    else return fail(h, ALG_EINVAL, "pipeline must be auto, fused or tiled");
  } else if (k == "fused_batch") {
    const int b = atoi(v.c_str());
    if (b < 1 || b > 64) return fail(h, ALG_EINVAL, "fused_batch must be in 1..64");
    h->fused_batch = b;
  } else if (k == "chunk_plan") {
    if (v == "device") h->tiled_plan_device = true; else if (v == "host") h->tiled_plan_device = false;
    else return fail(h, ALG_EINVAL, "chunk_plan must be device or host");
  } else if (k == "phase_align") {
    h->phase_align = v != "0";
  } else if (k == "max_neighbors") {
    h->max_neighbors = atol(v.c_str());
    if (h->max_neighbors < 0) return fail(h, ALG_EINVAL, "max_neighbors must be >= 0");
  } else if (k == "host_register") {
    h->host_register = v != "0";
  } else if (k == "neigh_ago") {
    h->neigh_ago = atoi(v.c_str());
    if (h->neigh_ago < 0) return fail(h, ALG_EINVAL, "neigh_ago must be >= 0");
  } else if (k == "precision") {
    if (v == "strict") h->tcw.passes = 3; else if (v == "tf32") h->tcw.passes = 1; else return fail(h, ALG_EINVAL, "precision must be strict or tf32");
  }
  else return fail(h, ALG_ENOTFOUND, "unknown option " + k);
  return ALG_OK;
}

// ------------------------------------------------------------------------------------------
// the step
// ------------------------------------------------------------------------------------------
static int ensure_chunk_buffers(alg_handle* h, long max_tiles, long max_centres) {
  const PipelineInfo& pi = h->pinfo;
  const size_t TM = pi.TM;
  for (int k = 0; k < h->nl; ++k) CK(h->c_X[k].ensure(sizeof(float) * max_tiles * S * TM));
  CK(h->c_W0.ensure(sizeof(float) * max_tiles * pi.ENVW * TM));
  for (int k = 1; k < h->nl; ++k) CK(h->c_V[k].ensure(sizeof(float) * max_tiles * U * pi.vstride[k] * TM));
  CK(h->c_dX.ensure(sizeof(float) * max_tiles * S * TM));
  if (h->use_tc) for (int k = 0; k <= h->nl; ++k) CK(h->c_ZD[k].ensure(sizeof(float) * max_tiles * 3 * 64 * TM));   // stage 0 = two-body, 1+k = layer k
  if (h->nl > 1) for (int q = 0; q < 2; ++q) CK(h->c_dV[q].ensure(sizeof(float) * max_tiles * U * pi.dvstride * TM));
  CK(h->c_dY.ensure(sizeof(float) * max_tiles * pi.NSH * TM));
  CK(h->c_du.ensure(sizeof(float) * max_tiles * TM));
  for (int k = 0; k < h->nl; ++k) {
    CK(h->c_gamma[k].ensure(sizeof(float) * max_centres * pi.F));
    CK(h->c_dgamma[k].ensure(sizeof(float) * max_centres * pi.F));
  }
  CK(h->c_carry.ensure(sizeof(float) * max_tiles * pi.F));
  CK(h->c_ecarry.ensure(sizeof(double) * max_tiles));
  return ALG_OK;
}

// row4 = the float4-packed row layout the tensor-core pipeline uses for x^k / dX (allegro_kernels_tc.cuh: st_row4)
static void detile(const std::vector<float>& raw, int ntiles, int rows, int TM, long E, std::vector<double>& out, bool row4 = false) {
  out.assign((size_t)E * rows, 0.0);
  for (long e = 0; e < E; ++e) {
    const long t = e / TM, m = e % TM;
    for (int r = 0; r < rows; ++r)
      out[(size_t)e * rows + r] = row4 ? raw[(size_t)t * rows * TM + ((size_t)(r >> 2) * TM + m) * 4 + (r & 3)] : raw[((size_t)t * rows + r) * TM + m];
  }
}

// everything one force evaluation needs, on the device (d_* = device pointers); acc describes the device-resident
// neighbour list.  Host-API extras: the caller's f is read-modified-written through d_f_inout, then copied back.
struct StepIO {
  int nlocal, nghost;
  const double* d_x; const int* d_type; const int* d_ilist;
  NeighAcc acc;
  int eflag_atom, vflag_global;
  double* d_f_inout; double* d_eatom;
  long cap_hint;                 // upper bound on the number of edges (0 = unknown)
  cudaEvent_t wait_f;            // host API: d_f_inout is being uploaded on another stream; wait for this before the store
  double* h_f; double* h_eatom_stage;   // host API: copy d_f_inout / d_eatom back (async, before the final synchronisation)
};

static void fill_args(alg_handle* h, const StepIO& io, ChunkArgs& a) {
  a = ChunkArgs{};
  a.rvec = h->d_rvec.as<float4>(); a.edge_j = h->d_edge_j.as<int>(); a.edge_c = h->d_edge_c.as<int>();
  a.rowptr = h->d_rowptr.as<int>(); a.ilist = io.d_ilist;
  for (int k = 0; k < 3; ++k) { a.X[k] = h->c_X[k].as<float>(); a.V[k] = h->c_V[k].as<float>(); a.gamma[k] = h->c_gamma[k].as<float>(); a.dgamma[k] = h->c_dgamma[k].as<float>(); }
  for (int k = 0; k < 4; ++k) a.ZD[k] = h->c_ZD[k].as<float>();
  a.W0 = h->c_W0.as<float>(); a.dX = h->c_dX.as<float>(); a.dV[0] = h->c_dV[0].as<float>(); a.dV[1] = h->c_dV[1].as<float>();
  a.dY = h->c_dY.as<float>(); a.du = h->c_du.as<float>(); a.carry = h->c_carry.as<float>(); a.ecarry = h->c_ecarry.as<double>();
  a.esum = h->d_esum.as<double>();
  a.edge_energy = h->debug ? h->d_edge_energy.as<float>() : nullptr;
  a.edge_grad = h->debug ? h->d_edge_grad.as<float>() : nullptr;
  a.facc = h->d_facc.as<unsigned long long>();
  a.vacc = h->d_vacc.as<unsigned long long>();
}

static int ensure_edge_arrays(alg_handle* h, long cap) {
  cap = std::max<long>(cap, 1);
  CK(h->d_edge_j.ensure(sizeof(int) * cap));
  CK(h->d_edge_c.ensure(sizeof(int) * cap));
  CK(h->d_rvec.ensure(sizeof(float4) * cap));
  if (h->keep_edges) CK(h->d_edge_index.ensure(sizeof(long long) * 2 * cap));
  if (h->debug) {
    CK(h->d_edge_energy.ensure(sizeof(float) * cap));
    CK(h->d_edge_grad.ensure(sizeof(float) * 3 * cap));
  }
  return ALG_OK;
}

static void launch_edge_fill(alg_handle* h, const StepIO& io, long cap) {
  cudaStream_t st = h->stream;
  const int wblocks = (int)(((long)io.nlocal * 32 + 255) / 256);
  if (!io.acc.flat && io.acc.stride_i == 1 && io.acc.stride_jj > 1)
    k_edges_tpa<true><<<(io.nlocal + 127) / 128, 128, 0, st>>>(io.nlocal, io.d_x, io.d_type, io.d_ilist, io.acc, h->d_cutsq.as<double>(), h->ntypes, h->filter_le ? 1 : 0,
                                                               nullptr, h->d_rowptr.as<int>(), h->d_mtype.as<int>(), h->d_edge_j.as<int>(), h->d_edge_c.as<int>(),
                                                               h->d_rvec.as<float4>(), cap);
  else
    k_edges<true><<<wblocks, 256, 0, st>>>(io.nlocal, io.d_x, io.d_type, io.d_ilist, io.acc, h->d_cutsq.as<double>(), h->ntypes, h->filter_le ? 1 : 0,
                                           nullptr, h->d_rowptr.as<int>(), h->d_mtype.as<int>(), h->d_edge_j.as<int>(), h->d_edge_c.as<int>(),
                                           h->d_rvec.as<float4>(), cap);
  if (h->keep_edges)
    k_edge_index<<<1024, 256, 0, st>>>(h->d_rowptr.as<int>(), io.nlocal, cap, h->d_edge_j.as<int>(), h->d_edge_c.as<int>(), io.d_ilist,
                                       h->d_edge_index.as<long long>());
}

// ---- chunked edge-tile pipeline (any neighbour count; synchronises once on the CSR row pointer)
static int step_tiled(alg_handle* h, const StepIO& io) {
  cudaStream_t st = h->stream;
  const int nlocal = io.nlocal;
  const PipelineInfo& pi = h->pinfo;
  const int TM = h->generic ? 128 : pi.TM;
  CK(h->h_rowptr.ensure(sizeof(int) * (nlocal + 1)));
  CK(cudaMemcpyAsync(h->h_rowptr.p, h->d_rowptr.p, sizeof(int) * (nlocal + 1), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const int* rowptr = h->h_rowptr.as<int>();
  const long E = rowptr[nlocal];
  h->last_E = E; h->last_E_known = E; h->last_E_nlocal = nlocal;
  int rc = ensure_edge_arrays(h, E);
  if (rc != ALG_OK) return rc;
  launch_edge_fill(h, io, E);
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->ev[1], st));
  // chunk plan (centre aligned)
  struct Chunk { int c0, c1, e0, e1; };
  std::vector<Chunk> chunks;
  const long CE = std::max<long>(h->generic ? std::min<long>(h->chunk_edges, h->gen_chunk_edges) : h->chunk_edges, TM);
  const long CC = CE;   // centres per chunk bound
  long max_tiles = 1, max_cent = 1;
  for (int c0 = 0; c0 < nlocal;) {
    // largest c1 with rowptr[c1]-rowptr[c0] <= CE and c1-c0 <= CC
    const long lim = (long)rowptr[c0] + CE;
    int c1 = (int)(std::upper_bound(rowptr + c0, rowptr + nlocal + 1, (int)std::min<long>(lim, 0x7fffffff)) - rowptr) - 1;
    if (c1 - c0 > CC) c1 = c0 + (int)CC;
    if (c1 <= c0) {
      if ((long)rowptr[c0 + 1] - rowptr[c0] > CE)
        return fail(h, ALG_EINVAL, "a single atom has more neighbours than chunk_edges; raise option chunk_edges");
      c1 = c0 + 1;
    }
    if (rowptr[c1] > rowptr[c0]) {
      chunks.push_back({c0, c1, rowptr[c0], rowptr[c1]});
      max_tiles = std::max<long>(max_tiles, ((long)rowptr[c1] - rowptr[c0] + TM - 1) / TM);
      max_cent = std::max<long>(max_cent, c1 - c0);
    }
    c0 = c1;
  }
  if (h->generic) {                                   // width-generic pipeline: own workspace, one chunk after the other
    long nmax = 1, cmax = 1;
    for (const Chunk& c : chunks) { nmax = std::max<long>(nmax, c.e1 - c.e0); cmax = std::max<long>(cmax, c.c1 - c.c0); }
    CK(h->gen_work.ensure(sizeof(float) * gen_work_floats(h->gm, nmax, cmax)));
    ChunkArgs a;
    fill_args(h, io, a);
    GenEdges ge{a.rvec, a.edge_j, a.edge_c, a.rowptr, a.ilist, a.esum, a.edge_energy, a.edge_grad, a.facc, a.vacc};
    for (const Chunk& c : chunks) CK(gen_run_chunk(h->gm, h->gtb, ge, c.c0, c.c1, c.e0, c.e1, h->gen_work.as<float>(), st, &h->prof.launches));
    h->step_stats[1] = (double)E; h->step_stats[2] = (double)chunks.size(); h->step_stats[3] = 0;
    h->dbg_ntiles = 0; h->dbg_c0 = 0; h->dbg_ncent = 0;
    return ALG_OK;
  }
  rc = ensure_chunk_buffers(h, max_tiles, max_cent);
  if (rc != ALG_OK) return rc;
  ChunkArgs a;
  fill_args(h, io, a);
  long tiles_total = 0;
  for (const Chunk& c : chunks) {
    a.e0 = c.e0; a.e1 = c.e1; a.c0 = c.c0;
    const int ntiles = (c.e1 - c.e0 + TM - 1) / TM;
    tiles_total += ntiles;
    CK(h->pipe->run_chunk(a, h->mw, &h->tcw, ntiles, st, &h->prof));
  }
  h->step_stats[1] = (double)E; h->step_stats[2] = (double)chunks.size(); h->step_stats[3] = (double)tiles_total;
  h->dbg_ntiles = chunks.size() == 1 ? (chunks[0].e1 - chunks[0].e0 + TM - 1) / TM : 0;
  h->dbg_c0 = chunks.size() == 1 ? chunks[0].c0 : 0;
  h->dbg_ncent = chunks.size() == 1 ? chunks[0].c1 - chunks[0].c0 : 0;
  return ALG_OK;
}

// ---- chunked pipeline with the device-built plan: no host synchronisation.  The grids cover the largest possible chunk
// and the number of launches the largest possible edge count (`cap`); CTAs / launches beyond the real plan exit at once.
constexpr int CHUNK_SLACK = 4096;                    // a chunk may exceed Q by the neighbour count of one atom
static int step_tiled_async(alg_handle* h, const StepIO& io, long cap) {
  cudaStream_t st = h->stream;
  const int nlocal = io.nlocal;
  const PipelineInfo& pi = h->pinfo;
  const int TM = pi.TM;
  int rc = ensure_edge_arrays(h, cap);
  if (rc != ALG_OK) return rc;
  const long Q = std::max<long>(h->chunk_edges, 4 * TM);
  const int nch_max = (int)((cap + Q - 1) / Q);
  const int max_edges = (int)Q + CHUNK_SLACK;
  const long max_tiles = (max_edges + TM - 1) / TM;
  const long max_cent = std::min<long>(std::max<long>(Q / 4, 1024), nlocal);
  CK(h->d_cplan.ensure(sizeof(int) * 3 * std::max(nch_max, 1)));
  k_chunk_plan<<<(nch_max + 127) / 128, 128, 0, st>>>(nlocal, h->d_rowptr.as<int>(), Q, nch_max, max_edges, (int)max_cent, cap,
                                                      h->d_cplan.as<int>(), h->d_info.as<int>());
  launch_edge_fill(h, io, cap);
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->ev[1], st));
  rc = ensure_chunk_buffers(h, max_tiles, max_cent);
  if (rc != ALG_OK) return rc;
  ChunkArgs a;
  fill_args(h, io, a);
  a.plan = h->d_cplan.as<int>();
  for (int ci = 0; ci < nch_max; ++ci) {
    a.ci = ci;
    CK(h->pipe->run_chunk(a, h->mw, &h->tcw, (int)max_tiles, st, &h->prof));
  }
  h->prof.launches += 1;                               // k_chunk_plan
  h->step_stats[2] = (double)nch_max; h->step_stats[3] = -1;
  h->dbg_ntiles = 0; h->dbg_c0 = 0; h->dbg_ncent = 0;
  return ALG_OK;
}

// ---- fused persistent pipeline: no host synchronisation; the tile plan is built on the device
static int step_fused(alg_handle* h, const StepIO& io, long cap) {
  cudaStream_t st = h->stream;
  const int nlocal = io.nlocal;
  const PipelineInfo& pi = h->pinfo;
  int rc = ensure_edge_arrays(h, cap);
  if (rc != ALG_OK) return rc;
  const int nblk = (nlocal + PLAN_CB - 1) / PLAN_CB;
  CK(h->d_blk.ensure(sizeof(int) * (nblk + 1)));
  CK(h->d_blk_base.ensure(sizeof(int) * (nblk + 1)));
  CK(h->d_tile_c0.ensure(sizeof(int) * ((size_t)nlocal + nblk + 2)));
  CK(cudaMemsetAsync(h->d_info.p, 0, sizeof(int) * 8, st));
  const int pb = (nblk + PLAN_TPB - 1) / PLAN_TPB;
  k_plan<false><<<pb, 256, 0, st>>>(nlocal, h->d_rowptr.as<int>(), h->d_blk.as<int>(), nullptr, nullptr, h->d_info.as<int>(), cap, h->fused_batch * 128);
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, h->d_blk.as<int>(), h->d_blk_base.as<int>(), nblk + 1, st);
  CK(h->d_scan_tmp.ensure(tmp_bytes));
  CK(cub::DeviceScan::ExclusiveSum(h->d_scan_tmp.p, tmp_bytes, h->d_blk.as<int>(), h->d_blk_base.as<int>(), nblk + 1, st));
  k_plan<true><<<pb, 256, 0, st>>>(nlocal, h->d_rowptr.as<int>(), nullptr, h->d_blk_base.as<int>(), h->d_tile_c0.as<int>(), h->d_info.as<int>(), cap, h->fused_batch * 128);
  launch_edge_fill(h, io, cap);
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->ev[1], st));
  const int grid = h->fused_grid;
  const int batch = h->fused_batch;
  rc = ensure_chunk_buffers(h, (long)grid * batch, (long)grid * batch * 128);
  if (rc != ALG_OK) return rc;
  ChunkArgs a;
  fill_args(h, io, a);
  a.e0 = 0; a.e1 = 0; a.c0 = 0;
  CK(h->d_sm_phase.ensure(sizeof(unsigned) * 1024));
  CK(cudaMemsetAsync(h->d_sm_phase.p, 0, sizeof(unsigned) * 1024, st));
  FusedPlan plan{h->d_tile_c0.as<int>(), h->d_info.as<int>(), batch, h->num_sms, h->phase_align ? h->d_sm_phase.as<unsigned>() : nullptr};      // info[5] (tile queue) was zeroed above
  CK(h->pipe->run_fused(a, h->mw, &h->tcw, plan, grid, st, &h->prof));
  h->prof.launches += 2;                               // k_plan x2
  h->dbg_ntiles = 0; h->dbg_c0 = 0; h->dbg_ncent = 0;
  (void)pi;
  return ALG_OK;
}

// the verdict of asynchronous steps is read here: non-blocking at the start of a call (every record whose copy has landed),
// blocking in the getters and when the ring is full
static int resolve_pending(alg_handle* h, bool block = true, unsigned keep = 0) {
  while (h->info_head - h->info_tail > keep) {
    const int slot = (int)(h->info_tail % alg_handle::INFO_RING);
    if (block || h->info_head - h->info_tail >= (unsigned)alg_handle::INFO_RING) CK(cudaEventSynchronize(h->ev_info[slot]));
    else {
      cudaError_t q = cudaEventQuery(h->ev_info[slot]);
      if (q == cudaErrorNotReady) { cudaGetLastError(); break; }
      CK(q);
    }
    ++h->info_tail;
    const int* info = h->h_info.as<int>() + 16 * slot;
    if (info[8] != 0)
      return fail(h, ALG_EINVAL, "an earlier asynchronous step met atom " + std::to_string(info[8] - 1) + " whose LAMMPS type has no model type");
    if (h->info_mode[slot] == 0) continue;
    h->last_E = info[2]; h->last_E_known = info[2];
    h->step_stats[1] = info[2]; h->step_stats[2] = info[0]; h->step_stats[3] = h->info_mode[slot] == 1 ? info[6] : (info[2] + 127) / 128;
    if (info[9]) {
      h->tiled_sync = true;
      return fail(h, ALG_ESTATE, "an earlier asynchronous alg_compute_device step needed a chunk larger than its buffers (an atom with more than 4096 "
                                 "neighbours, or fewer than 4 neighbours per atom on average) and produced no forces; the host-built chunk plan "
                                 "is selected from now on");
    }
    if (h->info_mode[slot] == 1 && info[1] > h->fused_batch * 128) {
      h->force_tiled = true;
      return fail(h, ALG_ESTATE, "an earlier asynchronous alg_compute_device step met an atom with more than fused_batch*128 neighbours inside the cutoff "
                                 "and produced no forces; the chunked pipeline is selected from now on (pass eng != NULL to have such steps "
                                 "re-run transparently)");
    }
    if (info[3]) return fail(h, ALG_ESTATE, "an earlier asynchronous alg_compute_device step overflowed its edge buffers and produced no forces "
                                            "(set option max_neighbors, or pass eng != NULL to have such steps re-run transparently)");
  }
  return ALG_OK;
}

static bool fused_selected(const alg_handle* h) {
  if (h->pipeline_mode == 2 || h->generic || !h->use_tc || !h->pipe || !h->pipe->run_fused || h->fused_grid <= 0) return false;
  if (h->pipeline_mode == 1) return true;
  // auto: the faster pipeline as measured on the B200 (bench.py, strict fp32) -- the chunked per-phase kernels: 4 % faster
  // for l_max = 1, 1.25x for l_max = 2, 1.5x for l_max = 3 (the persistent kernel streams the code of all phases through
  // the instruction cache and ptxas allocates a stand-alone phase kernel better than the same phase inlined into one
  // body).  With the device-built chunk plan (step_tiled_async) neither pipeline synchronises with the host, so the
  // fused kernel's remaining advantage is the launch count (10 vs ~600 per step); it stays selectable (pipeline=fused).
  return false;
}

static int run_step(alg_handle* h, const StepIO& io, double* eng, double* virial6) {
  cudaStream_t st = h->stream;
  const int nlocal = io.nlocal, ntot = io.nlocal + io.nghost;
  const bool want_sync = eng || virial6 || io.h_f;
  int rc = resolve_pending(h, want_sync, want_sync ? 0 : alg_handle::INFO_RING - 1);   // asynchronous call: only free one ring slot
  if (rc != ALG_OK) return rc;
  h->last_nlocal = nlocal; h->last_ntot = ntot; h->last_E = 0;
  h->outputs.clear();
  CK(cudaEventRecord(h->ev[0], st));
  // ---- K1: count + scan (the fill runs inside the pipeline step once the capacity is known)
  CK(h->d_cnt.ensure(sizeof(int) * (nlocal + 1)));
  CK(h->d_rowptr.ensure(sizeof(int) * (nlocal + 1)));
  CK(h->d_mtype.ensure(sizeof(int) * ntot));
  const int wblocks = (int)(((long)nlocal * 32 + 255) / 256);
  CK(h->d_info.ensure(sizeof(int) * 16));              // [0..3] fused tile plan (see k_plan), [8] 1 + index of an atom without model type
  CK(cudaMemsetAsync(h->d_info.p, 0, sizeof(int) * 16, st));
  k_mtype<<<(ntot + 255) / 256, 256, 0, st>>>(ntot, io.d_type, h->d_tmap.as<int>(), h->ntypes, h->d_mtype.as<int>(), h->d_info.as<int>() + 8);
  const bool tpa = !io.acc.flat && io.acc.stride_i == 1 && io.acc.stride_jj > 1;     // KOKKOS LayoutLeft neighbour view
  if (tpa)
    k_edges_tpa<false><<<(nlocal + 127) / 128, 128, 0, st>>>(nlocal, io.d_x, io.d_type, io.d_ilist, io.acc, h->d_cutsq.as<double>(), h->ntypes, h->filter_le ? 1 : 0,
                                                              h->d_cnt.as<int>(), nullptr, h->d_mtype.as<int>(), nullptr, nullptr, nullptr, 0);
  else
    k_edges<false><<<wblocks, 256, 0, st>>>(nlocal, io.d_x, io.d_type, io.d_ilist, io.acc, h->d_cutsq.as<double>(), h->ntypes, h->filter_le ? 1 : 0,
                                            h->d_cnt.as<int>(), nullptr, h->d_mtype.as<int>(), nullptr, nullptr, nullptr, 0);
  CK(cudaMemsetAsync(h->d_cnt.as<int>() + nlocal, 0, sizeof(int), st));
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, h->d_cnt.as<int>(), h->d_rowptr.as<int>(), nlocal + 1, st);
  CK(h->d_scan_tmp.ensure(tmp_bytes));
  CK(cub::DeviceScan::ExclusiveSum(h->d_scan_tmp.p, tmp_bytes, h->d_cnt.as<int>(), h->d_rowptr.as<int>(), nlocal + 1, st));
  CK(h->d_esum.ensure(sizeof(double) * nlocal));
  CK(h->d_facc.ensure(sizeof(unsigned long long) * 3 * ntot));
  CK(h->d_vacc.ensure(sizeof(unsigned long long) * 8));
  CK(h->d_forces.ensure(sizeof(double) * 3 * ntot));
  CK(h->d_eall.ensure(sizeof(double) * ntot));
  const int eblocks = (nlocal + 1023) / 1024;
  CK(h->d_red.ensure(sizeof(double) * (eblocks + 8)));
  CK(h->h_out.ensure(sizeof(double) * 8));
  CK(h->h_info.ensure(sizeof(int) * 16 * alg_handle::INFO_RING));
  bool fused = fused_selected(h);
  long cap = 0;
  // chunked pipeline: device-built chunk plan (no host synchronisation) unless intermediates are wanted (debug=1 needs the
  // host-side chunk bookkeeping) or an earlier step showed that the plan does not fit the buffers
  auto tiled_async = [&] { return !fused && !h->generic && !h->debug && !h->tiled_sync && h->tiled_plan_device; };
  for (int attempt = 0;; ++attempt) {
    CK(cudaMemsetAsync(h->d_esum.p, 0, sizeof(double) * nlocal, st));
    CK(cudaMemsetAsync(h->d_facc.p, 0, sizeof(unsigned long long) * 3 * ntot, st));
    CK(cudaMemsetAsync(h->d_vacc.p, 0, sizeof(unsigned long long) * 8, st));
    h->prof.reset();
    const bool tasync = tiled_async();
    if (fused || tasync) {
      if (cap == 0) {
        if (io.cap_hint > 0) cap = io.cap_hint;
        else if (h->last_E_known >= 0 && h->last_E_nlocal == nlocal) cap = h->last_E_known + h->last_E_known / 8 + 4096;
        else {                                           // first step on this system: read the edge count once
          int E0 = 0;
          CK(cudaMemcpyAsync(&E0, h->d_rowptr.as<int>() + nlocal, sizeof(int), cudaMemcpyDeviceToHost, st));
          CK(cudaStreamSynchronize(st));
          cap = (long)E0 + E0 / 8 + 4096;
          h->last_E_known = E0;
        }
        h->last_E_nlocal = nlocal;
      }
      rc = fused ? step_fused(h, io, cap) : step_tiled_async(h, io, cap);
    } else {
      rc = step_tiled(h, io);
    }
    if (rc != ALG_OK) return rc;
    h->last_fused = fused;
    const int mode = fused ? 1 : (tasync ? 2 : 0);
    h->last_mode = mode;
    // own kernels outside the pipeline: k_mtype, k_edges x2, [k_edge_index], 4 finalize kernels
    h->step_stats[0] = (double)h->prof.launches + 3 + (h->keep_edges ? 1 : 0) + 4;
    CK(cudaEventRecord(h->ev[2], st));
    // ---- finalize
    if (io.wait_f) CK(cudaStreamWaitEvent(st, io.wait_f, 0));
    k_forces<<<(unsigned)((3L * ntot + 255) / 256), 256, 0, st>>>(ntot, h->d_facc.as<unsigned long long>(), h->d_forces.as<double>(), io.d_f_inout,
                                                                  mode ? h->d_info.as<int>() : nullptr, fused ? h->fused_batch * 128 : 0x7fffffff);
    k_eall_ghost<<<(ntot + 255) / 256, 256, 0, st>>>(ntot, h->d_mtype.as<int>(), h->d_shift.as<double>(), h->d_eall.as<double>());
    k_eall_local<<<eblocks, 256, 0, st>>>(nlocal, io.d_ilist, h->d_mtype.as<int>(), h->d_esum.as<double>(), h->d_scale.as<double>(),
                                          h->d_shift.as<double>(), 1.0 / std::sqrt(h->avg_n), h->d_eall.as<double>(),
                                          io.eflag_atom ? io.d_eatom : nullptr, h->d_red.as<double>() + 8);
    k_final_scalars<<<1, 32, 0, st>>>(eblocks, h->d_red.as<double>() + 8, h->d_vacc.as<unsigned long long>(), h->d_red.as<double>());
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev[3], st));
    if (!want_sync) {
      const int slot = (int)(h->info_head % alg_handle::INFO_RING);
      CK(cudaMemcpyAsync(h->h_info.as<int>() + 16 * slot, h->d_info.p, sizeof(int) * 16, cudaMemcpyDeviceToHost, st));
      CK(cudaEventRecord(h->ev_info[slot], st));
      h->info_mode[slot] = mode;
      ++h->info_head;
      if (mode) { h->step_stats[1] = -1; h->step_stats[3] = -1; }
      return ALG_OK;
    }
    CK(cudaMemcpyAsync(h->h_info.p, h->d_info.p, sizeof(int) * 16, cudaMemcpyDeviceToHost, st));
    if (io.h_f) CK(cudaMemcpyAsync(io.h_f, io.d_f_inout, sizeof(double) * 3 * ntot, cudaMemcpyDeviceToHost, st));
    if (io.h_eatom_stage && io.eflag_atom) CK(cudaMemcpyAsync(io.h_eatom_stage, io.d_eatom, sizeof(double) * ntot, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h->h_out.p, h->d_red.p, sizeof(double) * 7, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h->h_info.as<int>()[8] != 0)
      return fail(h, ALG_EINVAL, "atom " + std::to_string(h->h_info.as<int>()[8] - 1) + " has a LAMMPS type that has no model type "
                                 "(its name in pair_coeff matches no type name of the model)");
    if (mode) {
      const int* info = h->h_info.as<int>();
      h->last_E = info[2]; h->last_E_known = info[2];
      h->step_stats[1] = info[2]; h->step_stats[2] = info[0]; h->step_stats[3] = fused ? info[6] : (info[2] + 127) / 128;
      if (info[9]) {                                     // a chunk of the device-built plan does not fit the buffers: host-built plan from now on
        h->tiled_sync = true;
        continue;
      }
      if (fused && info[1] > h->fused_batch * 128) {     // an atom with more neighbours than a batch holds: chunked pipeline from now on
        if (h->pipeline_mode == 1) return fail(h, ALG_EINVAL, "pipeline=fused: an atom has more than fused_batch*128 neighbours inside the cutoff");
        h->force_tiled = true; fused = false;
        continue;
      }
      if (info[3]) {                                     // edge arrays too small (capacity hysteresis): grow and repeat
        if (attempt >= 2) return fail(h, ALG_ECUDA, "edge buffer capacity could not be established");
        cap = (long)info[2] + info[2] / 8 + 4096;
        continue;
      }
    }
    break;
  }
  const double* o = h->h_out.as<double>();
  if (eng) *eng = o[0];
  if (virial6 && io.vflag_global) for (int q = 0; q < 6; ++q) virial6[q] = o[1 + q];
  float ms;
  for (int q = 0; q < 3; ++q) { cudaEventElapsedTime(&ms, h->ev[q], h->ev[q + 1]); h->timings[q] = ms; }
  if (h->prof.on) {
    for (int q = 0; q < KID_COUNT; ++q) { h->kernel_ms[q] = 0; h->kernel_n[q] = 0; }
    for (size_t r = 0; r < h->prof.ids.size(); ++r) {
      cudaEventElapsedTime(&ms, h->prof.ev[2 * r], h->prof.ev[2 * r + 1]);
      h->kernel_ms[h->prof.ids[r]] += ms; h->kernel_n[h->prof.ids[r]] += 1;
    }
  }
  return ALG_OK;
}

// pin a caller-owned host array for asynchronous DMA (one cached registration per role; re-registered when the
// caller's array moves or grows).  Returns false when the range cannot be pinned: the copy is then a plain
// (driver-staged) cudaMemcpyAsync.
static bool pin_host(alg_handle* h, int role, const void* p, size_t bytes) {
  if (!h->host_register || !p || bytes == 0) return false;
  auto& slot = h->reg[role];
  if (slot.first == p && slot.second >= bytes) return true;
  if (slot.first) { cudaHostUnregister(const_cast<void*>(slot.first)); cudaGetLastError(); slot = {nullptr, 0}; }
  cudaError_t e = cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return true; }   // already pinned by the caller
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  slot = {p, bytes};
  return true;
}

extern "C" int alg_compute_host(alg_handle* h, int nlocal, int nghost, const double* x, const int* type, const int* ilist,
                                const int* numneigh, int* const* firstneigh, int eflag_atom, int vflag_global,
                                double* f, double* eatom, double* eng, double* virial6) {
  if (!h) return ALG_EINVAL;
  if (!h->have_map) return fail(h, ALG_ESTATE, "alg_compute_host called before alg_set_type_map");
  if (nlocal < 0 || nghost < 0) return fail(h, ALG_EINVAL, "negative atom count");
  if (eng) *eng = 0.0;
  if (nlocal == 0) return ALG_OK;                   // empty domain: silent no-op (cpp:341)
  if (!x || !type || !ilist || !numneigh || !firstneigh || !f) return fail(h, ALG_EINVAL, "alg_compute_host: null pointer");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const int ntot = nlocal + nghost;
  const auto t_call = std::chrono::steady_clock::now();
  auto ms_since = [](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
  h->host_ms[0] = 0;
  CK(h->d_x.ensure(sizeof(double) * 3 * ntot));
  CK(h->d_type.ensure(sizeof(int) * ntot));
  CK(h->d_f_stage.ensure(sizeof(double) * 3 * ntot));
  // the caller's arrays (LAMMPS atom->x / f / type keep their address until they grow) are pinned once, so that the
  // per-step copies are asynchronous DMA instead of driver-staged pageable copies
  pin_host(h, 0, x, sizeof(double) * 3 * ntot);
  pin_host(h, 1, f, sizeof(double) * 3 * ntot);
  pin_host(h, 2, type, sizeof(int) * ntot);
  h->host_ms[2] = ms_since(t_call);
  CK(cudaMemcpyAsync(h->d_x.p, x, sizeof(double) * 3 * ntot, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(h->d_type.p, type, sizeof(int) * ntot, cudaMemcpyHostToDevice, st));
  // f travels to the device on a second stream while the network runs; the store kernel adds the model forces to it
  // (f += forces for ALL atoms incl. ghosts, cpp:370-377) and the sum is copied back before the final synchronisation
  CK(cudaMemcpyAsync(h->d_f_stage.p, f, sizeof(double) * 3 * ntot, cudaMemcpyHostToDevice, h->copy_stream));
  CK(cudaEventRecord(h->ev_copy, h->copy_stream));
  // LAMMPS rebuilds the neighbour list only every few steps (neighbor->ago == 0 on a rebuild step): with
  // option neigh_ago > 0 the device copy of the list uploaded by the last call is reused when the atom
  // counts still match; otherwise the paged list is flattened (ilist order, jlist order) and uploaded
  const bool reuse = h->neigh_ago > 0 && h->list_nlocal == nlocal && h->list_ntot == ntot && h->list_tot >= 0;
  if (!reuse) {
    const auto t_list = std::chrono::steady_clock::now();
    CK(h->h_first.ensure(sizeof(long long) * (nlocal + 1) + sizeof(int) * nlocal));
    long long* first = h->h_first.as<long long>();
    int* cnt = reinterpret_cast<int*>(first + nlocal + 1);
    long long tot = 0;
    for (int ii = 0; ii < nlocal; ++ii) { first[ii] = tot; cnt[ii] = numneigh[ilist[ii]]; tot += cnt[ii]; }
    first[nlocal] = tot;
    CK(h->h_stage.ensure(sizeof(int) * std::max<long long>(tot, 1)));
    int* stage = h->h_stage.as<int>();
    CK(h->d_ilist.ensure(sizeof(int) * nlocal));
    CK(h->d_cand.ensure(sizeof(int) * std::max<long long>(tot, 1)));
    CK(h->d_first.ensure(sizeof(long long) * (nlocal + 1)));
    CK(h->d_numneigh.ensure(sizeof(int) * nlocal));
    CK(cudaMemcpyAsync(h->d_ilist.p, ilist, sizeof(int) * nlocal, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_first.p, first, sizeof(long long) * (nlocal + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_numneigh.p, cnt, sizeof(int) * nlocal, cudaMemcpyHostToDevice, st));
    // flatten the paged list into pinned memory in slabs of ~16 MB; the upload of slab s overlaps the flattening of s+1.
    // Rows that are contiguous in the caller's memory (LAMMPS stores the rows of one neighbour page back to back) are
    // copied as one run, and every slab is split evenly over the host threads by BYTES (a row is only ~200 bytes).
    const long long SLAB = 4ll << 20;                  // ints per slab
    int s0 = 0;
    while (s0 < nlocal) {
      int s1 = s0;
      while (s1 < nlocal && first[s1 + 1] - first[s0] <= SLAB) ++s1;
      if (s1 == s0) s1 = s0 + 1;
      const long long nb = first[s1] - first[s0];
      const int nrows = s1 - s0;
#pragma omp parallel
      {
#ifdef _OPENMP
        const int nth = omp_get_num_threads(), th = omp_get_thread_num();
#else
        const int nth = 1, th = 0;
#endif
        // thread th copies the rows whose first element lies in its byte range of the slab
        const long long lo = first[s0] + nb * th / nth, hi = first[s0] + nb * (th + 1) / nth;
        int r0 = (int)(std::lower_bound(first + s0, first + s1, lo) - first);
        const int r1 = (int)(std::lower_bound(first + s0, first + s1, hi) - first);
        while (r0 < r1) {                              // coalesce rows that are contiguous in the caller's memory
          int r2 = r0 + 1;
          const int* src = firstneigh[ilist[r0]];
          while (r2 < r1 && firstneigh[ilist[r2]] == src + (first[r2] - first[r0])) ++r2;
          memcpy(stage + first[r0], src, sizeof(int) * (size_t)(first[r2] - first[r0]));
          r0 = r2;
        }
      }
      (void)nrows;
      if (nb > 0) CK(cudaMemcpyAsync(h->d_cand.as<int>() + first[s0], stage + first[s0], sizeof(int) * nb, cudaMemcpyHostToDevice, st));
      s0 = s1;
    }
    h->list_nlocal = nlocal; h->list_ntot = ntot; h->list_tot = tot;
    h->host_ms[0] = ms_since(t_list);
  }
  h->list_reused = reuse ? 1 : 0;
  const bool want_eatom = eflag_atom && eatom;
  if (want_eatom) { CK(h->d_eatom_out.ensure(sizeof(double) * ntot)); CK(h->h_eatom.ensure(sizeof(double) * ntot)); }
  double eng_l = 0.0, vir_l[6] = {0, 0, 0, 0, 0, 0};
  StepIO io{};
  io.nlocal = nlocal; io.nghost = nghost; io.d_x = h->d_x.as<double>(); io.d_type = h->d_type.as<int>(); io.d_ilist = h->d_ilist.as<int>();
  io.acc = NeighAcc{h->d_cand.as<int>(), h->d_first.as<long long>(), h->d_numneigh.as<int>(), 0, 0, 1};
  io.eflag_atom = want_eatom ? 1 : 0; io.vflag_global = 1;
  io.d_f_inout = h->d_f_stage.as<double>(); io.d_eatom = want_eatom ? h->d_eatom_out.as<double>() : nullptr;
  io.cap_hint = (long)std::max<long long>(h->list_tot, 1);      // the edge list is a subset of the candidates
  io.wait_f = h->ev_copy; io.h_f = f; io.h_eatom_stage = want_eatom ? h->h_eatom.as<double>() : nullptr;
  int rc = run_step(h, io, &eng_l, vir_l);
  if (rc != ALG_OK) return rc;
  if (want_eatom) {                                  // eatom for locals (cpp:378)
    const double* eo = h->h_eatom.as<double>();
    for (int ii = 0; ii < nlocal; ++ii) eatom[ilist[ii]] = eo[ilist[ii]];
  }
  if (eng) *eng = eng_l;
  if (vflag_global && virial6) for (int q = 0; q < 6; ++q) virial6[q] = vir_l[q];
  auto& vo = h->outputs["virial"];
  vo = {vir_l[0], vir_l[3], vir_l[4], vir_l[3], vir_l[1], vir_l[5], vir_l[4], vir_l[5], vir_l[2]};
  h->host_ms[1] = ms_since(t_call);
  return ALG_OK;
}

extern "C" int alg_compute_device(alg_handle* h, int nlocal, int nghost, const double* d_x, const int* d_type, const int* d_ilist,
                                  const int* d_numneigh, const int* d_neighbors, int64_t stride_i, int64_t stride_jj,
                                  int eflag_atom, int vflag_global, double* d_f, double* d_eatom, double* eng, double* virial6,
                                  void* stream) {
  if (!h) return ALG_EINVAL;
  if (!h->have_map) return fail(h, ALG_ESTATE, "alg_compute_device called before alg_set_type_map");
  if (eng) *eng = 0.0;
  if (nlocal == 0) return ALG_OK;
  if (!d_x || !d_type || !d_ilist || !d_numneigh || !d_neighbors || !d_f) return fail(h, ALG_EINVAL, "alg_compute_device: null pointer");
  CK(cudaSetDevice(h->device));
  // order our stream after the caller's stream and vice versa
  cudaStream_t cs = reinterpret_cast<cudaStream_t>(stream);
  CK(cudaEventRecord(h->ev_order, cs));
  CK(cudaStreamWaitEvent(h->stream, h->ev_order, 0));
  StepIO io{};
  io.nlocal = nlocal; io.nghost = nghost; io.d_x = d_x; io.d_type = d_type; io.d_ilist = d_ilist;
  io.acc = NeighAcc{d_neighbors, nullptr, d_numneigh, (long long)stride_i, (long long)stride_jj, 0};
  io.eflag_atom = eflag_atom && d_eatom ? 1 : 0; io.vflag_global = vflag_global;
  io.d_f_inout = d_f; io.d_eatom = d_eatom;
  io.cap_hint = h->max_neighbors > 0 ? (long)nlocal * h->max_neighbors : 0;
  int rc = run_step(h, io, eng, virial6);
  cudaEventRecord(h->ev_order, h->stream);
  cudaStreamWaitEvent(cs, h->ev_order, 0);
  return rc;
}


extern "C" int alg_get_edges(alg_handle* h, const int64_t** edge_index, int64_t* nedges) {
  if (!h || !edge_index || !nedges) return ALG_EINVAL;
  if (!h->keep_edges) return fail(h, ALG_ESTATE, "alg_get_edges needs option keep_edges=1 before the compute");
  CK(cudaSetDevice(h->device));
  { int rc = resolve_pending(h); if (rc != ALG_OK) return rc; }
  const long E = h->last_E;
  h->edges_host.resize((size_t)2 * E);
  if (E > 0) {
    CK(cudaMemcpyAsync(h->edges_host.data(), h->d_edge_index.p, sizeof(int64_t) * 2 * E, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  *edge_index = h->edges_host.data();
  *nedges = E;
  return ALG_OK;
}

extern "C" int alg_get_output(alg_handle* h, const char* name, const double** ptr, int64_t* n) {
  if (!h || !name || !ptr || !n) return ALG_EINVAL;
  CK(cudaSetDevice(h->device));
  { int rc = resolve_pending(h); if (rc != ALG_OK) return rc; }
  const std::string key(name);
  auto it = h->outputs.find(key);
  if (it == h->outputs.end()) {
    cudaStream_t st = h->stream;
    const long E = h->last_E;
    const int ntot = h->last_ntot;
    const PipelineInfo& pi = h->pinfo;
    auto fetchf = [&](const void* dptr, size_t count, std::vector<float>& out) -> cudaError_t {
      out.resize(count);
      cudaError_t e = cudaMemcpyAsync(out.data(), dptr, sizeof(float) * count, cudaMemcpyDeviceToHost, st);
      if (e != cudaSuccess) return e;
      return cudaStreamSynchronize(st);
    };
    std::vector<float> raw;
    std::vector<double> out;
    if (key == "forces" || key == "atomic_energy") {
      const bool fo = key == "forces";
      out.resize(fo ? (size_t)3 * ntot : (size_t)ntot);
      CK(cudaMemcpyAsync(out.data(), fo ? h->d_forces.p : h->d_eall.p, sizeof(double) * out.size(), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
    } else if (!h->debug) {
      return fail(h, ALG_ENOTFOUND, "output '" + key + "' not available (intermediates need option debug=1)");
    } else if (key == "edge_energy") {
      CK(fetchf(h->d_edge_energy.p, E, raw)); out.assign(raw.begin(), raw.end());
    } else if (key == "edge_grad") {
      CK(fetchf(h->d_edge_grad.p, 3 * E, raw)); out.assign(raw.begin(), raw.end());
    } else if (key == "edge_vec") {
      CK(fetchf(h->d_rvec.p, 4 * E, raw));
      out.resize((size_t)3 * E);
      for (long e = 0; e < E; ++e) for (int q = 0; q < 3; ++q) out[3 * e + q] = raw[4 * e + q];
    } else if (h->dbg_ntiles == 0) {
      return fail(h, ALG_ENOTFOUND, "intermediate outputs need a single-chunk run of the chunked pipeline (option pipeline=tiled, raise chunk_edges)");
    } else if (key.size() == 2 && key[0] == 'x' && key[1] >= '0' && key[1] < '0' + h->nl) {
      CK(fetchf(h->c_X[key[1] - '0'].p, (size_t)h->dbg_ntiles * S * pi.TM, raw));
      detile(raw, h->dbg_ntiles, S, pi.TM, E, out, h->use_tc);
    } else if (key == "dx0") {
      CK(fetchf(h->c_dX.p, (size_t)h->dbg_ntiles * S * pi.TM, raw));
      detile(raw, h->dbg_ntiles, S, pi.TM, E, out, h->use_tc);
    } else if (key == "du") {
      CK(fetchf(h->c_du.p, (size_t)h->dbg_ntiles * pi.TM, raw));
      detile(raw, h->dbg_ntiles, 1, pi.TM, E, out);
    } else if (key.size() == 2 && key[0] == 'V' && key[1] >= '1' && key[1] < '0' + h->nl) {
      const int k = key[1] - '0';
      const int vd = pi.vdim[k], vs = pi.vstride[k], TMv = pi.TM;
      CK(fetchf(h->c_V[k].p, (size_t)h->dbg_ntiles * U * vs * TMv, raw));
      out.assign((size_t)E * U * vd, 0.0);                              // [E][u][comp]
      const bool comp4 = h->use_tc && vs % 4 == 0;                      // tensor-core pipeline: comp4 packing inside a (padded) channel block
      for (long e = 0; e < E; ++e) {
        const long t = e / TMv, m = e % TMv;
        for (int u = 0; u < U; ++u)
          for (int cc = 0; cc < vd; ++cc)
            out[((size_t)e * U + u) * vd + cc] = comp4 ? raw[((size_t)t * U + u) * vs * TMv + ((size_t)(cc >> 2) * TMv + m) * 4 + (cc & 3)]
                                                       : raw[(((size_t)t * U + u) * vs + cc) * TMv + m];
      }
    } else if ((key.rfind("gamma", 0) == 0 && key.size() == 6) || (key.rfind("dgamma", 0) == 0 && key.size() == 7)) {
      const bool d = key[0] == 'd';
      const int k = key.back() - '0';
      if (k < 0 || k >= h->nl) return fail(h, ALG_ENOTFOUND, "no such layer");
      CK(fetchf(d ? h->c_dgamma[k].p : h->c_gamma[k].p, (size_t)h->dbg_ncent * pi.F, raw));
      out.assign(raw.begin(), raw.end());                              // [centre][lm][u]
    } else {
      return fail(h, ALG_ENOTFOUND, "unknown output '" + key + "'");
    }
    it = h->outputs.emplace(key, std::move(out)).first;
  }
  *ptr = it->second.data();
  *n = (int64_t)it->second.size();
  return ALG_OK;
}

extern "C" int alg_get_timings(alg_handle* h, double* ms3) {
  if (!h || !ms3) return ALG_EINVAL;
  for (int q = 0; q < 3; ++q) ms3[q] = h->timings[q];
  return ALG_OK;
}

extern "C" int alg_get_stats(alg_handle* h, const char* what, double* out, int n) {
  if (!h || !what || !out) return ALG_EINVAL;
  const std::string k(what);
  const double* src = nullptr; int m = 0;
  { int rc = resolve_pending(h); if (rc != ALG_OK) return rc; }
  if (k == "kernel_ms") { src = h->kernel_ms; m = KID_COUNT; }
  else if (k == "kernel_launches") { src = h->kernel_n; m = KID_COUNT; }
  else if (k == "step") { src = h->step_stats; m = 4; }
  else if (k == "host_ms") { src = h->host_ms; m = 3; }
  else if (k == "list_reused") { if (n > 0) out[0] = h->list_reused; return ALG_OK; }
  else if (k == "pipeline") {                       // [1 if the last step ran the fused kernel, CTAs of the fused grid, sticky tiled fallback]
    const double v[4] = {h->last_fused ? 1.0 : 0.0, (double)h->fused_grid, h->force_tiled ? 1.0 : 0.0, h->last_mode == 2 ? 1.0 : 0.0};
    for (int i = 0; i < n && i < 4; ++i) out[i] = v[i];
    return ALG_OK;
  }
  else return fail(h, ALG_ENOTFOUND, "unknown stats group " + k);
  for (int i = 0; i < n && i < m; ++i) out[i] = src[i];
  return ALG_OK;
}

extern "C" int alg_halo_pack(const double* d_x, const int* d_list, int n, const double* d_shift, double* d_buf, void* stream) {
  if (n <= 0) return ALG_OK;
  if (!d_x || !d_list || !d_buf) return ALG_EINVAL;
  k_halo_pack<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_x, d_list, n, d_shift, d_buf);
  return cudaGetLastError() == cudaSuccess ? ALG_OK : ALG_ECUDA;
}
extern "C" int alg_halo_unpack_add(double* d_f, const int* d_list, int n, const double* d_buf, void* stream) {
  if (n <= 0) return ALG_OK;
  if (!d_f || !d_list || !d_buf) return ALG_EINVAL;
  k_halo_unpack_add<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_f, d_list, n, d_buf);
  return cudaGetLastError() == cudaSuccess ? ALG_OK : ALG_ECUDA;
}
