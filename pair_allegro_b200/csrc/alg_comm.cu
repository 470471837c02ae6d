// Ghost halo exchange of the spatial-domain multi-GPU run, as product code behind the C-ABI (include/allegro_b200.h,
// alg_comm_*): what LAMMPS' Comm does for the reference around PairNequIPAllegro::compute -- comm->forward_comm() of the
// ghost positions before the pair style and comm->reverse_comm() of the ghost forces after it, required by `newton on`
// (/root/reference/pair_nequip_allegro.cpp:149; owner-accumulation pattern of
// /root/reference/compute/compute_allegro.cpp:159-189).  One rank per GPU; transport = NCCL point-to-point over NVLink
// (grouped ncclSend / ncclRecv to every neighbouring domain at once -- NVSwitch is uniform, so there are no staged
// 6-direction brick swaps), pack / unpack = the kernels below.  Images a rank owns itself (periodic self-images) never
// leave the device.  The reverse unpack is a sorted segmented sum: every owner atom adds its image contributions in a
// fixed order (peer order, then send order), so forces are bit-reproducible run to run -- no floating-point atomics.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy PyTorch already loaded, or the system one), so the
// single-GPU library has no link-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/allegro_b200.h"

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
  bool load() {
    if (lib) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { err = std::string("cannot load NCCL (libnccl.so.2): ") + dlerror(); return false; }
#define ALG_SYM(field, sym) \
    field = reinterpret_cast<decltype(field)>(dlsym(lib, sym)); \
    if (!field) { err = std::string("NCCL symbol missing: ") + sym; lib = nullptr; return false; }
    ALG_SYM(GetUniqueId, "ncclGetUniqueId");
    ALG_SYM(CommInitRank, "ncclCommInitRank");
    ALG_SYM(CommDestroy, "ncclCommDestroy");
    ALG_SYM(Send, "ncclSend");
    ALG_SYM(Recv, "ncclRecv");
    ALG_SYM(GroupStart, "ncclGroupStart");
    ALG_SYM(GroupEnd, "ncclGroupEnd");
    ALG_SYM(AllReduce, "ncclAllReduce");
    ALG_SYM(GetErrorString, "ncclGetErrorString");
#undef ALG_SYM
    return true;
  }
};
NcclApi g_nccl;
std::string g_comm_error;

template <class T> struct DBuf {
  T* p = nullptr; size_t n = 0;
  cudaError_t ensure(size_t count) {
    if (count <= n) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; n = 0;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

// buf[e] = x[idx[e]] (+ shift[e]): ghost positions as the receiving domain sees them (periodic image shift applied by the owner)
__global__ void k_comm_pack(const double* __restrict__ x, const int* __restrict__ idx, const double* __restrict__ shift, long n, double* __restrict__ buf) {
  const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (e >= n) return;
  const size_t i = (size_t)idx[e];
  const double sx = shift ? shift[3 * e + 0] : 0.0, sy = shift ? shift[3 * e + 1] : 0.0, sz = shift ? shift[3 * e + 2] : 0.0;
  buf[3 * e + 0] = x[3 * i + 0] + sx;
  buf[3 * e + 1] = x[3 * i + 1] + sy;
  buf[3 * e + 2] = x[3 * i + 2] + sz;
}
// owner accumulation: atom u_atom[k] adds the forces of all its images, rbuf[u_slot[u_ptr[k] .. u_ptr[k+1])], in that fixed order
__global__ void k_comm_unpack(double* __restrict__ f, const int* __restrict__ u_atom, const int* __restrict__ u_ptr, const int* __restrict__ u_slot,
                              int nu, const double* __restrict__ rbuf) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nu) return;
  double ax = 0.0, ay = 0.0, az = 0.0;
  for (int q = u_ptr[k]; q < u_ptr[k + 1]; ++q) {
    const size_t s = (size_t)u_slot[q];
    ax += rbuf[3 * s + 0]; ay += rbuf[3 * s + 1]; az += rbuf[3 * s + 2];
  }
  const size_t i = (size_t)u_atom[k];
  f[3 * i + 0] += ax; f[3 * i + 1] += ay; f[3 * i + 2] += az;
}

}  // namespace

struct alg_comm {
  int device = 0, nranks = 1, rank = 0;
  ncclComm_t comm = nullptr;
  std::string err;
  // plan
  int npeer = 0;
  std::vector<int> peer, send_cnt, send_off, recv_begin, recv_cnt;
  long nsend = 0;
  bool have_shift = false;
  DBuf<int> d_idx, d_uatom, d_uptr, d_uslot;
  DBuf<double> d_shift, d_sbuf, d_rbuf, d_red;
  int nu = 0;
  long bytes_fwd = 0;       // bytes this rank sends to OTHER ranks per forward exchange (the reverse moves the ghost slices back)
  long bytes_rev = 0;
};

#define CCK(call)                                                                                               \
  do {                                                                                                          \
    cudaError_t _e = (call);                                                                                    \
    if (_e != cudaSuccess) { c->err = std::string("CUDA error: ") + cudaGetErrorString(_e) + " (" #call ")"; return ALG_ECUDA; } \
  } while (0)
#define NCK(call)                                                                                               \
  do {                                                                                                          \
    ncclResult_t _r = (call);                                                                                   \
    if (_r != ncclSuccess) { c->err = std::string("NCCL error: ") + g_nccl.GetErrorString(_r) + " (" #call ")"; return ALG_ECUDA; } \
  } while (0)

extern "C" const char* alg_comm_last_error(const alg_comm* c) { return c ? c->err.c_str() : g_comm_error.c_str(); }

extern "C" int alg_comm_unique_id(char* id128) {
  if (!id128) return ALG_EINVAL;
  if (!g_nccl.load()) { g_comm_error = g_nccl.err; return ALG_ECUDA; }
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) { g_comm_error = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return ALG_ECUDA; }
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, 128);
  return ALG_OK;
}

extern "C" int alg_comm_create(int cuda_device, int nranks, int rank, const char* id128, alg_comm** out) {
  if (!out || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id128)) { g_comm_error = "alg_comm_create: bad arguments"; return ALG_EINVAL; }
  *out = nullptr;
  alg_comm* c = new alg_comm();
  c->device = cuda_device; c->nranks = nranks; c->rank = rank;
  cudaError_t e = cudaSetDevice(cuda_device);
  if (e != cudaSuccess) { g_comm_error = std::string("CUDA error: ") + cudaGetErrorString(e); delete c; return ALG_ECUDA; }
  if (nranks > 1) {                                  // a single rank only has periodic self-images: no communicator needed
    if (!g_nccl.load()) { g_comm_error = g_nccl.err; delete c; return ALG_ECUDA; }
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) { g_comm_error = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r); delete c; return ALG_ECUDA; }
  }
  *out = c;
  return ALG_OK;
}

extern "C" void alg_comm_destroy(alg_comm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  if (c->comm) g_nccl.CommDestroy(c->comm);
  c->d_idx.release(); c->d_uatom.release(); c->d_uptr.release(); c->d_uslot.release();
  c->d_shift.release(); c->d_sbuf.release(); c->d_rbuf.release(); c->d_red.release();
  delete c;
}

extern "C" int alg_comm_set_plan(alg_comm* c, int npeer, const int* peer_rank, const int* send_count, const int* const* send_index,
                                 const double* const* send_shift, const int* recv_begin, const int* recv_count) {
  if (!c) return ALG_EINVAL;
  if (npeer < 0 || (npeer > 0 && (!peer_rank || !send_count || !send_index || !recv_begin || !recv_count))) { c->err = "alg_comm_set_plan: bad arguments"; return ALG_EINVAL; }
  CCK(cudaSetDevice(c->device));
  c->npeer = npeer;
  c->peer.assign(peer_rank, peer_rank + npeer);
  c->send_cnt.assign(send_count, send_count + npeer);
  c->recv_begin.assign(recv_begin, recv_begin + npeer);
  c->recv_cnt.assign(recv_count, recv_count + npeer);
  c->send_off.assign(npeer + 1, 0);
  for (int p = 0; p < npeer; ++p) {
    if (c->peer[p] < 0 || c->peer[p] >= c->nranks || send_count[p] < 0 || recv_count[p] < 0 || (send_count[p] > 0 && !send_index[p])) {
      c->err = "alg_comm_set_plan: bad peer entry"; return ALG_EINVAL;
    }
    if (c->peer[p] == c->rank && send_count[p] != recv_count[p]) { c->err = "alg_comm_set_plan: self-image send and receive counts differ"; return ALG_EINVAL; }
    c->send_off[p + 1] = c->send_off[p] + send_count[p];
  }
  const long n = c->send_off[npeer];
  c->nsend = n;
  std::vector<int> idx((size_t)std::max<long>(n, 1));
  std::vector<double> shift((size_t)std::max<long>(3 * n, 1), 0.0);
  c->have_shift = false;
  c->bytes_fwd = 0; c->bytes_rev = 0;
  for (int p = 0; p < npeer; ++p) {
    if (send_count[p] > 0) memcpy(idx.data() + c->send_off[p], send_index[p], sizeof(int) * send_count[p]);
    if (send_shift && send_shift[p] && send_count[p] > 0) {
      memcpy(shift.data() + 3 * (size_t)c->send_off[p], send_shift[p], sizeof(double) * 3 * send_count[p]);
      c->have_shift = true;
    }
    if (c->peer[p] != c->rank) { c->bytes_fwd += 24L * send_count[p]; c->bytes_rev += 24L * recv_count[p]; }
  }
  // reverse unpack: entries sorted by (owner atom, slot) -> unique atoms + CSR of slots (fixed summation order)
  std::vector<int> order((size_t)n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return idx[a] < idx[b]; });
  std::vector<int> uatom, uptr, uslot((size_t)std::max<long>(n, 1));
  for (long q = 0; q < n; ++q) {
    const int e = order[q];
    if (q == 0 || idx[e] != idx[order[q - 1]]) { uatom.push_back(idx[e]); uptr.push_back((int)q); }
    uslot[q] = e;
  }
  uptr.push_back((int)n);
  c->nu = (int)uatom.size();
  if (uatom.empty()) uatom.push_back(0);
  CCK(c->d_idx.ensure(idx.size())); CCK(c->d_shift.ensure(shift.size())); CCK(c->d_sbuf.ensure(shift.size())); CCK(c->d_rbuf.ensure(shift.size()));
  CCK(c->d_uatom.ensure(uatom.size())); CCK(c->d_uptr.ensure(uptr.size())); CCK(c->d_uslot.ensure(uslot.size())); CCK(c->d_red.ensure(64));
  CCK(cudaMemcpy(c->d_idx.p, idx.data(), sizeof(int) * idx.size(), cudaMemcpyHostToDevice));
  CCK(cudaMemcpy(c->d_shift.p, shift.data(), sizeof(double) * shift.size(), cudaMemcpyHostToDevice));
  CCK(cudaMemcpy(c->d_uatom.p, uatom.data(), sizeof(int) * uatom.size(), cudaMemcpyHostToDevice));
  CCK(cudaMemcpy(c->d_uptr.p, uptr.data(), sizeof(int) * uptr.size(), cudaMemcpyHostToDevice));
  CCK(cudaMemcpy(c->d_uslot.p, uslot.data(), sizeof(int) * uslot.size(), cudaMemcpyHostToDevice));
  return ALG_OK;
}

// ghost x <- owner x (+ image shift).  Stream-ordered; no host synchronisation.
extern "C" int alg_comm_forward(alg_comm* c, double* d_x, void* stream) {
  if (!c || !d_x) return ALG_EINVAL;
  CCK(cudaSetDevice(c->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long n = c->nsend;
  if (n > 0) {
    k_comm_pack<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_x, c->d_idx.p, c->have_shift ? c->d_shift.p : nullptr, n, c->d_sbuf.p);
    CCK(cudaGetLastError());
  }
  bool remote = false;
  for (int p = 0; p < c->npeer; ++p) remote = remote || c->peer[p] != c->rank;
  if (remote) {
    NCK(g_nccl.GroupStart());
    for (int p = 0; p < c->npeer; ++p) {
      if (c->peer[p] == c->rank) continue;
      if (c->send_cnt[p] > 0) NCK(g_nccl.Send(c->d_sbuf.p + 3 * (size_t)c->send_off[p], 3 * (size_t)c->send_cnt[p], ncclDouble, c->peer[p], c->comm, st));
      if (c->recv_cnt[p] > 0) NCK(g_nccl.Recv(d_x + 3 * (size_t)c->recv_begin[p], 3 * (size_t)c->recv_cnt[p], ncclDouble, c->peer[p], c->comm, st));
    }
    NCK(g_nccl.GroupEnd());
  }
  for (int p = 0; p < c->npeer; ++p)
    if (c->peer[p] == c->rank && c->send_cnt[p] > 0)
      CCK(cudaMemcpyAsync(d_x + 3 * (size_t)c->recv_begin[p], c->d_sbuf.p + 3 * (size_t)c->send_off[p], sizeof(double) * 3 * c->send_cnt[p], cudaMemcpyDeviceToDevice, st));
  return ALG_OK;
}

// owner f += ghost f (newton on).  Stream-ordered; deterministic summation order.
extern "C" int alg_comm_reverse(alg_comm* c, double* d_f, void* stream) {
  if (!c || !d_f) return ALG_EINVAL;
  CCK(cudaSetDevice(c->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  bool remote = false;
  for (int p = 0; p < c->npeer; ++p) remote = remote || c->peer[p] != c->rank;
  if (remote) {
    NCK(g_nccl.GroupStart());
    for (int p = 0; p < c->npeer; ++p) {
      if (c->peer[p] == c->rank) continue;
      if (c->recv_cnt[p] > 0) NCK(g_nccl.Send(d_f + 3 * (size_t)c->recv_begin[p], 3 * (size_t)c->recv_cnt[p], ncclDouble, c->peer[p], c->comm, st));
      if (c->send_cnt[p] > 0) NCK(g_nccl.Recv(c->d_rbuf.p + 3 * (size_t)c->send_off[p], 3 * (size_t)c->send_cnt[p], ncclDouble, c->peer[p], c->comm, st));
    }
    NCK(g_nccl.GroupEnd());
  }
  for (int p = 0; p < c->npeer; ++p)
    if (c->peer[p] == c->rank && c->send_cnt[p] > 0)
      CCK(cudaMemcpyAsync(c->d_rbuf.p + 3 * (size_t)c->send_off[p], d_f + 3 * (size_t)c->recv_begin[p], sizeof(double) * 3 * c->send_cnt[p], cudaMemcpyDeviceToDevice, st));
  if (c->nu > 0) {
    k_comm_unpack<<<(c->nu + 255) / 256, 256, 0, st>>>(d_f, c->d_uatom.p, c->d_uptr.p, c->d_uslot.p, c->nu, c->d_rbuf.p);
    CCK(cudaGetLastError());
  }
  return ALG_OK;
}

// sum of per-rank scalars (eng_vdwl, virial[6]: LAMMPS sums them over ranks with MPI_Allreduce); synchronises the stream
extern "C" int alg_comm_allreduce_sum(alg_comm* c, double* values, int n, void* stream) {
  if (!c || !values || n < 0 || n > 64) return ALG_EINVAL;
  if (c->nranks == 1 || n == 0) return ALG_OK;
  CCK(cudaSetDevice(c->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CCK(c->d_red.ensure(64));
  CCK(cudaMemcpyAsync(c->d_red.p, values, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  NCK(g_nccl.AllReduce(c->d_red.p, c->d_red.p, (size_t)n, ncclDouble, ncclSum, c->comm, st));
  CCK(cudaMemcpyAsync(values, c->d_red.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  CCK(cudaStreamSynchronize(st));
  return ALG_OK;
}

extern "C" int alg_comm_stats(const alg_comm* c, double* out4) {
  if (!c || !out4) return ALG_EINVAL;
  out4[0] = (double)c->bytes_fwd; out4[1] = (double)c->bytes_rev; out4[2] = (double)c->nsend; out4[3] = (double)c->nu;
  return ALG_OK;
}
