// Minimal stand-in for the parts of Kokkos and of LAMMPS' KOKKOS package that `pair_style allegro/kk` touches
// (/root/reference/pair_nequip_allegro_kokkos.h:16-113, .cpp:87-353): device views as (pointer, extents, strides),
// DualView host/device pairs with sync/modify, AtomKokkos, NeighListKokkos, MemoryKokkos, KokkosLMP.  Test harness
// standing in for LAMMPS core + Kokkos (neither is in the image); written from the public Kokkos / LAMMPS developer
// documentation, no code of either.  Device memory is plain CUDA runtime memory.
//
// Layouts follow LAMMPS on CUDA: x / f are `T*[3]` LayoutRight, the 2-D neighbour view is the device default
// LayoutLeft (d_neighbors(i,jj) at data[i + jj*extent(0)]); the harness can also hand out a LayoutRight view.
#pragma once
#include <cuda_runtime.h>

#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "lammps.h"
#include "pair.h"

namespace Kokkos {
struct LayoutLeft {};
struct LayoutRight {};
struct Cuda {};
struct Serial {};
template <class T> struct View1 {
  T* p = nullptr; size_t n = 0;
  T* data() const { return p; }
  size_t extent(int) const { return n; }
};
template <class T> struct View2 {
  T* p = nullptr; size_t n0 = 0, n1 = 0, s0 = 0, s1 = 0;
  T* data() const { return p; }
  size_t extent(int d) const { return d == 0 ? n0 : n1; }
  size_t stride(int d) const { return d == 0 ? s0 : s1; }
};
}  // namespace Kokkos

namespace LAMMPS_NS {
typedef Kokkos::Cuda LMPDeviceType;
typedef Kokkos::Serial LMPHostType;
typedef double X_FLOAT;
typedef double F_FLOAT;
typedef double E_FLOAT;
enum ExecutionSpace { Host, Device };
template <class D> struct ExecutionSpaceFromDevice { static const ExecutionSpace space = std::is_same<D, LMPDeviceType>::value ? Device : Host; };
enum { X_MASK = 1, V_MASK = 2, F_MASK = 4, TAG_MASK = 8, TYPE_MASK = 16, ENERGY_MASK = 1 << 20, VIRIAL_MASK = 1 << 21 };
enum { FULL = 1, HALFTHREAD = 2, HALF = 4 };

// host array + device mirror; sync<>() copies towards the side that is stale
template <class T> struct DualView1 {
  T* h = nullptr; T* d = nullptr; size_t n = 0; int modified_on = 0;   // 1 = host newer, 2 = device newer
  template <class D> Kokkos::View1<T> view() const { return Kokkos::View1<T>{std::is_same<D, LMPDeviceType>::value ? d : h, n}; }
  template <class D> void modify() { modified_on = std::is_same<D, LMPDeviceType>::value ? 2 : 1; }
  template <class D> void sync() {
    const bool to_dev = std::is_same<D, LMPDeviceType>::value;
    if (to_dev && modified_on == 1) cudaMemcpy(d, h, sizeof(T) * n, cudaMemcpyHostToDevice);
    if (!to_dev && modified_on == 2) cudaMemcpy(h, d, sizeof(T) * n, cudaMemcpyDeviceToHost);
    modified_on = 0;
  }
};

template <class D> struct ArrayTypes {
  typedef Kokkos::View2<const X_FLOAT> t_x_array_randomread;
  typedef Kokkos::View2<F_FLOAT> t_f_array;
  typedef Kokkos::View1<const tagint> t_tagint_1d;
  typedef Kokkos::View1<const int> t_int_1d_randomread;
  typedef Kokkos::View2<const int> t_neighbors_2d;
  typedef Kokkos::View1<E_FLOAT> t_efloat_1d;
};
struct DAT { typedef DualView1<E_FLOAT> tdual_efloat_1d; };

class AtomKokkos : public Atom {
 public:
  // device mirrors of x [n][3], f [n][3], type, tag
  double* d_x = nullptr; double* d_f = nullptr; int* d_type = nullptr; tagint* d_tag = nullptr;
  size_t cap = 0;
  int host_modified = X_MASK | F_MASK | TYPE_MASK | TAG_MASK, device_modified = 0;
  ~AtomKokkos() { cudaFree(d_x); cudaFree(d_f); cudaFree(d_type); cudaFree(d_tag); }
  void ensure() {
    const size_t n = (size_t)nlocal + nghost;
    if (n <= cap) return;
    cudaFree(d_x); cudaFree(d_f); cudaFree(d_type); cudaFree(d_tag);
    cap = n;
    cudaMalloc(&d_x, sizeof(double) * 3 * cap); cudaMalloc(&d_f, sizeof(double) * 3 * cap);
    cudaMalloc(&d_type, sizeof(int) * cap); cudaMalloc(&d_tag, sizeof(tagint) * cap);
    host_modified = X_MASK | F_MASK | TYPE_MASK | TAG_MASK;
  }
  void sync(ExecutionSpace space, int mask) {
    ensure();
    const size_t n = (size_t)nlocal + nghost;
    if (n == 0) return;
    if (space == Device) {
      const int m = mask & host_modified;
      if (m & X_MASK) cudaMemcpy(d_x, x[0], sizeof(double) * 3 * n, cudaMemcpyHostToDevice);
      if (m & F_MASK) cudaMemcpy(d_f, f[0], sizeof(double) * 3 * n, cudaMemcpyHostToDevice);
      if (m & TYPE_MASK) cudaMemcpy(d_type, type, sizeof(int) * n, cudaMemcpyHostToDevice);
      if (m & TAG_MASK) cudaMemcpy(d_tag, tag, sizeof(tagint) * n, cudaMemcpyHostToDevice);
      host_modified &= ~m;
    } else {
      const int m = mask & device_modified;
      if (m & F_MASK) cudaMemcpy(f[0], d_f, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost);
      if (m & X_MASK) cudaMemcpy(x[0], d_x, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost);
      device_modified &= ~m;
    }
  }
  void modified(ExecutionSpace space, int mask) { if (space == Device) device_modified |= mask; else host_modified |= mask; }
  struct DualX { AtomKokkos* a; template <class D> Kokkos::View2<const X_FLOAT> view() const { const size_t n = (size_t)a->nlocal + a->nghost; return {a->d_x, n, 3, 3, 1}; } } k_x{this};
  struct DualF { AtomKokkos* a; template <class D> Kokkos::View2<F_FLOAT> view() const { const size_t n = (size_t)a->nlocal + a->nghost; return {a->d_f, n, 3, 3, 1}; } } k_f{this};
  struct DualT { AtomKokkos* a; template <class D> Kokkos::View1<const int> view() const { return {a->d_type, (size_t)a->nlocal + a->nghost}; } } k_type{this};
  struct DualG { AtomKokkos* a; template <class D> Kokkos::View1<const tagint> view() const { return {a->d_tag, (size_t)a->nlocal + a->nghost}; } } k_tag{this};
};

template <class D> class NeighListKokkos : public NeighList {
 public:
  Kokkos::View1<const int> d_ilist, d_numneigh;
  Kokkos::View2<const int> d_neighbors;
};

class MemoryKokkos {
 public:
  void create_kokkos(DualView1<E_FLOAT>& k, double*& host, int n, const char*) {
    destroy_kokkos(k, host);
    k.n = n > 0 ? n : 1;
    k.h = (double*)calloc(k.n, sizeof(double));
    cudaMalloc(&k.d, sizeof(double) * k.n);
    cudaMemset(k.d, 0, sizeof(double) * k.n);
    host = k.h;
  }
  void destroy_kokkos(DualView1<E_FLOAT>& k, double*& host) {
    if (k.d) cudaFree(k.d);
    if (k.h) free(k.h);
    k = DualView1<E_FLOAT>();
    host = nullptr;
  }
};

class KokkosLMP {
 public:
  int neighflag = HALF;       // `-pk kokkos neigh half` (the reference rejects FULL, pair_nequip_allegro_kokkos.cpp:402-405)
  int newtonflag = 1;
};
}  // namespace LAMMPS_NS
