// C driver around the device-resident pair style (pair_style allegro/kk) compiled against lmpshim/kokkos_shim.h: plays
// LAMMPS + the KOKKOS package for src/pair_allegro_b200_kokkos.cpp -- atoms and the 2-D neighbour view live in device
// memory, the pair style only ever sees device views.  C API for ctypes (lmpshim/driver.py: ShimLammpsKK).
#include <cstring>
#include <string>
#include <vector>

#include "kokkos_shim.h"
#include "pair_allegro_b200_kokkos.h"

using namespace LAMMPS_NS;

namespace {
struct ShimKK {
  LAMMPS lmp;
  KokkosLMP kk;
  AtomKokkos* atomKK = nullptr;
  NeighListKokkos<LMPDeviceType> list;
  Pair* pair = nullptr;
  std::vector<double> x, f;
  std::vector<double*> xrow, frow;
  std::vector<int> type, ilist, numneigh;
  std::vector<tagint> tag;
  int* d_ilist = nullptr; int* d_numneigh = nullptr; int* d_neighbors = nullptr;
  std::string err;
  ~ShimKK() { cudaFree(d_ilist); cudaFree(d_numneigh); cudaFree(d_neighbors); }
};
template <class F> int guard(ShimKK* s, F&& fn) {
  try { fn(); return 0; }
  catch (const std::exception& e) { s->err = e.what(); return -1; }
  catch (...) { s->err = "unknown C++ exception"; return -1; }
}
}  // namespace

#define API extern "C" __attribute__((visibility("default")))

API void* shimkk_create(int ntypes, int nlocal, int nghost, const double* x, const int* type, const long long* tag, int neighflag) {
  ShimKK* s = new ShimKK();
  delete s->lmp.atom;
  s->atomKK = new AtomKokkos();
  s->lmp.atom = s->atomKK;
  s->lmp.comm->atom = s->atomKK;
  s->kk.neighflag = neighflag;
  s->lmp.kokkos = &s->kk;
  const int n = nlocal + nghost;
  s->x.assign(x, x + 3 * (size_t)n);
  s->f.assign(3 * (size_t)n, 0.0);
  s->type.assign(type, type + n);
  s->tag.assign(tag, tag + n);
  s->xrow.resize(n > 0 ? n : 1); s->frow.resize(n > 0 ? n : 1);
  for (int i = 0; i < n; ++i) { s->xrow[i] = s->x.data() + 3 * (size_t)i; s->frow[i] = s->f.data() + 3 * (size_t)i; }
  Atom* a = s->lmp.atom;
  a->ntypes = ntypes; a->nlocal = nlocal; a->nghost = nghost; a->nmax = n;
  a->x = s->xrow.data(); a->f = s->frow.data(); a->type = s->type.data(); a->tag = s->tag.data();
  return s;
}
API void shimkk_destroy(void* p) { ShimKK* s = (ShimKK*)p; if (!s) return; delete s->pair; delete s; }
API const char* shimkk_last_error(void* p) { return ((ShimKK*)p)->err.c_str(); }

// full neighbour list of the local atoms as the KOKKOS 2-D view d_neighbors(i, jj): layout_left != 0 -> data[i + jj*nrows]
// (the device default), else data[i*maxneighs + jj]
API void shimkk_set_list(void* p, int inum, int gnum, const int* ilist, const int* numneigh, const int* neigh_flat, const long long* first, int layout_left) {
  ShimKK* s = (ShimKK*)p;
  const int n = inum + gnum;
  s->ilist.assign(ilist, ilist + n);
  s->numneigh.assign(numneigh, numneigh + n);
  int maxn = 1;
  for (int i = 0; i < inum; ++i) maxn = std::max(maxn, numneigh[ilist[i]]);
  const size_t rows = (size_t)n;
  std::vector<int> nb(rows * maxn, 0);
  for (int ii = 0; ii < inum; ++ii) {
    const int i = ilist[ii];
    for (int jj = 0; jj < numneigh[i]; ++jj) nb[layout_left ? (size_t)i + (size_t)jj * rows : (size_t)i * maxn + jj] = neigh_flat[first[i] + jj];
  }
  cudaFree(s->d_ilist); cudaFree(s->d_numneigh); cudaFree(s->d_neighbors);
  cudaMalloc(&s->d_ilist, sizeof(int) * std::max(n, 1)); cudaMalloc(&s->d_numneigh, sizeof(int) * std::max(n, 1)); cudaMalloc(&s->d_neighbors, sizeof(int) * nb.size());
  cudaMemcpy(s->d_ilist, ilist, sizeof(int) * n, cudaMemcpyHostToDevice);
  cudaMemcpy(s->d_numneigh, numneigh, sizeof(int) * n, cudaMemcpyHostToDevice);
  cudaMemcpy(s->d_neighbors, nb.data(), sizeof(int) * nb.size(), cudaMemcpyHostToDevice);
  s->list.inum = inum; s->list.gnum = gnum;
  s->list.ilist = s->ilist.data(); s->list.numneigh = s->numneigh.data();
  s->list.d_ilist = {s->d_ilist, (size_t)n};
  s->list.d_numneigh = {s->d_numneigh, (size_t)n};
  s->list.d_neighbors = layout_left ? Kokkos::View2<const int>{s->d_neighbors, rows, (size_t)maxn, 1, rows}
                                    : Kokkos::View2<const int>{s->d_neighbors, rows, (size_t)maxn, (size_t)maxn, 1};
  if (s->pair) s->pair->init_list(0, &s->list);
}
API int shimkk_set_newton(void* p, int newton_pair) { ((ShimKK*)p)->lmp.force->newton_pair = newton_pair; return 0; }
API int shimkk_pair_create(void* p) {
  ShimKK* s = (ShimKK*)p;
  return guard(s, [&] { s->pair = new PairAllegroB200Kokkos(&s->lmp); s->lmp.force->pair = s->pair; s->pair->init_list(0, &s->list); });
}
API int shimkk_pair_settings(void* p, int narg, char** arg) { ShimKK* s = (ShimKK*)p; return guard(s, [&] { s->pair->settings(narg, arg); }); }
API int shimkk_pair_coeff(void* p, int narg, char** arg) { ShimKK* s = (ShimKK*)p; return guard(s, [&] { s->pair->coeff(narg, arg); }); }
API int shimkk_pair_init_style(void* p) { ShimKK* s = (ShimKK*)p; return guard(s, [&] { s->pair->init_style(); }); }
API int shimkk_pair_compute(void* p, int eflag, int vflag, int zero_forces) {
  ShimKK* s = (ShimKK*)p;
  return guard(s, [&] {
    if (zero_forces) { std::fill(s->f.begin(), s->f.end(), 0.0); s->atomKK->modified(Host, F_MASK); }
    s->pair->compute(eflag, vflag);
  });
}
// what LAMMPS does before a host-side consumer reads f: atomKK->sync(Host, F_MASK)
API void shimkk_get_forces(void* p, double* out) {
  ShimKK* s = (ShimKK*)p;
  s->atomKK->sync(Host, F_MASK);
  memcpy(out, s->f.data(), sizeof(double) * s->f.size());
}
API double shimkk_get_eng(void* p) { return ((ShimKK*)p)->pair->eng_vdwl; }
API void shimkk_get_virial(void* p, double* out6) { memcpy(out6, ((ShimKK*)p)->pair->virial, sizeof(double) * 6); }
API int shimkk_get_eatom(void* p, double* out) {
  ShimKK* s = (ShimKK*)p;
  if (!s->pair->eatom) return -1;
  memcpy(out, s->pair->eatom, sizeof(double) * (s->lmp.atom->nlocal + s->lmp.atom->nghost));
  return 0;
}
