// C driver around ONE pair style compiled against the lmpshim headers.  The pair class is
// selected at compile time:
//   -DSHIM_PAIR_HEADER='<pair_nequip_allegro.h>' -DSHIM_PAIR_CLASS='PairNequIPAllegro<false>'   (oracle/_ref)
//   -DSHIM_PAIR_HEADER='"pair_allegro_b200.h"'   -DSHIM_PAIR_CLASS='PairAllegroB200'            (this repo)
// so the reference's unmodified sources and this repo's pair style run under the same harness
// on the same inputs.  Exposes a C API for ctypes (tests/, bench.py).
#include <cstring>
#include <string>
#include <vector>

#include "lammps.h"
#include "pair.h"
#include SHIM_PAIR_HEADER
#ifdef SHIM_COMPUTE_HEADER
//   -DSHIM_COMPUTE_HEADER='<compute_allegro.h>'        -DSHIM_COMPUTE_TEMPLATE=ComputeAllegro       (oracle/_ref)
//   -DSHIM_COMPUTE_HEADER='"compute_allegro_b200.h"'   -DSHIM_COMPUTE_TEMPLATE=ComputeAllegroB200   (this repo)
#include "compute.h"
#include SHIM_COMPUTE_HEADER
#endif

using namespace LAMMPS_NS;

namespace {
struct Shim {
  LAMMPS lmp;
  NeighList list;
  Pair* pair = nullptr;
  std::vector<double> x, f;
  std::vector<double*> xrow, frow;
  std::vector<int> type, ilist, numneigh, neigh;
  std::vector<tagint> tag;
  std::vector<int*> firstneigh;
  std::vector<int> ghost_owner;
#ifdef SHIM_COMPUTE_HEADER
  std::vector<Compute*> computes;
#endif
  std::string err;
};
template <class F> int guard(Shim* s, F&& fn) {
  try {
    fn();
    return 0;
  } catch (const std::exception& e) {
    s->err = e.what();
    return -1;
  } catch (...) {
    s->err = "unknown C++ exception";
    return -1;
  }
}
}  // namespace

#define API extern "C" __attribute__((visibility("default")))

API void* shim_create(int ntypes, int nlocal, int nghost, const double* x, const int* type, const long long* tag) {
  Shim* s = new Shim();
  const int n = nlocal + nghost;
  s->x.assign(x, x + 3 * (size_t)n);
  s->f.assign(3 * (size_t)n, 0.0);
  s->type.assign(type, type + n);
  s->tag.assign(tag, tag + n);
  s->xrow.resize(n > 0 ? n : 1);
  s->frow.resize(n > 0 ? n : 1);
  for (int i = 0; i < n; ++i) { s->xrow[i] = s->x.data() + 3 * (size_t)i; s->frow[i] = s->f.data() + 3 * (size_t)i; }
  Atom* a = s->lmp.atom;
  a->ntypes = ntypes; a->nlocal = nlocal; a->nghost = nghost; a->nmax = n;
  a->x = s->xrow.data(); a->f = s->frow.data(); a->type = s->type.data(); a->tag = s->tag.data();
  return s;
}

API void shim_destroy(void* p) {
  Shim* s = (Shim*)p;
  if (!s) return;
#ifdef SHIM_COMPUTE_HEADER
  for (Compute* c : s->computes) delete c;
#endif
  delete s->pair;
  delete s;
}

API const char* shim_last_error(void* p) { return ((Shim*)p)->err.c_str(); }

API void shim_set_positions(void* p, const double* x) {
  Shim* s = (Shim*)p;
  memcpy(s->x.data(), x, sizeof(double) * s->x.size());
}
API void shim_zero_forces(void* p) {
  Shim* s = (Shim*)p;
  std::fill(s->f.begin(), s->f.end(), 0.0);
}

// full neighbour list of the local atoms: ilist[inum+gnum], numneigh[ntot], flat neighbours + offsets per atom
API void shim_set_list(void* p, int inum, int gnum, const int* ilist, const int* numneigh, const int* neigh_flat, const long long* first) {
  Shim* s = (Shim*)p;
  const int n = inum + gnum;
  s->ilist.assign(ilist, ilist + n);
  s->numneigh.assign(numneigh, numneigh + n);
  long long tot = 0;
  for (int i = 0; i < n; ++i) tot = std::max(tot, first[i] + numneigh[i]);
  s->neigh.assign(neigh_flat, neigh_flat + tot);
  s->firstneigh.resize(n > 0 ? n : 1);
  for (int i = 0; i < n; ++i) s->firstneigh[i] = s->neigh.data() + first[i];
  s->list.inum = inum; s->list.gnum = gnum;
  s->list.ilist = s->ilist.data(); s->list.numneigh = s->numneigh.data(); s->list.firstneigh = s->firstneigh.data();
  if (s->pair) s->pair->init_list(0, &s->list);
}

API void shim_set_neigh_ago(void* p, int ago) { ((Shim*)p)->lmp.neighbor->ago = ago; }
API int shim_set_newton(void* p, int newton_pair) { ((Shim*)p)->lmp.force->newton_pair = newton_pair; return 0; }

API int shim_pair_create(void* p) {
  Shim* s = (Shim*)p;
  return guard(s, [&] {
    s->pair = new SHIM_PAIR_CLASS(&s->lmp);
    s->lmp.force->pair = s->pair;
    s->pair->init_list(0, &s->list);
  });
}
API int shim_pair_settings(void* p, int narg, char** arg) { Shim* s = (Shim*)p; return guard(s, [&] { s->pair->settings(narg, arg); }); }
API int shim_pair_coeff(void* p, int narg, char** arg) { Shim* s = (Shim*)p; return guard(s, [&] { s->pair->coeff(narg, arg); }); }
API int shim_pair_init_style(void* p) { Shim* s = (Shim*)p; return guard(s, [&] { s->pair->init_style(); }); }
API double shim_pair_init_one(void* p, int i, int j) { return ((Shim*)p)->pair->init_one(i, j); }
API int shim_neigh_request_flags(void* p) { return ((Shim*)p)->lmp.neighbor->last_request.flags; }
API int shim_pair_flags(void* p, int* restartinfo, int* manybody_flag) {
  Shim* s = (Shim*)p;
  *restartinfo = s->pair->restartinfo; *manybody_flag = s->pair->manybody_flag;
  return 0;
}
API int shim_pair_compute(void* p, int eflag, int vflag) { Shim* s = (Shim*)p; return guard(s, [&] { s->pair->compute(eflag, vflag); }); }

API void shim_get_forces(void* p, double* out) { Shim* s = (Shim*)p; memcpy(out, s->f.data(), sizeof(double) * s->f.size()); }
API double shim_get_eng(void* p) { return ((Shim*)p)->pair->eng_vdwl; }
API void shim_get_virial(void* p, double* out6) { memcpy(out6, ((Shim*)p)->pair->virial, sizeof(double) * 6); }
API int shim_get_eatom(void* p, double* out) {
  Shim* s = (Shim*)p;
  if (!s->pair->eatom) return -1;
  memcpy(out, s->pair->eatom, sizeof(double) * (s->lmp.atom->nlocal + s->lmp.atom->nghost));
  return 0;
}

// ---- `compute allegro` / `compute allegro/atom` (reference: compute/compute_allegro.cpp) -------------------
API void shim_set_ghost_owner(void* p, const int* owner /*[nghost]*/) {
  Shim* s = (Shim*)p;
  s->ghost_owner.assign(owner, owner + s->lmp.atom->nghost);
  s->lmp.comm->ghost_owner = s->ghost_owner.data();
}
API void shim_set_timestep(void* p, long long step) { ((Shim*)p)->lmp.update->ntimestep = step; }
#ifdef SHIM_COMPUTE_HEADER
// arg = the words of the LAMMPS command `compute ID group style args...`; returns the compute's index or -1
API int shim_compute_create(void* p, int narg, char** arg) {
  Shim* s = (Shim*)p;
  int idx = -1;
  const int rc = guard(s, [&] {
    if (narg < 3) throw std::runtime_error("compute: too few arguments");
    Compute* c = nullptr;
    if (strcmp(arg[2], "allegro") == 0) c = new SHIM_COMPUTE_TEMPLATE<0>(&s->lmp, narg, arg);
    else if (strcmp(arg[2], "allegro/atom") == 0) c = new SHIM_COMPUTE_TEMPLATE<1>(&s->lmp, narg, arg);
    else throw std::runtime_error(std::string("unknown compute style ") + arg[2]);
    c->init();
    s->computes.push_back(c);
    idx = (int)s->computes.size() - 1;
  });
  return rc == 0 ? idx : -1;
}
API int shim_compute_vector(void* p, int idx, double* out, int n) {
  Shim* s = (Shim*)p;
  return guard(s, [&] {
    Compute* c = s->computes.at(idx);
    if (!c->vector_flag || n != c->size_vector) throw std::runtime_error("compute is not a global vector of that length");
    c->compute_vector();
    memcpy(out, c->vector, sizeof(double) * n);
  });
}
// out[nlocal][ncols]
API int shim_compute_peratom(void* p, int idx, double* out, int ncols) {
  Shim* s = (Shim*)p;
  return guard(s, [&] {
    Compute* c = s->computes.at(idx);
    if (!c->peratom_flag || ncols != (c->size_peratom_cols ? c->size_peratom_cols : 1)) throw std::runtime_error("compute is not per-atom with that many columns");
    c->compute_peratom();
    const int nlocal = s->lmp.atom->nlocal;
    for (int i = 0; i < nlocal; ++i)
      for (int j = 0; j < ncols; ++j) out[(size_t)i * ncols + j] = c->size_peratom_cols ? c->array_atom[i][j] : c->vector_atom[i];
  });
}
API long long shim_compute_invoked(void* p, int idx, int peratom) {
  Compute* c = ((Shim*)p)->computes.at(idx);
  return peratom ? c->invoked_peratom : c->invoked_vector;
}
#endif
