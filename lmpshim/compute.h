// lmpshim: the slice of LAMMPS' Compute base class that `compute allegro[/atom]` uses
// (reference: /root/reference/compute/compute_allegro.{h,cpp}); written from the public LAMMPS
// developer documentation, see lammps.h for the scope of this harness.
#pragma once
#include <vector>

#include "lammps.h"
namespace LAMMPS_NS {
class Compute : protected Pointers {
 public:
  int vector_flag = 0, extvector = 0, size_vector = 0;
  int peratom_flag = 0, size_peratom_cols = 0, comm_reverse = 0;
  double* vector = nullptr;
  double* vector_atom = nullptr;
  double** array_atom = nullptr;
  bigint invoked_vector = -1, invoked_peratom = -1;
  int copymode = 0;

  Compute(LAMMPS* lmp, int /*narg*/, char** /*arg*/) : Pointers(lmp) {}
  virtual void init() = 0;
  virtual void compute_vector() {}
  virtual void compute_peratom() {}
  virtual int pack_reverse_comm(int, int, double*) { return 0; }
  virtual void unpack_reverse_comm(int, int*, double*) {}
};

// single-rank reverse communication of a compute's per-atom values: every ghost is an image of a
// local atom (Comm::ghost_owner, set by the harness), so "send to the owner" is a local fold
inline void Comm::reverse_comm(Compute* c) {
  if (!atom || c->comm_reverse <= 0 || atom->nghost == 0) return;
  const int n = atom->nghost;
  std::vector<double> buf((size_t)n * c->comm_reverse);
  c->pack_reverse_comm(n, atom->nlocal, buf.data());
  c->unpack_reverse_comm(n, ghost_owner, buf.data());
}
}  // namespace LAMMPS_NS
#define ComputeStyle(key, Class)
