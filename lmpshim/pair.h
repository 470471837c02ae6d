#pragma once
#include <cstring>

#include "lammps.h"
namespace LAMMPS_NS {
class Pair : protected Pointers {
 public:
  double eng_vdwl = 0.0, eng_coul = 0.0;
  double virial[6] = {0, 0, 0, 0, 0, 0};
  double* eatom = nullptr;
  double** vatom = nullptr;
  int restartinfo = 1, manybody_flag = 0, respa_enable = 0, no_virial_fdotr_compute = 0;
  int allocated = 0, copymode = 0;
  int** setflag = nullptr;
  double** cutsq = nullptr;
  NeighList* list = nullptr;
  int eflag_either = 0, eflag_global = 0, eflag_atom = 0;
  int vflag_either = 0, vflag_global = 0, vflag_atom = 0, vflag_fdotr = 0;
  int maxeatom = 0, maxvatom = 0;

  explicit Pair(LAMMPS* lmp) : Pointers(lmp) {}
  ~Pair() override { free(eatom); }
  virtual void compute(int, int) = 0;
  virtual void settings(int, char**) = 0;
  virtual void coeff(int, char**) = 0;
  virtual void init_style() {}
  virtual double init_one(int, int) { return 0.0; }
  virtual void init_list(int, NeighList* ptr) { list = ptr; }

  // LAMMPS: eflag bit 1 = global energy, 2 = per-atom; vflag bits 1|2 = global virial, 4 = per-atom
  void ev_init(int eflag, int vflag, int /*alloc*/ = 1) {
    eflag_either = eflag; eflag_global = eflag & 1; eflag_atom = eflag & 2;
    vflag_either = vflag; vflag_global = vflag & 3; vflag_atom = vflag & 4;
    eng_vdwl = eng_coul = 0.0;
    for (double& v : virial) v = 0.0;
    const int n = atom->nlocal + atom->nghost;
    if (eflag_atom) {
      if (n > maxeatom) { free(eatom); maxeatom = n; eatom = (double*)malloc(sizeof(double) * (n > 0 ? n : 1)); }
      memset(eatom, 0, sizeof(double) * n);
    }
  }
};
}  // namespace LAMMPS_NS
#define PairStyle(key, Class)
