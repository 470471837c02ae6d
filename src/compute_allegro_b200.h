/* `compute allegro` / `compute allegro/atom` for pair_style allegro (B200 build).

   Drop-in for /root/reference/compute/compute_allegro.{h,cpp}: same command syntax
       compute ID all allegro      <quantity> <length>
       compute ID all allegro/atom <quantity> <length-per-atom> <newton 0|1>
   same registration hook on the pair style (add_custom_output, pair_nequip_allegro.h:80-82), same
   error messages.  The quantity is read from PairAllegroB200::custom_output (host doubles filled
   from alg_get_output after every compute()) instead of a torch::Tensor. */
#ifdef COMPUTE_CLASS
// clang-format off
ComputeStyle(allegro, ComputeAllegroB200<0>)
ComputeStyle(allegro/atom, ComputeAllegroB200<1>)
// clang-format on
#else

#ifndef LMP_COMPUTE_ALLEGRO_B200_H
#define LMP_COMPUTE_ALLEGRO_B200_H

#include "compute.h"
#include <vector>

#include <string>

namespace LAMMPS_NS {

template <int peratom> class ComputeAllegroB200 : public Compute {
 public:
  ComputeAllegroB200(class LAMMPS *, int, char **);
  ~ComputeAllegroB200() override;
  void init() override {}
  void compute_vector() override;
  void compute_peratom() override;
  int pack_reverse_comm(int, int, double *) override;
  void unpack_reverse_comm(int, int *, double *) override;

 protected:
  std::string quantity;
  const double *rows = nullptr;    // custom_output[quantity] of the current step, [ntot][nperatom]
  std::vector<double> zero_rows;   // what an empty domain sends in the reverse communication
  int newton = 0, nperatom = 0, nmax = 0;
};

}    // namespace LAMMPS_NS
#endif
#endif
