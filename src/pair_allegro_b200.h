/* -*- c++ -*- ----------------------------------------------------------
   pair_style allegro (B200-native backend)

   Drop-in for the reference's PairNequIPAllegro<false> (pair_nequip_allegro.h:41-94): same
   LAMMPS-facing virtuals, same public members, same pair_coeff syntax.  The model execution
   (libtorch in the reference) is replaced by the C-ABI of include/allegro_b200.h.
------------------------------------------------------------------------- */
#ifdef PAIR_CLASS
// clang-format off
PairStyle(allegro,PairAllegroB200)
// clang-format on
#else
#ifndef LMP_PAIR_ALLEGRO_B200_H
#define LMP_PAIR_ALLEGRO_B200_H

#include "pair.h"

#include <map>
#include <string>
#include <vector>

struct alg_handle;

namespace LAMMPS_NS {

class PairAllegroB200 : public Pair {
 public:
  PairAllegroB200(class LAMMPS *);
  ~PairAllegroB200() override;
  void compute(int, int) override;
  void settings(int, char **) override;
  void coeff(int, char **) override;
  double init_one(int, int) override;
  void init_style() override;
  void allocate();

  double cutoff;
  int device_index = 0;
  std::vector<int> type_mapper;
  std::string model_path;

  // in/out precision at the LAMMPS boundary (pair_nequip_allegro.h:73-75)
  typedef double inputtype;
  typedef double outputtype;

  // hook used by `compute allegro[/atom]` (pair_nequip_allegro.h:80-82)
  std::vector<std::string> custom_output_names;
  std::map<std::string, std::vector<double>> custom_output;
  void add_custom_output(std::string);

 protected:
  int debug_mode = 0;
  double **cutoff_matrix = nullptr;
  alg_handle *handle = nullptr;
  std::string resolve_weight_path(const std::string &) const;
};

}    // namespace LAMMPS_NS
#endif
#endif
