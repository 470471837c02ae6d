/* -*- c++ -*- ----------------------------------------------------------
   pair_style allegro/kk -- B200-native Allegro force evaluation on device-resident LAMMPS/KOKKOS data.
   Drop-in for the reference's PairAllegroKokkos<false> (pair_nequip_allegro_kokkos.h:16, .cpp:87-353): same style
   name, same neighbour request, same `neigh half` requirement; x / type / neighbour views are handed to the C-ABI
   (alg_compute_device, include/allegro_b200.h) as raw device pointers + strides, forces are accumulated on the device.
------------------------------------------------------------------------- */
#ifdef PAIR_CLASS
// clang-format off
PairStyle(allegro/kk,PairAllegroB200Kokkos)
// clang-format on
#else

#ifndef LMP_PAIR_ALLEGRO_B200_KOKKOS_H
#define LMP_PAIR_ALLEGRO_B200_KOKKOS_H

#include "pair_allegro_b200.h"
#ifdef ALLEGRO_B200_KOKKOS_SHIM
#include "kokkos_shim.h"        // test harness (lmpshim/): no LAMMPS / Kokkos in this image
#else
#include "kokkos_type.h"
#include "pair_kokkos.h"
#endif

namespace LAMMPS_NS {

class PairAllegroB200Kokkos : public PairAllegroB200 {
 public:
  using DeviceType = LMPDeviceType;
  enum { EnabledNeighFlags = FULL | HALFTHREAD | HALF };
  enum { COUL_FLAG = 0 };
  typedef LMPDeviceType device_type;
  typedef ArrayTypes<DeviceType> AT;

  PairAllegroB200Kokkos(class LAMMPS *);
  ~PairAllegroB200Kokkos() override;
  void compute(int, int) override;
  void coeff(int, char **) override;
  void init_style() override;

  typename AT::t_efloat_1d d_eatom;

 protected:
  typename AT::t_x_array_randomread x;
  typename AT::t_f_array f;
  typename AT::t_tagint_1d tag;
  typename AT::t_int_1d_randomread type;
  DAT::tdual_efloat_1d k_eatom;

  typename AT::t_neighbors_2d d_neighbors;
  typename AT::t_int_1d_randomread d_ilist;
  typename AT::t_int_1d_randomread d_numneigh;

  AtomKokkos *atomKK = nullptr;
  MemoryKokkos *memoryKK = nullptr;
  ExecutionSpace execution_space = Device;
  unsigned int datamask_read = 0, datamask_modify = 0;
  int neighflag = 0, newton_pair = 1;
  long max_neighs_told = -1;
};

}    // namespace LAMMPS_NS
#endif
#endif
