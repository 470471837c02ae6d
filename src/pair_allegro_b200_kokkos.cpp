/* ----------------------------------------------------------------------
   pair_style allegro/kk on the B200-native backend: the device-resident twin of PairAllegroB200.
   Mirrors /root/reference/pair_nequip_allegro_kokkos.cpp:
     constructor / masks :40-69, compute :87-353, coeff :365-386 (device copies of type map and cutoff matrix: here
     alg_set_type_map, already done by the base class), init_style :393-406.
   What the reference does with eight Kokkos kernels + libtorch (short neighbour list :165-195, scan + BLOCKING edge count
   :196-206, padded edge / position tensors :218-264, model call :292, force store :305-319) happens behind ONE
   asynchronous C-ABI call; only the scalars the caller asked for (eng_vdwl, virial) synchronise.
------------------------------------------------------------------------- */
#include "pair_allegro_b200_kokkos.h"

#include <string>

#include "allegro_b200.h"
#ifndef ALLEGRO_B200_KOKKOS_SHIM
#include "atom_kokkos.h"
#include "atom_masks.h"
#include "comm.h"
#include "error.h"
#include "force.h"
#include "kokkos.h"
#include "memory_kokkos.h"
#include "neigh_list_kokkos.h"
#include "neigh_request.h"
#include "neighbor.h"
#endif

using namespace LAMMPS_NS;

// kokkos.cpp:40-69
PairAllegroB200Kokkos::PairAllegroB200Kokkos(LAMMPS *lmp) : PairAllegroB200(lmp)
{
  respa_enable = 0;
  atomKK = (AtomKokkos *) atom;
#ifdef ALLEGRO_B200_KOKKOS_SHIM
  static MemoryKokkos shim_memory;
  memoryKK = &shim_memory;
#else
  memoryKK = (MemoryKokkos *) memory;
#endif
  execution_space = ExecutionSpaceFromDevice<DeviceType>::space;
  datamask_read = X_MASK | F_MASK | TAG_MASK | TYPE_MASK | ENERGY_MASK | VIRIAL_MASK;
  datamask_modify = F_MASK | ENERGY_MASK | VIRIAL_MASK;
  // the model virial is ASSIGNED (host path, pair_nequip_allegro.cpp:382-393); the reference's Kokkos twin may add an
  // f.r virial on top when vflag_fdotr is set (kokkos.cpp:344) -- a double count this backend does not reproduce
  no_virial_fdotr_compute = 1;
}

// kokkos.cpp:75-83
PairAllegroB200Kokkos::~PairAllegroB200Kokkos()
{
  if (!copymode) {
    memoryKK->destroy_kokkos(k_eatom, eatom);
    eatom = nullptr;
  }
}

// kokkos.cpp:365-386: the device copies of type_mapper / cutoff_matrix live inside the handle (alg_set_type_map, called by
// the base class); the Kokkos path filters with a strict `<` (kokkos.cpp:189), the host path with `<=` (cpp:507)
void PairAllegroB200Kokkos::coeff(int narg, char **arg)
{
  PairAllegroB200::coeff(narg, arg);
  if (alg_set_option(handle, "filter", "lt") != ALG_OK) error->all(FLERR, "pair_allegro/kk: {}", std::string(alg_last_error(handle)));
  max_neighs_told = -1;
}

// kokkos.cpp:393-406
void PairAllegroB200Kokkos::init_style()
{
  PairAllegroB200::init_style();
  auto request = neighbor->find_request(this);
  request->set_kokkos_host(std::is_same<DeviceType, LMPHostType>::value && !std::is_same<DeviceType, LMPDeviceType>::value);
  request->set_kokkos_device(std::is_same<DeviceType, LMPDeviceType>::value);
  neighflag = ((KokkosLMP *) lmp->kokkos)->neighflag;
  if (neighflag == FULL) error->all(FLERR, "pair style allegro/kk requires the 'neigh half' flag due to 'newton on'");
}

// kokkos.cpp:87-353
void PairAllegroB200Kokkos::compute(int eflag_in, int vflag_in)
{
  ev_init(eflag_in, vflag_in, 0);

  if (eflag_atom) {    // (re)allocate the per-atom energy on both sides (kokkos.cpp:99-103)
    maxeatom = atom->nlocal + atom->nghost > maxeatom ? atom->nlocal + atom->nghost : maxeatom;
    memoryKK->destroy_kokkos(k_eatom, eatom);
    memoryKK->create_kokkos(k_eatom, eatom, maxeatom, "pair:eatom");
    d_eatom = k_eatom.view<DeviceType>();
  }
  if (vflag_atom) error->all(FLERR, "Pair style Allegro does not support per-atom virial");

  atomKK->sync(execution_space, datamask_read);
  if (eflag_in || vflag_in) atomKK->modified(execution_space, datamask_modify);
  else atomKK->modified(execution_space, F_MASK);

  x = atomKK->k_x.view<DeviceType>();
  f = atomKK->k_f.view<DeviceType>();
  tag = atomKK->k_tag.view<DeviceType>();
  type = atomKK->k_type.view<DeviceType>();
  newton_pair = force->newton_pair;

  const int inum = list->inum;
  NeighListKokkos<DeviceType> *k_list = static_cast<NeighListKokkos<DeviceType> *>(list);
  d_ilist = k_list->d_ilist;
  d_numneigh = k_list->d_numneigh;
  d_neighbors = k_list->d_neighbors;
  if (inum == 0) return;    // empty domain (kokkos.cpp:128)

  copymode = 1;
  static_assert(sizeof(X_FLOAT) == sizeof(double) && sizeof(F_FLOAT) == sizeof(double), "allegro/kk needs a double-precision KOKKOS build");
  // extent(1) of the neighbour view bounds the edge count: the call then never has to look at the device (no blocking
  // `nedges` read-back as in kokkos.cpp:203-206)
  const long max_neighs = (long) d_neighbors.extent(1);
  if (max_neighs != max_neighs_told) {
    alg_set_option(handle, "max_neighbors", std::to_string(max_neighs).c_str());
    max_neighs_told = max_neighs;
  }
  const bool want_scalars = eflag_global || vflag_global;
  double eng = 0.0, vir[6] = {0, 0, 0, 0, 0, 0};
  const int rc = alg_compute_device(handle, inum, atom->nghost, &x.data()[0], type.data(), d_ilist.data(), d_numneigh.data(), d_neighbors.data(),
                                    (int64_t) d_neighbors.stride(0), (int64_t) d_neighbors.stride(1), eflag_atom ? 1 : 0, vflag_global ? 1 : 0,
                                    f.data(), eflag_atom ? d_eatom.data() : nullptr, want_scalars ? &eng : nullptr, want_scalars ? vir : nullptr,
                                    nullptr /* Kokkos' default CUDA stream */);
  if (rc != ALG_OK) { copymode = 0; error->one(FLERR, "pair_allegro/kk: {}", std::string(alg_last_error(handle))); }

  if (eflag_global) eng_vdwl = eng;                               // sum over the local atoms (kokkos.cpp:303-319)
  if (eflag_atom) {
    k_eatom.modify<DeviceType>();
    k_eatom.sync<LMPHostType>();
  }
  if (vflag_global)
    for (int q = 0; q < 6; q++) virial[q] = vir[q];               // xx yy zz xy xz yz (kokkos.cpp:328-339)

  for (const std::string &output_name : custom_output_names) {
    const double *ptr;
    int64_t n;
    if (alg_get_output(handle, output_name.c_str(), &ptr, &n) != ALG_OK) { copymode = 0; error->all(FLERR, "missing {}", output_name); }
    custom_output[output_name].assign(ptr, ptr + n);
  }
  copymode = 0;
}
