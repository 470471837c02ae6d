/* see compute_allegro_b200.h; behaviour follows /root/reference/compute/compute_allegro.cpp
   (constructor :40-85, compute_vector :104-130, compute_peratom :133-162, reverse comm :165-189) */
#include "compute_allegro_b200.h"

#include "atom.h"
#include "comm.h"
#include "error.h"
#include "force.h"
#include "memory.h"
#include "pair_allegro_b200.h"
#include "update.h"

#include <cstdlib>
#include <cstring>
#include <vector>

using namespace LAMMPS_NS;

template <int peratom>
ComputeAllegroB200<peratom>::ComputeAllegroB200(LAMMPS *lmp, int narg, char **arg) : Compute(lmp, narg, arg)
{
  // compute ID all allegro quantity length            | compute ID all allegro/atom quantity length newton
  if (!peratom && narg != 5) error->all(FLERR, "Incorrect args for compute allegro");
  if (peratom && narg != 6) error->all(FLERR, "Incorrect args for compute allegro/atom");
  if (strcmp(arg[1], "all") != 0) error->all(FLERR, "compute allegro can only operate on group 'all'");

  quantity = arg[3];
  const int length = std::atoi(arg[4]);
  if (peratom) {
    peratom_flag = 1;
    nperatom = length;
    if (nperatom <= 0) error->all(FLERR, "Incorrect vector length!");
    newton = std::atoi(arg[5]);
    if (newton) comm_reverse = nperatom;
    size_peratom_cols = nperatom == 1 ? 0 : nperatom;    // LAMMPS: 0 columns = per-atom vector
    if (comm->me == 0)
      error->message(FLERR, "compute allegro/atom will evaluate the quantity {} of length {} with newton {}", quantity,
                     size_peratom_cols, newton);
  } else {
    vector_flag = 1;
    extvector = 1;    // vector quantities are assumed extensive (reference README)
    size_vector = length;
    if (size_vector <= 0) error->all(FLERR, "Incorrect vector length!");
    memory->create(vector, size_vector, "ComputeAllegro:vector");
    if (comm->me == 0)
      error->message(FLERR, "compute allegro will evaluate the quantity {} of length {}", quantity, size_vector);
  }

  if (force->pair == nullptr) error->all(FLERR, "no pair style; compute allegro must be defined after pair style");
  auto *pair = dynamic_cast<PairAllegroB200 *>(force->pair);
  if (pair == nullptr) error->all(FLERR, "compute allegro requires pair_style allegro");
  pair->add_custom_output(quantity);
}

template <int peratom> ComputeAllegroB200<peratom>::~ComputeAllegroB200()
{
  if (copymode) return;
  memory->destroy(array_atom);
  memory->destroy(vector);
}

template <int peratom> void ComputeAllegroB200<peratom>::compute_vector()
{
  invoked_vector = update->ntimestep;
  for (int i = 0; i < size_vector; i++) vector[i] = 0.0;
  if (atom->nlocal > 0) {    // an empty domain stores nothing on the pair style but still joins the reduction
    const std::vector<double> &q = static_cast<PairAllegroB200 *>(force->pair)->custom_output.at(quantity);
    if ((int) q.size() != size_vector)
      error->one(FLERR, "size {} of quantity tensor {} does not match expected {} on rank {}", q.size(), quantity, size_vector,
                 comm->me);
    for (int i = 0; i < size_vector; i++) vector[i] = q[i];
  }
  MPI_Allreduce(MPI_IN_PLACE, vector, size_vector, MPI_DOUBLE, MPI_SUM, world);
}

template <int peratom> void ComputeAllegroB200<peratom>::compute_peratom()
{
  invoked_peratom = update->ntimestep;
  if (atom->nmax > nmax || array_atom == nullptr) {
    nmax = atom->nmax;
    memory->destroy(array_atom);
    memory->create(array_atom, nmax, nperatom, "allegro/atom:array");
    vector_atom = nperatom == 1 ? &array_atom[0][0] : nullptr;
  }
  const int nlocal = atom->nlocal;
  if (nlocal > 0) {
    const std::vector<double> &q = static_cast<PairAllegroB200 *>(force->pair)->custom_output.at(quantity);
    const size_t need = (size_t) (newton ? nlocal + atom->nghost : nlocal) * nperatom;
    if (q.size() < need || q.size() % nperatom != 0)
      error->one(FLERR, "size {} of quantity tensor {} does not match expected {} on rank {}", q.size(), quantity, need, comm->me);
    rows = q.data();
    for (int i = 0; i < nlocal; i++)
      for (int j = 0; j < nperatom; j++) array_atom[i][j] = rows[(size_t) i * nperatom + j];
  } else {
    // empty domain: pair->compute returned early (cpp:341) and produced nothing this step; the ghost rows this rank
    // still has to send in the reverse communication are zeros (never a stale or null pointer)
    zero_rows.assign((size_t) (atom->nlocal + atom->nghost) * nperatom, 0.0);
    rows = zero_rows.data();
  }
  if (newton) comm->reverse_comm(this);    // ghost rows are added to their owners, even if this domain is empty
}

template <int peratom> int ComputeAllegroB200<peratom>::pack_reverse_comm(int n, int first, double *buf)
{
  int m = 0;
  for (int i = first; i < first + n; i++)
    for (int j = 0; j < nperatom; j++) buf[m++] = rows[(size_t) i * nperatom + j];
  return m;
}

template <int peratom> void ComputeAllegroB200<peratom>::unpack_reverse_comm(int n, int *list, double *buf)
{
  int m = 0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < nperatom; j++) array_atom[list[i]][j] += buf[m++];
}

namespace LAMMPS_NS {
template class ComputeAllegroB200<0>;
template class ComputeAllegroB200<1>;
}    // namespace LAMMPS_NS
