/* ----------------------------------------------------------------------
   pair_style allegro, B200-native backend.

   Host-side pair style with the behaviour of the reference's PairNequIPAllegro<false>
   (/root/reference/pair_nequip_allegro.cpp); each method cites the lines it mirrors.  All
   numerical work is delegated to liballegro_b200.so through the C-ABI (include/allegro_b200.h):
   coeff() -> alg_create/alg_metadata/alg_set_type_map, compute() -> alg_compute_host.
   Compiles against real LAMMPS headers or against lmpshim/ (tests).
------------------------------------------------------------------------- */
#include "pair_allegro_b200.h"

#include "atom.h"
#include "comm.h"
#include "error.h"
#include "force.h"
#include "memory.h"
#include "neigh_list.h"
#include "neigh_request.h"
#include "neighbor.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <sys/stat.h>

#include <mpi.h>

#include "allegro_b200.h"

using namespace LAMMPS_NS;

// cpp:66-125
PairAllegroB200::PairAllegroB200(LAMMPS *lmp) : Pair(lmp)
{
  restartinfo = 0;
  manybody_flag = 1;

  if (comm->me == 0)
    std::cout << "Allegro (B200 backend " << alg_version() << ") is using input precision double and output precision double"
              << std::endl;

  if (const char *env_p = std::getenv("_NEQUIP_LOG_LEVEL")) {
    if (std::string(env_p) == "DEBUG") {
      std::cout << "Debug mode enabled, since _NEQUIP_LOG_LEVEL is set to DEBUG\n";
      debug_mode = 1;
    }
  }

  // device = node-local rank (cpp:91-120).  More ranks than visible devices is an error, except in debug mode
  // where the rank wraps around so that multi-rank tests can share one GPU (cpp:102-118).
  device_index = 0;
  if (comm->nprocs > 1) {
    MPI_Comm shmcomm;
    MPI_Comm_split_type(MPI_COMM_WORLD, MPI_COMM_TYPE_SHARED, 0, MPI_INFO_NULL, &shmcomm);
    int shmrank;
    MPI_Comm_rank(shmcomm, &shmrank);
    device_index = shmrank;
  }
  if (const char *env_d = std::getenv("ALLEGRO_B200_DEVICE")) device_index = std::atoi(env_d);
  const int devicecount = alg_device_count();
  if (devicecount > 0 && device_index >= devicecount) {
    if (debug_mode) {
      std::cerr << "WARNING (Allegro): my rank (" << device_index << ") is bigger than the number of visible devices (" << devicecount
                << "), wrapping around to use device " << device_index % devicecount << " again!!!";
      device_index = device_index % devicecount;
    } else {
      std::cerr << "ERROR (Allegro): my rank (" << device_index << ") is bigger than the number of visible devices (" << devicecount << ")!!!";
      error->all(FLERR, "pair_allegro: mismatch between number of ranks and number of available GPUs");
    }
  }
  if (debug_mode) std::cout << "Allegro is using device cuda:" << device_index << "\n";
}

// cpp:127-135
PairAllegroB200::~PairAllegroB200()
{
  if (copymode) return;
  if (handle) alg_destroy(handle);
  if (allocated) {
    memory->destroy(setflag);
    memory->destroy(cutsq);
    memory->destroy(cutoff_matrix);
  }
}

// cpp:137-151
void PairAllegroB200::init_style()
{
  if (atom->tag_enable == 0) error->all(FLERR, "Pair style Allegro requires atom IDs");

  // full neighbour list of the local atoms; the host path also asks for ghost lists like the
  // reference does ("to avoid segfaults", cpp:145-146) although only local rows are read
  if (lmp->kokkos) {
    neighbor->add_request(this, NeighConst::REQ_FULL);
  } else {
    neighbor->add_request(this, NeighConst::REQ_FULL | NeighConst::REQ_GHOST);
  }

  if (force->newton_pair == 0) error->all(FLERR, "Pair style allegro requires newton pair on");
}

// cpp:153-156
double PairAllegroB200::init_one(int /*i*/, int /*j*/)
{
  return cutoff;
}

// cpp:158-166
void PairAllegroB200::allocate()
{
  allocated = 1;
  int n = atom->ntypes;

  memory->create(setflag, n + 1, n + 1, "pair:setflag");
  memory->create(cutsq, n + 1, n + 1, "pair:cutsq");
  memory->create(cutoff_matrix, n, n, "pair:cutoff_matrix");
}

// cpp:168-172
void PairAllegroB200::settings(int narg, char ** /*arg*/)
{
  if (narg > 0) error->all(FLERR, "Illegal pair_style command, too many arguments");
}

static bool ends_with(const std::string &s, const std::string &ext)
{
  return s.size() >= ext.size() && s.compare(s.size() - ext.size(), ext.size(), ext) == 0;
}

// `pair_coeff * * <model> <types...>` keeps the reference syntax (cpp:195-206): a
// `.nequip.pth` / `.nequip.pt2` path resolves to the `.alg` file the offline exporter wrote
// next to it; a `.alg` path is used directly.
std::string PairAllegroB200::resolve_weight_path(const std::string &path) const
{
  if (ends_with(path, ".alg")) return path;
  for (const char *ext : {".nequip.pth", ".nequip.pt2"}) {
    if (ends_with(path, ext)) {
      std::string cand = path.substr(0, path.size() - strlen(ext)) + ".alg";
      struct stat st;
      if (stat(cand.c_str(), &st) == 0) return cand;
      // the exporter reads TorchScript files written by this repo's model builder (they carry an `allegro_b200_config`
      // entry); an AOT-Inductor package cannot be read back at all, and a nequip-compile'd TorchScript file needs the
      // nequip/allegro importer that is not part of this build
      if (std::string(ext) == ".nequip.pt2")
        throw std::runtime_error("no exported weights " + cand + " for " + path +
                                 ": an AOT-Inductor package cannot be converted; export the weights of the same model to " + cand +
                                 " (python -m pair_allegro_b200.export <model>.nequip.pth " + cand + ")");
      throw std::runtime_error("no exported weights " + cand + " for " + path +
                               ": run `python -m pair_allegro_b200.export " + path + " " + cand +
                               "` (supports TorchScript files carrying an allegro_b200_config entry)");
    }
  }
  throw std::runtime_error("Only accepts model paths with extension `.nequip.pth`, `.nequip.pt2` or `.alg`, but found" + path);
}

// cpp:174-330
void PairAllegroB200::coeff(int narg, char **arg)
{
  if (!allocated) allocate();

  int ntypes = atom->ntypes;

  for (int i = 1; i <= ntypes; i++)
    for (int j = i; j <= ntypes; j++) setflag[i][j] = 0;

  if (narg != (3 + ntypes)) {
    error->all(FLERR,
               "Incorrect args for pair coefficients, should be * * <model>.nequip.pth/pt2 <type1> <type2> ... <typen>");
  }
  if (strcmp(arg[0], "*") != 0 || strcmp(arg[1], "*") != 0) error->all(FLERR, "Incorrect args for pair coefficients");

  model_path = std::string(arg[2]);
  const std::string weight_path = resolve_weight_path(model_path);
  if (comm->me == 0) std::cout << "Allegro: Loading model from " << weight_path << "\n";
  if (handle) { alg_destroy(handle); handle = nullptr; }
  if (alg_create(weight_path.c_str(), device_index, &handle) != ALG_OK) {
    std::string msg = alg_last_error(nullptr);
    error->all(FLERR, "pair_allegro: {}", msg);
  }

  double r_max;
  int num_model_types, allow_tf32;
  const char *type_names;
  const double *per_edge;
  alg_metadata(handle, &r_max, &num_model_types, &type_names, &per_edge, &allow_tf32);
  if (debug_mode) {
    std::cout << "Allegro: Information from model: r_max=" << r_max << " num_types=" << num_model_types << " type_names=["
              << type_names << "] allow_tf32=" << allow_tf32 << "\n";
  }
  cutoff = r_max;

  type_mapper.assign(ntypes, -1);
  std::stringstream ss;
  ss << type_names;
  if (comm->me == 0) std::cout << "Type mapping:\nAllegro type | Allegro name | LAMMPS type | LAMMPS name\n";
  for (int i = 0; i < num_model_types; i++) {
    std::string ele;
    ss >> ele;
    for (int itype = 1; itype <= ntypes; itype++) {
      if (ele.compare(arg[itype + 3 - 1]) == 0) {
        type_mapper[itype - 1] = i;
        if (comm->me == 0) std::cout << i << " | " << ele << " | " << itype << " | " << arg[itype + 3 - 1] << "\n";
      }
    }
  }

  for (int i = 1; i <= ntypes; i++) {
    for (int j = i; j <= ntypes; j++) {
      if ((type_mapper[i - 1] >= 0) && (type_mapper[j - 1] >= 0)) { setflag[i][j] = 1; }
    }
  }

  if (per_edge) {
    std::vector<int> reverse_type_mapper(num_model_types, -1);
    // the reference indexes reverse_type_mapper[-1] for unmapped LAMMPS types (cpp:308, UB); guarded here
    for (int i = 0; i < ntypes; i++)
      if (type_mapper[i] >= 0) reverse_type_mapper[type_mapper[i]] = i;
    for (int i = 0; i < ntypes; i++)
      for (int j = 0; j < ntypes; j++) cutoff_matrix[i][j] = 0.0;
    for (int i = 0; i < num_model_types; i++) {
      for (int j = 0; j < num_model_types; j++) {
        double cutij = per_edge[i * num_model_types + j];
        if (reverse_type_mapper[i] >= 0 && reverse_type_mapper[j] >= 0) {
          if (comm->me == 0) {
            printf("%s %s si=%d sj=%d ti=%d tj=%d cut=%.2f\n", arg[reverse_type_mapper[i] + 3], arg[reverse_type_mapper[j] + 3], i, j,
                   reverse_type_mapper[i], reverse_type_mapper[j], cutij);
          }
          cutoff_matrix[reverse_type_mapper[i]][reverse_type_mapper[j]] = cutij;
        }
      }
    }
  } else {
    for (int i = 0; i < ntypes; i++) {
      for (int j = 0; j < ntypes; j++) { cutoff_matrix[i][j] = cutoff; }
    }
  }

  std::vector<double> cm((size_t)ntypes * ntypes);
  for (int i = 0; i < ntypes; i++)
    for (int j = 0; j < ntypes; j++) cm[(size_t)i * ntypes + j] = cutoff_matrix[i][j];
  if (alg_set_type_map(handle, ntypes, type_mapper.data(), cm.data()) != ALG_OK)
    error->all(FLERR, "pair_allegro: {}", std::string(alg_last_error(handle)));
  if (debug_mode) alg_set_option(handle, "keep_edges", "1");
}

// cpp:333-407
void PairAllegroB200::compute(int eflag, int vflag)
{
  ev_init(eflag, vflag);

  double **f = atom->f;
  double **x = atom->x;

  int inum = list->inum;
  if (inum == 0) return;
  int nghost = list->gnum;
  int ntotal = inum + nghost;

  if (vflag_atom) { error->all(FLERR, "Pair styles nequip and allegro do not support per-atom virial"); }

  // neighbor->ago == 0 on the steps LAMMPS rebuilt the list: in between the device copy is reused
  alg_set_option(handle, "neigh_ago", std::to_string(neighbor->ago).c_str());

  // atom->x / atom->f are contiguous [ntotal][3] (LAMMPS memory->create layout): x[0], f[0]
  double eng = 0.0, vir[6] = {0, 0, 0, 0, 0, 0};
  int rc = alg_compute_host(handle, inum, nghost, &x[0][0], atom->type, list->ilist, list->numneigh, list->firstneigh,
                            eflag_atom ? 1 : 0, vflag_global ? 1 : 0, &f[0][0], eflag_atom ? eatom : nullptr, &eng, vir);
  if (rc != ALG_OK) error->one(FLERR, "pair_allegro: {}", std::string(alg_last_error(handle)));
  (void) ntotal;

  eng_vdwl = eng;                                           // sum over LOCAL atoms (cpp:379)
  if (vflag) {
    for (int q = 0; q < 6; q++) virial[q] = vir[q];         // xx yy zz xy xz yz (cpp:387-392)
  }

  if (debug_mode) {    // the reference's edge dump (cpp:562-565, 620-633): tag-1 indices and |x_i - x_j|
    const int64_t *ei;
    int64_t ne;
    if (alg_get_edges(handle, &ei, &ne) == ALG_OK) {
      printf("Allegro edges: i j rij\n");
      for (int64_t e = 0; e < ne; e++) {
        const int i = (int) ei[e], j = (int) ei[ne + e];
        const double dx = x[i][0] - x[j][0], dy = x[i][1] - x[j][1], dz = x[i][2] - x[j][2];
        printf("%d %d %.10g\n", (int) atom->tag[i] - 1, (int) atom->tag[j] - 1, sqrt(dx * dx + dy * dy + dz * dz));
      }
      printf("end Allegro edges\n");
    }
  }

  for (const std::string &output_name : custom_output_names) {
    const double *ptr;
    int64_t n;
    if (alg_get_output(handle, output_name.c_str(), &ptr, &n) != ALG_OK) error->all(FLERR, "missing {}", output_name);
    custom_output[output_name].assign(ptr, ptr + n);
  }
}

// cpp:681-684
void PairAllegroB200::add_custom_output(std::string name)
{
  custom_output_names.push_back(name);
}
