/* allegro_b200.h -- C-ABI of the B200-native Allegro force evaluation.
 *
 * Drop-in boundary for the hot path of mir-group/pair_allegro: everything the reference
 * does between `PairNequIPAllegro<false>::compute()` entering and leaving
 * (pair_nequip_allegro.cpp:333-407: preprocess() :457-650, call() :409-454 i.e. the whole
 * libtorch model execution, and the output store :358-393) and its Kokkos twin
 * (pair_nequip_allegro_kokkos.cpp:87-353) happens behind these entry points, in hand-written
 * sm_100a CUDA kernels.  No C++ types, no libtorch, no exceptions cross this boundary.
 *
 * Conventions
 *   - every function returns 0 on success or a negative ALG_E* code; the message is
 *     available through alg_last_error() (never throws; the host pair style converts to
 *     LAMMPS error->all / error->one, cf. pair_nequip_allegro.cpp:87,115,139,149,186,394).
 *   - the caller owns every pointer it passes; the handle owns all device scratch (grown
 *     geometrically, never per-step cudaMalloc in steady state); pointers returned by
 *     alg_get_* stay valid until the next alg_compute_* / alg_destroy on that handle.
 *   - one handle = one CUDA device = one caller thread (one LAMMPS rank), all work on one
 *     stream.  alg_compute_host is synchronous; alg_compute_device is stream-ordered except
 *     for the scalar outputs (see below).
 *   - dtypes at the boundary are LAMMPS': double x/f/energies/virial
 *     (pair_nequip_allegro.h:73-75), 32-bit int indices; int64 only in alg_get_edges
 *     (the reference's edge_index tensor, pair_nequip_allegro.cpp:526-527).
 */
#ifndef ALLEGRO_B200_H
#define ALLEGRO_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define ALG_API __attribute__((visibility("default")))
#else
#define ALG_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct alg_handle alg_handle;

enum {
  ALG_OK = 0,
  ALG_EINVAL = -1,   /* bad argument / unsupported model hyper-parameters */
  ALG_EIO = -2,      /* weight file missing or malformed */
  ALG_ECUDA = -3,    /* CUDA runtime error (no device, out of memory, launch failure) */
  ALG_ESTATE = -4,   /* call order violated (e.g. compute before set_type_map) */
  ALG_ENOTFOUND = -5 /* unknown output / option name */
};

/* Replaces the model load in coeff(): torch::jit::load(model_path, device, metadata)
 * (pair_nequip_allegro.cpp:213-232) / AOTIModelPackageLoader (:238).  `weight_path` is the
 * `.alg` file written by the offline exporter (python -m pair_allegro_b200.export).
 * `cuda_device` is the device index chosen by the pair style (node-local rank,
 * pair_nequip_allegro.cpp:91-120).  There is no CPU fallback: without a usable device this
 * returns ALG_ECUDA. */
ALG_API int alg_create(const char* weight_path, int cuda_device, alg_handle** out);
ALG_API void alg_destroy(alg_handle* h);

/* message of the last failing call on `h`; with h == NULL: of the last failing alg_create. */
ALG_API const char* alg_last_error(const alg_handle* h);

/* The five metadata keys the reference reads from the model file
 * (pair_nequip_allegro.cpp:214-220, used at :267-328).  `type_names` is space separated;
 * `per_edge_type_cutoff` is num_types x num_types row-major [centre][neighbour] in MODEL type
 * order, or NULL when the model has a single r_max.  Any out pointer may be NULL. */
ALG_API int alg_metadata(const alg_handle* h, double* r_max, int* num_types, const char** type_names,
                 const double** per_edge_type_cutoff, int* allow_tf32);

/* Replaces the tail of coeff() (pair_nequip_allegro.cpp:274-328; Kokkos copy
 * pair_nequip_allegro_kokkos.cpp:365-386): `lammps_type_to_model[t-1]` = model type index of
 * LAMMPS type t (type_mapper, -1 = unmapped), `cutoff_matrix` = ntypes x ntypes row-major,
 * indexed [centre LAMMPS type-1][neighbour LAMMPS type-1] (may be asymmetric).  An unmapped LAMMPS type is legal as
 * long as no atom carries it (the reference fails inside the model's type embedding otherwise): a compute that meets
 * such an atom builds no edge for it and returns ALG_EINVAL ("atom with a LAMMPS type that has no model type"). */
ALG_API int alg_set_type_map(alg_handle* h, int ntypes, const int* lammps_type_to_model,
                     const double* cutoff_matrix);

/* Options (string key/value, all optional):
 *   "filter"       "le" (default; host path rsq <= cut^2, pair_nequip_allegro.cpp:507,599)
 *                  | "lt" (Kokkos path rsq < cut^2, pair_nequip_allegro_kokkos.cpp:189)
 *   "pipeline"     "auto" (default) | "fused" | "tiled".  fused = ONE persistent kernel: the edge list is cut on the device
 *                  into centre-aligned batches of fused_batch*128 edges, a CTA runs a batch through all phases and
 *                  combines the per-atom sums itself; inter-phase state in CTA-private scratch; no host synchronisation,
 *                  no fix-up launches.  Needs <= fused_batch*128 neighbours inside the cutoff per atom.  tiled = the
 *                  chunked edge-tile pipeline (any neighbour count; one kernel per phase and chunk, fix-up kernels
 *                  between them; see chunk_plan).  auto = tiled: measured 4 % (l_max 1) to 50 % (l_max 3) faster on
 *                  the B200, and with the device-built chunk plan just as free of host synchronisation.
 *   "fused_batch"  fused pipeline: 128-edge tiles per batch (default 8).  A CTA runs the tiles of a batch phase by phase, so
 *                  the code of one phase stays in the instruction cache for the whole batch
 *   "phase_align"  fused pipeline: "1" (default) keeps the CTAs that share an SM in the same phase (bounded wait at every phase
 *                  change) so that they share the instruction cache; "0" lets them drift
 *   "max_neighbors" alg_compute_device only: extent(1) of the caller's 2-D neighbour view.  Sizes the edge arrays to
 *                  nlocal*max_neighbors so that a fully asynchronous step can never overflow them; without it they
 *                  are sized from the previous step's edge count (+12.5 %), like the reference's 1.05 padding
 *                  (pair_nequip_allegro_kokkos.cpp:218-229)
 *   "host_register" "1" (default) | "0": alg_compute_host pins the caller's x / f / type arrays with cudaHostRegister
 *                  (cached per array address and size; LAMMPS keeps atom->x / f at the same address until they grow)
 *                  so that the per-step copies are asynchronous DMA.  Switch off when the caller frees and
 *                  re-allocates these arrays between calls.
 *   "chunk_edges"  tiled pipeline: edges processed per pipeline pass (activation buffers are sized by this)
 *   "chunk_plan"   tiled pipeline: "device" (default) builds the centre-aligned chunk plan on the device -- no host
 *                  synchronisation; grids and launch counts are sized by the capacity of the edge arrays, surplus CTAs and
 *                  launches exit at once -- | "host" copies the CSR row pointer to the host every step (round-1 behaviour;
 *                  also used with debug=1 and after a step whose chunks did not fit the device plan's buffers)
 *   "keep_edges"   "1": materialise the int64 [2,E] edge_index for alg_get_edges
 *   "debug"        "1": keep per-edge gradients / intermediates for alg_get_output
 *   "profile"      "1": time every pipeline kernel with CUDA events (alg_get_stats)
 *   "gemm"         "tc": dense contractions on the tcgen05 tensor cores (default; l_max = 1..3) | "ffma": the same tiled
 *                  kernels on the FP32 pipe | "generic": width-generic per-operation kernels (csrc/alg_generic.cu).  "tc" and
 *                  "ffma" exist for num_scalar_features=64, num_tensor_features=32, MLP 2x64, readout 32; a model of any
 *                  other widths (mlp_depth 1..4) loads and runs on "generic" automatically (slower: no tensor cores, every
 *                  intermediate through HBM, one host synchronisation per step)
 *   "precision"    "strict": fp32-level accuracy (3xTF32 split on the tensor cores) | "tf32": one
 *                  TF32 pass (fast mode, ~1e-3 relative accuracy; tensor-core path only)
 *   "neigh_ago"    LAMMPS' neighbor->ago (steps since the last neighbour-list rebuild), set before
 *                  alg_compute_host: when > 0 and the atom counts are unchanged, the device copy of the
 *                  list uploaded by the previous call is reused (the reference re-reads and re-uploads
 *                  its edge tensors every step, pair_nequip_allegro.cpp:524-529,638-641); default 0 */
ALG_API int alg_set_option(alg_handle* h, const char* key, const char* value);

/* Host-pointer force evaluation: replaces the body of PairNequIPAllegro<false>::compute()
 * (pair_nequip_allegro.cpp:333-407).
 *   nlocal, nghost : atom->nlocal (== list->inum, :470) and list->gnum (:344)
 *   x              : atom->x, [nlocal+nghost][3] contiguous doubles
 *   type           : atom->type, 1-based LAMMPS types, [nlocal+nghost]
 *   ilist,numneigh,firstneigh : the FULL neighbour list of the local atoms (:340-350,476-480);
 *                    numneigh/firstneigh are indexed by atom index i = ilist[ii];
 *                    neighbour entries are masked with NEIGHMASK (:496)
 *   eflag_atom     : write eatom[i] (assigned, locals only, :378) when non-zero and eatom != NULL
 *   vflag_global   : write virial6 (assigned, LAMMPS order xx,yy,zz,xy,xz,yz, :387-392)
 *   f              : atom->f, [nlocal+nghost][3]; model forces are ADDED for local AND ghost
 *                    atoms (newton on, :370-377); LAMMPS reverse-communicates afterwards
 *   eng            : eng_vdwl = sum of LOCAL atomic energies (:379)
 * nlocal == 0 is a valid no-op (:341). */
ALG_API int alg_compute_host(alg_handle* h, int nlocal, int nghost, const double* x, const int* type,
                     const int* ilist, const int* numneigh, int* const* firstneigh,
                     int eflag_atom, int vflag_global,
                     double* f, double* eatom, double* eng, double* virial6);

/* Device-pointer force evaluation: replaces PairAllegroKokkos<false>::compute()
 * (pair_nequip_allegro_kokkos.cpp:87-353).  All d_* pointers are device memory on the
 * handle's device.  Neighbours are the Kokkos 2-D view d_neighbors(i,jj) addressed as
 * d_neighbors[i*stride_i + jj*stride_jj] (:126,176).  d_f is accumulated in place (:308-310),
 * d_eatom (may be NULL) assigned for locals (:311-313).  `eng` and `virial6` are HOST
 * pointers written before return (the call synchronises the stream once for them, like
 * the reference's parallel_reduce result :318 and virial .cpu() :329); pass NULL for both
 * to keep the call fully asynchronous: nothing in the call then waits for the device (the chunk plan of the default
 * chunked pipeline and the tile plan of the fused pipeline are built on the device; the reference blocks on its edge count
 * every step, pair_nequip_allegro_kokkos.cpp:203-206; models on the width-generic pipeline, option gemm=generic, build their
 * chunk plan on the host and synchronise once inside the call).  Such a step is verified lazily: if it met an atom with more than fused_batch*128
 * neighbours, or overflowed edge arrays sized without max_neighbors, it wrote no forces and the NEXT call on the handle
 * returns ALG_ESTATE (then switches to the tiled pipeline / larger arrays).  Calls that pass `eng` or `virial6` are
 * verified before they return and such steps are repeated transparently.  `stream` is a cudaStream_t (0 = legacy default). */
ALG_API int alg_compute_device(alg_handle* h, int nlocal, int nghost, const double* d_x, const int* d_type,
                       const int* d_ilist, const int* d_numneigh, const int* d_neighbors,
                       int64_t stride_i, int64_t stride_jj,
                       int eflag_atom, int vflag_global,
                       double* d_f, double* d_eatom, double* eng, double* virial6, void* stream);

/* The edge list of the last compute in the reference's tensor layout
 * (edge_index int64 [2][E], row 0 = centre atom index, row 1 = neighbour atom index with
 * ghost indices kept; pair_nequip_allegro.cpp:601-602).  Host memory.  Requires option
 * keep_edges=1 before the compute.  This is the parity hook for the reference's debug
 * dump "Allegro edges: i j rij" (:562-565,620-633). */
ALG_API int alg_get_edges(alg_handle* h, const int64_t** edge_index, int64_t* nedges);

/* Named outputs of the last compute (host memory, doubles) -- the hook `compute allegro`
 * / `compute allegro/atom` use (compute/compute_allegro.cpp:113-116,148-150 read
 * pair->custom_output[name]).  Always available: "atomic_energy" [ntot] (ghost rows =
 * per-type shift, as the model returns them), "forces" [ntot*3], "virial" [9],
 * "edge_energy" [E] (needs debug=1).  With debug=1 single-chunk runs also expose
 * intermediates for parity tests ("x0","x1","gamma0","edge_grad",...). */
ALG_API int alg_get_output(alg_handle* h, const char* name, const double** ptr, int64_t* n);

/* Per-phase device time of the last compute in milliseconds (CUDA events on the handle's
 * stream): [0] edge build, [1] network forward+backward, [2] finalize/store. */
ALG_API int alg_get_timings(alg_handle* h, double* ms3);

/* Counters of the last compute (doubles): what = "step" -> [own kernel launches, edges, chunks,
 * tiles]; with option profile=1 also "kernel_ms" / "kernel_launches" -> per kernel family
 * [F0, FK, T, BK, B0, fixup, fused] summed CUDA-event durations (ms) and launch counts; "pipeline" -> [1 if the
 * last step ran the fused kernel, CTAs in the fused grid, 1 once the tiled fallback became sticky, 1 if the last chunked
 * step used the device-built plan]; "host_ms" -> wall times of the last alg_compute_host call in ms [neighbour-list
 * flatten + upload issue (0 when the device copy was reused), whole call, pinning of the caller's arrays]. */
ALG_API int alg_get_stats(alg_handle* h, const char* what, double* out, int n);

/* ---- ghost halo exchange of spatial-domain multi-GPU runs (one rank per GPU) -------------------------------------------
 * Replaces what LAMMPS' Comm does around the reference pair style: comm->forward_comm() (ghost x <- owner x + image
 * shift) before compute() and comm->reverse_comm() (owner f += ghost f) after it -- required because Allegro runs with
 * `newton on` (pair_nequip_allegro.cpp:149; same owner-accumulation protocol as compute/compute_allegro.cpp:159-189).
 * Transport: grouped ncclSend/ncclRecv to all neighbouring domains at once over NVLink; periodic self-images of a rank stay
 * on the device.  The reverse unpack is a sorted segmented sum (fixed order), so forces are bit-reproducible.
 * NCCL is bound at run time (libnccl.so.2); alg_comm_create(nranks = 1) needs no NCCL at all.
 *
 * alg_comm_unique_id : rank 0 creates the 128-byte NCCL id and distributes it by any out-of-band means (MPI_Bcast in
 *                      LAMMPS, torch.distributed in bench.py).
 * alg_comm_set_plan  : per peer p (peer_rank[p] may be the rank itself = periodic self-images):
 *                      send_index[p][0..send_count[p])  local atoms whose positions peer p needs as ghosts (host ints),
 *                      send_shift[p]                    [send_count][3] image shift added to those positions, or NULL,
 *                      recv_begin[p], recv_count[p]     the contiguous ghost slice [begin, begin+count) of x / f that
 *                                                       holds peer p's atoms, in the order of p's send_index for this rank.
 * alg_comm_forward / alg_comm_reverse : device pointers x / f of [nlocal+nghost][3] doubles; stream-ordered, no host sync.
 * alg_comm_allreduce_sum : in-place sum over ranks of n <= 64 host doubles (eng_vdwl, virial: LAMMPS' MPI_Allreduce). */
typedef struct alg_comm alg_comm;
ALG_API int alg_comm_unique_id(char* id128);
ALG_API int alg_comm_create(int cuda_device, int nranks, int rank, const char* id128, alg_comm** out);
ALG_API void alg_comm_destroy(alg_comm* c);
ALG_API const char* alg_comm_last_error(const alg_comm* c);
ALG_API int alg_comm_set_plan(alg_comm* c, int npeer, const int* peer_rank, const int* send_count, const int* const* send_index,
                      const double* const* send_shift, const int* recv_begin, const int* recv_count);
ALG_API int alg_comm_forward(alg_comm* c, double* d_x, void* stream);
ALG_API int alg_comm_reverse(alg_comm* c, double* d_f, void* stream);
ALG_API int alg_comm_allreduce_sum(alg_comm* c, double* values, int n, void* stream);
/* [bytes sent to other ranks per forward, per reverse, packed entries, distinct owner atoms] */
ALG_API int alg_comm_stats(const alg_comm* c, double* out4);

/* ---- neighbour-list build on the device (the step on the CALLER's side of the path) ------------------------------------
 * What LAMMPS' Neighbor class builds before Pair::compute reads list->ilist / numneigh / firstneigh
 * (pair_nequip_allegro.cpp:340-350, 469-480; requested with neighbor->add_request(this, REQ_FULL), :142-147): for every
 * LOCAL atom i all atoms j != i (locals + ghosts) with |x_i - x_j|^2 <= rneigh^2 (rneigh = r_max + skin), as the 2-D view
 * d_neighbors[i*stride_i + jj*stride_jj] + d_numneigh that alg_compute_device consumes.  Cell list (cells of rneigh inside
 * the bounding box lo..hi of ALL atoms incl. ghosts), atoms sorted by (cell, index) with a stable radix sort, one warp per
 * atom: the same positions always give the same list in the same order.  `max_count` (may be NULL) returns the largest
 * neighbour count (one small synchronisation); a count above max_neigh truncates the rows and returns ALG_ESTATE.
 * alg_neigh_check: Verlet-skin criterion -- *rebuild = 1 when some atom moved further than skin/2 since the last build. */
typedef struct alg_neigh alg_neigh;
ALG_API int alg_neigh_create(int cuda_device, alg_neigh** out);
ALG_API void alg_neigh_destroy(alg_neigh* n);
ALG_API const char* alg_neigh_last_error(const alg_neigh* n);
ALG_API int alg_neigh_build(alg_neigh* n, int nlocal, int nghost, const double* d_x, const double* lo, const double* hi, double rneigh,
                    int max_neigh, int64_t stride_i, int64_t stride_jj, int* d_neighbors, int* d_numneigh, int* max_count, void* stream);
ALG_API int alg_neigh_check(alg_neigh* n, int ntot, const double* d_x, double skin, int* rebuild, void* stream);

/* Building blocks of the above for callers that bring their own transport (LAMMPS pack/unpack_forward/reverse_comm):
 *   pack:        buf[k][0..2] = x[list[k]][0..2] + shift[k][0..2]   (d_shift may be NULL)
 *   unpack_add:  f[list[k]][0..2] += buf[k][0..2]   (fp64 atomics: list entries may repeat, and then the summation
 *                order -- hence the last bits -- is not fixed; alg_comm_reverse is the deterministic path) */
ALG_API int alg_halo_pack(const double* d_x, const int* d_list, int n, const double* d_shift, double* d_buf,
                  void* stream);
ALG_API int alg_halo_unpack_add(double* d_f, const int* d_list, int n, const double* d_buf, void* stream);

/* Number of CUDA devices visible to this process (torch::cuda::device_count() in the reference's device
 * selection, pair_nequip_allegro.cpp:102-118: rank >= count is an error, or wraps around in debug mode); returns 0
 * when no device is usable (then alg_create fails with ALG_ECUDA -- there is no CPU fallback). */
ALG_API int alg_device_count(void);

/* library / build identification, e.g. "allegro_b200 0.1.0 sm_100a" */
ALG_API const char* alg_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ALLEGRO_B200_H */
