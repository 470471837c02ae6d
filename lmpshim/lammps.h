// Minimal LAMMPS-compatible declarations ("lmpshim"): just enough of the LAMMPS core API for
// a pair style to compile and run outside LAMMPS -- the reference's
// /root/reference/pair_nequip_allegro.cpp (unmodified, for oracle/_ref) and this repo's
// src/pair_allegro_b200.cpp.  Written from the public LAMMPS developer documentation; it is a
// test harness standing in for LAMMPS core (out of scope), not a copy of it.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <stdexcept>
#include <string>

#include "mpi.h"

#define FLERR __FILE__, __LINE__
#define NEIGHMASK 0x1FFFFFFF

namespace LAMMPS_NS {

typedef int64_t tagint;
typedef int64_t bigint;

class LAMMPSException : public std::runtime_error {
 public:
  explicit LAMMPSException(const std::string& m) : std::runtime_error(m) {}
};

class LAMMPS;
class Pair;
class Compute;

class Memory {
 public:
  template <typename T> T** create(T**& array, int n1, int n2, const char* /*name*/) {
    T* data = (T*)calloc((size_t)n1 * n2 > 0 ? (size_t)n1 * n2 : 1, sizeof(T));
    array = (T**)malloc(sizeof(T*) * (n1 > 0 ? n1 : 1));
    for (int i = 0; i < n1; ++i) array[i] = data + (size_t)i * n2;
    return array;
  }
  template <typename T> T* create(T*& array, int n, const char* /*name*/) {
    array = (T*)calloc(n > 0 ? n : 1, sizeof(T));
    return array;
  }
  template <typename T> void destroy(T**& array) {
    if (!array) return;
    free(array[0]);
    free(array);
    array = nullptr;
  }
  template <typename T> void destroy(T*& array) {
    free(array);
    array = nullptr;
  }
};

class Error {
 public:
  static void fmt_into(std::ostringstream& os, const std::string& f, size_t pos) { os << f.substr(pos); }
  template <typename A, typename... R>
  static void fmt_into(std::ostringstream& os, const std::string& f, size_t pos, const A& a, const R&... r) {
    size_t p = f.find("{}", pos);
    if (p == std::string::npos) { os << f.substr(pos); return; }
    os << f.substr(pos, p - pos) << a;
    fmt_into(os, f, p + 2, r...);
  }
  template <typename... Args> [[noreturn]] void all(const char* file, int line, const std::string& f, const Args&... args) {
    std::ostringstream os;
    os << "ERROR: ";
    fmt_into(os, f, 0, args...);
    os << " (" << file << ":" << line << ")";
    throw LAMMPSException(os.str());
  }
  template <typename... Args> [[noreturn]] void one(const char* file, int line, const std::string& f, const Args&... args) {
    all(file, line, f, args...);
  }
  template <typename... Args> void message(const char*, int, const std::string& f, const Args&... args) {
    std::ostringstream os;
    fmt_into(os, f, 0, args...);
    fprintf(stderr, "%s\n", os.str().c_str());
  }
  template <typename... Args> void warning(const char* a, int b, const std::string& f, const Args&... args) { message(a, b, f, args...); }
};

class Atom {
 public:
  int tag_enable = 1;
  int ntypes = 0;
  int nlocal = 0, nghost = 0, nmax = 0;
  double** x = nullptr;
  double** f = nullptr;
  int* type = nullptr;
  tagint* tag = nullptr;
};

class Comm {
 public:
  int me = 0, nprocs = 1;
  Atom* atom = nullptr;          // harness wiring for reverse_comm(Compute*) (compute.h)
  int* ghost_owner = nullptr;    // [nghost] local index of the atom each ghost is an image of
  inline void reverse_comm(Compute*);
  void reverse_comm() {}
  void forward_comm() {}
};

class Domain {
 public:
  double boxlo[3] = {0, 0, 0}, boxhi[3] = {1, 1, 1};
  double xy = 0, xz = 0, yz = 0;
};

class Force {
 public:
  int newton_pair = 1;
  Pair* pair = nullptr;
};

class Update {
 public:
  bigint ntimestep = 0;
};
class Output {};

namespace NeighConst {
enum { REQ_DEFAULT = 0, REQ_FULL = 1 << 0, REQ_GHOST = 1 << 1, REQ_SIZE = 1 << 2 };
}

class NeighRequest {
 public:
  int flags = 0;
  void set_kokkos_host(int) {}
  void set_kokkos_device(int) {}
};

class NeighList {
 public:
  int inum = 0, gnum = 0;
  int* ilist = nullptr;
  int* numneigh = nullptr;
  int** firstneigh = nullptr;
};

class Neighbor {
 public:
  NeighRequest last_request;
  int nrequest = 0;
  int ago = 0;
  NeighRequest* add_request(Pair*, int flags = 0) {
    last_request.flags = flags;
    ++nrequest;
    return &last_request;
  }
  NeighRequest* find_request(Pair*) { return &last_request; }
};

class LAMMPS {
 public:
  Memory* memory;
  Error* error;
  Atom* atom;
  Comm* comm;
  Domain* domain;
  Force* force;
  Neighbor* neighbor;
  Update* update;
  Output* output;
  void* kokkos = nullptr;
  MPI_Comm world = MPI_COMM_WORLD;
  LAMMPS()
      : memory(new Memory), error(new Error), atom(new Atom), comm(new Comm), domain(new Domain), force(new Force),
        neighbor(new Neighbor), update(new Update), output(new Output) { comm->atom = atom; }
  ~LAMMPS() {
    delete memory; delete error; delete atom; delete comm; delete domain; delete force; delete neighbor; delete update; delete output;
  }
};

class Pointers {
 public:
  explicit Pointers(LAMMPS* ptr)
      : lmp(ptr), memory(ptr->memory), error(ptr->error), atom(ptr->atom), comm(ptr->comm), domain(ptr->domain), force(ptr->force),
        neighbor(ptr->neighbor), update(ptr->update), output(ptr->output), world(ptr->world) {}
  virtual ~Pointers() = default;

 protected:
  LAMMPS* lmp;
  Memory*& memory;
  Error*& error;
  Atom*& atom;
  Comm*& comm;
  Domain*& domain;
  Force*& force;
  Neighbor*& neighbor;
  Update*& update;
  Output*& output;
  MPI_Comm& world;
};

}  // namespace LAMMPS_NS
