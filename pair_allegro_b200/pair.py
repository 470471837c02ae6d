"""Python mirror of the reference pair style's interface on top of the C-ABI.

`PairAllegroB200` has the methods LAMMPS calls on `PairNequIPAllegro<false>`
(/root/reference/pair_nequip_allegro.h:43-50): settings, coeff, init_style, init_one,
compute -- same argument meaning and the same error behaviour (messages follow
pair_nequip_allegro.cpp) so that the parity tests read like the reference's own.  The C++
pair style for a real LAMMPS build is src/pair_allegro_b200.cpp; both bind the same C-ABI.

`atom` / `list` arguments are duck-typed LAMMPS stand-ins with the fields the reference
reads: atom.x,type,tag,nlocal,nghost,ntypes,f ; list.inum,gnum,ilist,numneigh and the
neighbour storage as (neigh_flat, first) -- see lmpshim/harness.py for the test harness.
"""
import os

import numpy as np

from . import capi


class PairAllegroB200:
    def __init__(self, device=0, debug_mode=None, pin_host=False):
        # pair_nequip_allegro.cpp:66-125
        # pin_host: let the library cudaHostRegister the x / f / type arrays it is handed (what the C++ pair style does
        # for LAMMPS' long-lived atom arrays).  Off by default here: numpy arrays of a test may be freed and re-allocated
        # at the same address between two calls, which a cached registration cannot see.
        self.pin_host = pin_host
        self.restartinfo = 0
        self.manybody_flag = 1
        self.device = device
        self.debug_mode = (os.environ.get("_NEQUIP_LOG_LEVEL") == "DEBUG") if debug_mode is None else debug_mode
        self.allocated = False
        self.handle = None
        self.custom_output_names = []
        self.custom_output = {}
        self.eng_vdwl = 0.0
        self.virial = np.zeros(6)
        self.eatom = None
        self.debug_lines = []
        self._neigh_ago = 0

    # pair_nequip_allegro.cpp:168-172
    def settings(self, args):
        if len(args) > 0:
            raise RuntimeError("Illegal pair_style command, too many arguments")

    @staticmethod
    def resolve_weight_path(model_path):
        """`pair_coeff * * <model> ...` keeps the reference syntax (cpp:195-206): a `.nequip.pth`
        / `.nequip.pt2` path is accepted and resolved to the `.alg` file the offline exporter wrote
        next to it; a `.alg` path is taken as is."""
        if model_path.endswith(".alg"):
            return model_path
        for ext in (".nequip.pth", ".nequip.pt2"):
            if model_path.endswith(ext):
                cand = model_path[:-len(ext)] + ".alg"
                if os.path.exists(cand):
                    return cand
                if ext == ".nequip.pt2":
                    raise RuntimeError("no exported weights %s for %s: an AOT-Inductor package cannot be converted; export the "
                                       "weights of the same model to %s (python -m pair_allegro_b200.export <model>.nequip.pth %s)"
                                       % (cand, model_path, cand, cand))
                raise RuntimeError("no exported weights %s for %s: run `python -m pair_allegro_b200.export %s %s` "
                                   "(supports TorchScript files carrying an allegro_b200_config entry)"
                                   % (cand, model_path, model_path, cand))
        raise RuntimeError("Only accepts model paths with extension `.nequip.pth`, `.nequip.pt2` or `.alg`, but found" + model_path)

    # pair_nequip_allegro.cpp:174-330
    def coeff(self, args, ntypes):
        self.ntypes = ntypes
        self.setflag = np.zeros((ntypes + 1, ntypes + 1), dtype=np.int32)
        self.cutoff_matrix = np.zeros((ntypes, ntypes))
        self.allocated = True
        if len(args) != 3 + ntypes:
            raise RuntimeError("Incorrect args for pair coefficients, should be * * <model>.nequip.pth/pt2 <type1> <type2> ... <typen>")
        if args[0] != "*" or args[1] != "*":
            raise RuntimeError("Incorrect args for pair coefficients")
        self.model_path = args[2]
        self.handle = capi.Handle(self.resolve_weight_path(self.model_path), self.device)
        md = self.handle.metadata()
        self.metadata = md
        self.cutoff = md["r_max"]
        self.type_mapper = [-1] * ntypes
        for i, ele in enumerate(md["type_names"][:md["num_types"]]):
            for itype in range(1, ntypes + 1):
                if ele == args[itype + 3 - 1]:
                    self.type_mapper[itype - 1] = i
        for i in range(1, ntypes + 1):
            for j in range(i, ntypes + 1):
                if self.type_mapper[i - 1] >= 0 and self.type_mapper[j - 1] >= 0:
                    self.setflag[i][j] = 1
        if md["per_edge_type_cutoff"] is not None:
            rev = [-1] * md["num_types"]
            for i in range(ntypes):
                if self.type_mapper[i] >= 0:       # guards the reference's UB at cpp:308
                    rev[self.type_mapper[i]] = i
            for i in range(md["num_types"]):
                for j in range(md["num_types"]):
                    if rev[i] >= 0 and rev[j] >= 0:
                        self.cutoff_matrix[rev[i]][rev[j]] = md["per_edge_type_cutoff"][i][j]
        else:
            self.cutoff_matrix[:, :] = self.cutoff
        self.handle.set_type_map(self.type_mapper, self.cutoff_matrix)
        self.handle.set_option("host_register", "1" if self.pin_host else "0")
        if self.debug_mode:
            self.handle.set_option("keep_edges", "1")

    # pair_nequip_allegro.cpp:137-151
    def init_style(self, newton_pair=1, tag_enable=1):
        if tag_enable == 0:
            raise RuntimeError("Pair style Allegro requires atom IDs")
        if newton_pair == 0:
            raise RuntimeError("Pair style allegro requires newton pair on")

    # pair_nequip_allegro.cpp:153-156
    def init_one(self, i, j):
        return self.cutoff

    # pair_nequip_allegro.cpp:333-407
    def compute(self, atom, lst, eflag=1, vflag=1, eflag_atom=1, vflag_atom=0, neigh_ago=0):
        """neigh_ago = LAMMPS' neighbor->ago (0 on a rebuild step): > 0 reuses the device copy of the list"""
        if lst.inum == 0:
            return
        if neigh_ago != self._neigh_ago:
            self.handle.set_option("neigh_ago", str(int(neigh_ago)))
            self._neigh_ago = neigh_ago
        if vflag_atom:
            raise RuntimeError("Pair styles nequip and allegro do not support per-atom virial")
        ntot = lst.inum + lst.gnum
        if eflag_atom:
            self.eatom = np.zeros(ntot)
        eng, vir = self.handle.compute_host(atom.x, atom.type, lst.ilist[:lst.inum], lst.numneigh, lst.neigh_flat, lst.first,
                                            lst.inum, lst.gnum, atom.f, self.eatom if eflag_atom else None, bool(vflag))
        self.eng_vdwl = eng
        if vflag:
            self.virial[:] = vir
        if self.debug_mode:   # the reference's edge dump (cpp:562-565, 620-633): tag-1 indices and |x_i - x_j|
            e = self.handle.get_edges()
            d = np.linalg.norm(atom.x[e[0]] - atom.x[e[1]], axis=1)
            self.debug_lines = ["Allegro edges: i j rij"] + ["%d %d %.10g" % (atom.tag[a] - 1, atom.tag[b] - 1, r)
                                                             for a, b, r in zip(e[0], e[1], d)] + ["end Allegro edges"]
        for name in self.custom_output_names:
            self.custom_output[name] = self.handle.get_output(name)

    # pair_nequip_allegro.cpp:681-684
    def add_custom_output(self, name):
        self.custom_output_names.append(name)
