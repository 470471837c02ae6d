"""ctypes driver of the lmpshim C API (lmpshim/shim.cpp): runs ONE LAMMPS pair style compiled
against the shim headers -- either the reference's unmodified PairNequIPAllegro<false>
(oracle/_ref/libref_pair_allegro.so) or this repo's PairAllegroB200
(src/libpair_allegro_b200_shim.so) -- through the calls LAMMPS makes:
settings / coeff / init_style / init_one / compute(eflag, vflag)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_pair_allegro.so")
OURS_LIB = os.path.join(ROOT, "src", "libpair_allegro_b200_shim.so")
OURS_KK_LIB = os.path.join(ROOT, "src", "libpair_allegro_b200_kk_shim.so")


class ShimError(RuntimeError):
    pass


def _load(path):
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    vp = C.c_void_p
    lib.shim_create.restype = vp
    lib.shim_create.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, vp]
    lib.shim_destroy.argtypes = [vp]
    lib.shim_last_error.restype = C.c_char_p
    lib.shim_last_error.argtypes = [vp]
    lib.shim_set_list.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp]
    lib.shim_set_positions.argtypes = [vp, vp]
    lib.shim_zero_forces.argtypes = [vp]
    lib.shim_set_newton.argtypes = [vp, C.c_int]
    for fn in ("shim_pair_create", "shim_pair_init_style", "shim_neigh_request_flags"):
        getattr(lib, fn).argtypes = [vp]
        getattr(lib, fn).restype = C.c_int
    lib.shim_pair_settings.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p)]
    lib.shim_pair_coeff.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p)]
    lib.shim_pair_init_one.argtypes = [vp, C.c_int, C.c_int]
    lib.shim_pair_init_one.restype = C.c_double
    lib.shim_pair_flags.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.shim_pair_compute.argtypes = [vp, C.c_int, C.c_int]
    lib.shim_pair_compute.restype = C.c_int
    lib.shim_get_forces.argtypes = [vp, vp]
    lib.shim_get_eng.argtypes = [vp]
    lib.shim_get_eng.restype = C.c_double
    lib.shim_get_virial.argtypes = [vp, vp]
    lib.shim_get_eatom.argtypes = [vp, vp]
    lib.shim_get_eatom.restype = C.c_int
    lib.shim_set_ghost_owner.argtypes = [vp, vp]
    lib.shim_set_timestep.argtypes = [vp, C.c_longlong]
    lib.shim_set_neigh_ago.argtypes = [vp, C.c_int]
    if hasattr(lib, "shim_compute_create"):
        lib.shim_compute_create.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p)]
        lib.shim_compute_create.restype = C.c_int
        lib.shim_compute_vector.argtypes = [vp, C.c_int, vp, C.c_int]
        lib.shim_compute_vector.restype = C.c_int
        lib.shim_compute_peratom.argtypes = [vp, C.c_int, vp, C.c_int]
        lib.shim_compute_peratom.restype = C.c_int
    return lib


class ShimLammps:
    """one 'LAMMPS instance' holding atoms + a full neighbour list + one pair style"""

    def __init__(self, lib_path, atom, lst):
        if not os.path.exists(lib_path):
            raise FileNotFoundError(lib_path)
        self.lib = _load(lib_path)
        x = np.ascontiguousarray(atom.x, dtype=np.float64)
        t = np.ascontiguousarray(atom.type, dtype=np.int32)
        tag = np.ascontiguousarray(atom.tag, dtype=np.int64)
        self.ntot = atom.nlocal + atom.nghost
        self.nlocal = atom.nlocal
        self.h = self.lib.shim_create(atom.ntypes, atom.nlocal, atom.nghost, x.ctypes.data, t.ctypes.data, tag.ctypes.data)
        il = np.ascontiguousarray(lst.ilist[:self.ntot], dtype=np.int32)
        nn = np.ascontiguousarray(lst.numneigh[:self.ntot], dtype=np.int32)
        nf = np.ascontiguousarray(lst.neigh_flat, dtype=np.int32)
        fi = np.ascontiguousarray(lst.first[:self.ntot], dtype=np.int64)
        self.lib.shim_set_list(self.h, lst.inum, lst.gnum, il.ctypes.data, nn.ctypes.data, nf.ctypes.data, fi.ctypes.data)
        self._ck(self.lib.shim_pair_create(self.h))

    def _ck(self, rc):
        if rc != 0:
            raise ShimError(self.lib.shim_last_error(self.h).decode())

    @staticmethod
    def _argv(args):
        arr = (C.c_char_p * max(1, len(args)))()
        for i, a in enumerate(args):
            arr[i] = a.encode()
        return arr

    def pair_style(self, args=()):
        self._ck(self.lib.shim_pair_settings(self.h, len(args), self._argv(args)))

    def pair_coeff(self, args):
        self._ck(self.lib.shim_pair_coeff(self.h, len(args), self._argv(args)))

    def init(self, newton_pair=1):
        self.lib.shim_set_newton(self.h, newton_pair)
        self._ck(self.lib.shim_pair_init_style(self.h))

    def init_one(self, i, j):
        return self.lib.shim_pair_init_one(self.h, i, j)

    def flags(self):
        a, b = C.c_int(), C.c_int()
        self.lib.shim_pair_flags(self.h, C.byref(a), C.byref(b))
        return dict(restartinfo=a.value, manybody_flag=b.value, neigh_request=self.lib.shim_neigh_request_flags(self.h))

    def set_positions(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.lib.shim_set_positions(self.h, x.ctypes.data)

    def compute(self, eflag=3, vflag=1, zero=True, neigh_ago=0):
        """eflag: 1 global energy | 2 per-atom ; vflag: 1 global virial | 4 per-atom (LAMMPS bit flags);
        neigh_ago = neighbor->ago (0 = the list was rebuilt this step)"""
        self.lib.shim_set_neigh_ago(self.h, neigh_ago)
        if zero:
            self.lib.shim_zero_forces(self.h)
        self._ck(self.lib.shim_pair_compute(self.h, eflag, vflag))
        f = np.zeros((self.ntot, 3))
        self.lib.shim_get_forces(self.h, f.ctypes.data)
        vir = np.zeros(6)
        self.lib.shim_get_virial(self.h, vir.ctypes.data)
        eatom = np.zeros(self.ntot)
        has = self.lib.shim_get_eatom(self.h, eatom.ctypes.data) == 0
        return dict(f=f, eng_vdwl=self.lib.shim_get_eng(self.h), virial=vir, eatom=eatom if has else None)

    # ---- compute allegro / compute allegro/atom (shim.cpp: shim_compute_*) ----------------------------
    def set_ghost_owner(self, owner_of_ghosts):
        """[nghost] local index each ghost is an image of: lets comm->reverse_comm(compute) fold ghost rows"""
        o = np.ascontiguousarray(owner_of_ghosts, dtype=np.int32)
        assert len(o) == self.ntot - self.nlocal
        self.lib.shim_set_ghost_owner(self.h, o.ctypes.data)

    def compute_create(self, words):
        """words = the LAMMPS command after `compute`, e.g. ["ae", "all", "allegro/atom", "atomic_energy", "1", "0"]"""
        idx = self.lib.shim_compute_create(self.h, len(words), self._argv(words))
        if idx < 0:
            raise ShimError(self.lib.shim_last_error(self.h).decode())
        return idx

    def compute_vector(self, idx, n):
        out = np.zeros(n)
        self._ck(self.lib.shim_compute_vector(self.h, idx, out.ctypes.data, n))
        return out

    def compute_peratom(self, idx, ncols):
        out = np.zeros((self.nlocal, ncols))
        self._ck(self.lib.shim_compute_peratom(self.h, idx, out.ctypes.data, ncols))
        return out

    def close(self):
        if self.h:
            self.lib.shim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShimLammpsKK:
    """one 'LAMMPS + KOKKOS instance': device-resident atoms + the 2-D device neighbour view + pair_style allegro/kk
    (src/pair_allegro_b200_kokkos.cpp compiled against lmpshim/kokkos_shim.h; lmpshim/shim_kk.cpp)"""
    FULL, HALFTHREAD, HALF = 1, 2, 4

    def __init__(self, atom, lst, layout_left=True, neighflag=4, lib_path=OURS_KK_LIB):
        if not os.path.exists(lib_path):
            raise FileNotFoundError(lib_path)
        lib = C.CDLL(lib_path, mode=C.RTLD_LOCAL)
        vp = C.c_void_p
        lib.shimkk_create.restype = vp
        lib.shimkk_create.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_int]
        lib.shimkk_destroy.argtypes = [vp]
        lib.shimkk_last_error.restype = C.c_char_p
        lib.shimkk_last_error.argtypes = [vp]
        lib.shimkk_set_list.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int]
        lib.shimkk_set_newton.argtypes = [vp, C.c_int]
        for fn in ("shimkk_pair_create", "shimkk_pair_init_style"):
            getattr(lib, fn).argtypes = [vp]
            getattr(lib, fn).restype = C.c_int
        lib.shimkk_pair_settings.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p)]
        lib.shimkk_pair_coeff.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p)]
        lib.shimkk_pair_compute.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        lib.shimkk_pair_compute.restype = C.c_int
        lib.shimkk_get_forces.argtypes = [vp, vp]
        lib.shimkk_get_eng.argtypes = [vp]
        lib.shimkk_get_eng.restype = C.c_double
        lib.shimkk_get_virial.argtypes = [vp, vp]
        lib.shimkk_get_eatom.argtypes = [vp, vp]
        lib.shimkk_get_eatom.restype = C.c_int
        self.lib = lib
        x = np.ascontiguousarray(atom.x, dtype=np.float64)
        t = np.ascontiguousarray(atom.type, dtype=np.int32)
        tag = np.ascontiguousarray(atom.tag, dtype=np.int64)
        self.ntot, self.nlocal = atom.nlocal + atom.nghost, atom.nlocal
        self.h = lib.shimkk_create(atom.ntypes, atom.nlocal, atom.nghost, x.ctypes.data, t.ctypes.data, tag.ctypes.data, neighflag)
        il = np.ascontiguousarray(lst.ilist[:self.ntot], dtype=np.int32)
        nn = np.ascontiguousarray(lst.numneigh[:self.ntot], dtype=np.int32)
        nf = np.ascontiguousarray(lst.neigh_flat, dtype=np.int32)
        fi = np.ascontiguousarray(lst.first[:self.ntot], dtype=np.int64)
        lib.shimkk_set_list(self.h, lst.inum, lst.gnum, il.ctypes.data, nn.ctypes.data, nf.ctypes.data, fi.ctypes.data, 1 if layout_left else 0)
        self._ck(lib.shimkk_pair_create(self.h))

    def _ck(self, rc):
        if rc != 0:
            raise ShimError(self.lib.shimkk_last_error(self.h).decode())

    def pair_style(self, args=()):
        self._ck(self.lib.shimkk_pair_settings(self.h, len(args), ShimLammps._argv(args)))

    def pair_coeff(self, args):
        self._ck(self.lib.shimkk_pair_coeff(self.h, len(args), ShimLammps._argv(args)))

    def init(self, newton_pair=1):
        self.lib.shimkk_set_newton(self.h, newton_pair)
        self._ck(self.lib.shimkk_pair_init_style(self.h))

    def compute(self, eflag=3, vflag=1, zero=True):
        self._ck(self.lib.shimkk_pair_compute(self.h, eflag, vflag, 1 if zero else 0))
        f = np.zeros((self.ntot, 3))
        self.lib.shimkk_get_forces(self.h, f.ctypes.data)
        vir = np.zeros(6)
        self.lib.shimkk_get_virial(self.h, vir.ctypes.data)
        eatom = np.zeros(self.ntot)
        has = self.lib.shimkk_get_eatom(self.h, eatom.ctypes.data) == 0
        return dict(f=f, eng_vdwl=self.lib.shimkk_get_eng(self.h), virial=vir, eatom=eatom if has else None)

    def close(self):
        if self.h:
            self.lib.shimkk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
