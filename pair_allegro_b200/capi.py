"""ctypes binding of the C-ABI (include/allegro_b200.h).

This is the reference-side binding a maintainer would write for a Python driver; the LAMMPS
pair style (src/pair_allegro_b200.cpp) binds the same symbols from C++.  There is NO fallback:
if the shared library is missing or no CUDA device is usable, an exception is raised.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liballegro_b200.so")
DEBUG_LIB_PATH = os.path.join(_HERE, "liballegro_b200_debug.so")      # tcgen05 primitive unit tests / micro-benchmarks (not the product)

EXPORTS = ["alg_create", "alg_destroy", "alg_last_error", "alg_metadata", "alg_set_type_map", "alg_set_option",
           "alg_compute_host", "alg_compute_device", "alg_get_edges", "alg_get_output", "alg_get_timings", "alg_get_stats",
           "alg_halo_pack", "alg_halo_unpack_add", "alg_version", "alg_device_count",
           "alg_comm_unique_id", "alg_comm_create", "alg_comm_destroy", "alg_comm_last_error", "alg_comm_set_plan", "alg_comm_forward",
           "alg_comm_reverse", "alg_comm_allreduce_sum", "alg_comm_stats",
           "alg_neigh_create", "alg_neigh_destroy", "alg_neigh_last_error", "alg_neigh_build", "alg_neigh_check"]

_lib = None


class AllegroError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("allegro_b200 error %d: %s" % (code, msg))
        self.code = code


def load_library(path=None):
    """dlopen liballegro_b200.so and declare prototypes.  Raises if the library is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C pair_allegro_b200/csrc`); there is no CPU fallback" % p)
    lib = C.CDLL(p)
    vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
    lib.alg_create.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    lib.alg_create.restype = C.c_int
    lib.alg_destroy.argtypes = [vp]
    lib.alg_destroy.restype = None
    lib.alg_last_error.argtypes = [vp]
    lib.alg_last_error.restype = C.c_char_p
    lib.alg_metadata.argtypes = [vp, dp, ip, C.POINTER(C.c_char_p), C.POINTER(dp), ip]
    lib.alg_metadata.restype = C.c_int
    lib.alg_set_type_map.argtypes = [vp, C.c_int, ip, dp]
    lib.alg_set_type_map.restype = C.c_int
    lib.alg_set_option.argtypes = [vp, C.c_char_p, C.c_char_p]
    lib.alg_set_option.restype = C.c_int
    lib.alg_compute_host.argtypes = [vp, C.c_int, C.c_int, dp, ip, ip, ip, C.POINTER(ip), C.c_int, C.c_int, dp, dp, dp, dp]
    lib.alg_compute_host.restype = C.c_int
    lib.alg_compute_device.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                       vp, vp, dp, dp, vp]
    lib.alg_compute_device.restype = C.c_int
    lib.alg_get_edges.argtypes = [vp, C.POINTER(C.POINTER(C.c_int64)), C.POINTER(C.c_int64)]
    lib.alg_get_edges.restype = C.c_int
    lib.alg_get_output.argtypes = [vp, C.c_char_p, C.POINTER(dp), C.POINTER(C.c_int64)]
    lib.alg_get_output.restype = C.c_int
    lib.alg_get_timings.argtypes = [vp, dp]
    lib.alg_get_timings.restype = C.c_int
    lib.alg_get_stats.argtypes = [vp, C.c_char_p, dp, C.c_int]
    lib.alg_get_stats.restype = C.c_int
    lib.alg_halo_pack.argtypes = [vp, vp, C.c_int, vp, vp, vp]
    lib.alg_halo_pack.restype = C.c_int
    lib.alg_halo_unpack_add.argtypes = [vp, vp, C.c_int, vp, vp]
    lib.alg_halo_unpack_add.restype = C.c_int
    lib.alg_comm_unique_id.argtypes = [C.c_char_p]
    lib.alg_comm_unique_id.restype = C.c_int
    lib.alg_comm_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(vp)]
    lib.alg_comm_create.restype = C.c_int
    lib.alg_comm_destroy.argtypes = [vp]
    lib.alg_comm_destroy.restype = None
    lib.alg_comm_last_error.argtypes = [vp]
    lib.alg_comm_last_error.restype = C.c_char_p
    lib.alg_comm_set_plan.argtypes = [vp, C.c_int, ip, ip, C.POINTER(ip), C.POINTER(dp), ip, ip]
    lib.alg_comm_set_plan.restype = C.c_int
    lib.alg_comm_forward.argtypes = [vp, vp, vp]
    lib.alg_comm_forward.restype = C.c_int
    lib.alg_comm_reverse.argtypes = [vp, vp, vp]
    lib.alg_comm_reverse.restype = C.c_int
    lib.alg_comm_allreduce_sum.argtypes = [vp, dp, C.c_int, vp]
    lib.alg_comm_allreduce_sum.restype = C.c_int
    lib.alg_comm_stats.argtypes = [vp, dp]
    lib.alg_comm_stats.restype = C.c_int
    lib.alg_neigh_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.alg_neigh_create.restype = C.c_int
    lib.alg_neigh_destroy.argtypes = [vp]
    lib.alg_neigh_destroy.restype = None
    lib.alg_neigh_last_error.argtypes = [vp]
    lib.alg_neigh_last_error.restype = C.c_char_p
    lib.alg_neigh_build.argtypes = [vp, C.c_int, C.c_int, vp, dp, dp, C.c_double, C.c_int, C.c_int64, C.c_int64, vp, vp, ip, vp]
    lib.alg_neigh_build.restype = C.c_int
    lib.alg_neigh_check.argtypes = [vp, C.c_int, vp, C.c_double, ip, vp]
    lib.alg_neigh_check.restype = C.c_int
    lib.alg_device_count.argtypes = []
    lib.alg_device_count.restype = C.c_int
    lib.alg_version.argtypes = []
    lib.alg_version.restype = C.c_char_p
    if path is None:
        _lib = lib
    return lib


def load_debug_library():
    """the separate library with the tcgen05 unit-test kernels (csrc/alg_debug.cu)"""
    if not os.path.exists(DEBUG_LIB_PATH):
        raise FileNotFoundError("%s not found: build it with `make -C pair_allegro_b200/csrc`" % DEBUG_LIB_PATH)
    return C.CDLL(DEBUG_LIB_PATH)


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Handle:
    """thin RAII wrapper over alg_handle*"""

    def __init__(self, weight_path, device=0):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.alg_create(os.fsencode(weight_path), int(device), C.byref(self.h))
        if rc != 0:
            msg = self.lib.alg_last_error(None).decode()
            self.h = None
            raise AllegroError(rc, msg)

    def close(self):
        if getattr(self, "h", None):
            self.lib.alg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise AllegroError(rc, self.lib.alg_last_error(self.h).decode())

    def metadata(self):
        r = C.c_double()
        n = C.c_int()
        names = C.c_char_p()
        pc = C.POINTER(C.c_double)()
        tf = C.c_int()
        self._check(self.lib.alg_metadata(self.h, C.byref(r), C.byref(n), C.byref(names), C.byref(pc), C.byref(tf)))
        T = n.value
        cut = None
        if pc:
            cut = np.array([pc[i] for i in range(T * T)]).reshape(T, T)
        return dict(r_max=r.value, num_types=T, type_names=names.value.decode().split(), per_edge_type_cutoff=cut,
                    allow_tf32=bool(tf.value))

    def set_type_map(self, type_mapper, cutoff_matrix):
        tm = np.ascontiguousarray(type_mapper, dtype=np.int32)
        cm = np.ascontiguousarray(cutoff_matrix, dtype=np.float64)
        assert cm.shape == (len(tm), len(tm))
        self._check(self.lib.alg_set_type_map(self.h, len(tm), _iptr(tm), _dptr(cm)))

    def set_option(self, key, value):
        self._check(self.lib.alg_set_option(self.h, key.encode(), str(value).encode()))

    def compute_host(self, x, type_, ilist, numneigh, neigh_flat, first, nlocal, nghost, f, eatom=None, vflag=True):
        """x [ntot,3] f64, type_ [ntot] i32, list given as flat neighbours + per-atom offsets
        (converted here into the int** firstneigh LAMMPS passes).  f is accumulated in place.
        returns (eng_vdwl, virial6)"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        type_ = np.ascontiguousarray(type_, dtype=np.int32)
        ilist = np.ascontiguousarray(ilist, dtype=np.int32)
        numneigh = np.ascontiguousarray(numneigh, dtype=np.int32)
        neigh_flat = np.ascontiguousarray(neigh_flat, dtype=np.int32)
        assert f.dtype == np.float64 and f.flags.c_contiguous and f.shape == x.shape
        base = neigh_flat.ctypes.data if neigh_flat.size else 0
        # the firstneigh pointer table LAMMPS hands over is rebuilt only when the list storage changes (this is caller-side
        # marshalling, ~2 ms for 1 M atoms, not part of the plugin call)
        key = (base, first.ctypes.data if isinstance(first, np.ndarray) else id(first), len(first))
        if getattr(self, "_ptr_key", None) != key:
            self._ptrs = np.ascontiguousarray((base + np.asarray(first, dtype=np.int64) * 4).astype(np.uint64))
            self._ptr_key = key
        ptrs = self._ptrs
        eng = C.c_double()
        vir = np.zeros(6)
        self._check(self.lib.alg_compute_host(
            self.h, int(nlocal), int(nghost), _dptr(x), _iptr(type_), _iptr(ilist), _iptr(numneigh),
            ptrs.ctypes.data_as(C.POINTER(C.POINTER(C.c_int))), 1 if eatom is not None else 0, 1 if vflag else 0,
            _dptr(f), _dptr(eatom) if eatom is not None else None, C.byref(eng), _dptr(vir)))
        self._keep = (x, type_, ilist, numneigh, neigh_flat, ptrs)
        return eng.value, vir

    def compute_device(self, nlocal, nghost, d_x, d_type, d_ilist, d_numneigh, d_neighbors, stride_i, stride_jj,
                       d_f, d_eatom=0, want_scalars=True, vflag=True, stream=0):
        """all d_* are raw device addresses (ints).  returns (eng, virial6) or (None, None)."""
        eng = C.c_double()
        vir = np.zeros(6)
        self._check(self.lib.alg_compute_device(
            self.h, int(nlocal), int(nghost), d_x, d_type, d_ilist, d_numneigh, d_neighbors, int(stride_i), int(stride_jj),
            1 if d_eatom else 0, 1 if vflag else 0, d_f, d_eatom or None,
            C.byref(eng) if want_scalars else None, _dptr(vir) if want_scalars else None, stream or None))
        return (eng.value, vir) if want_scalars else (None, None)

    def get_edges(self):
        p = C.POINTER(C.c_int64)()
        n = C.c_int64()
        self._check(self.lib.alg_get_edges(self.h, C.byref(p), C.byref(n)))
        E = n.value
        if E == 0:
            return np.zeros((2, 0), dtype=np.int64)
        return np.ctypeslib.as_array(p, shape=(2 * E,)).reshape(2, E).copy()

    def get_output(self, name):
        p = C.POINTER(C.c_double)()
        n = C.c_int64()
        self._check(self.lib.alg_get_output(self.h, name.encode(), C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros(0)
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def stats(self, what, n):
        t = np.zeros(n)
        self._check(self.lib.alg_get_stats(self.h, what.encode(), _dptr(t), n))
        return t

    def timings(self):
        t = np.zeros(3)
        self._check(self.lib.alg_get_timings(self.h, _dptr(t)))
        return t


class Comm:
    """ghost halo exchange over NCCL (alg_comm_*): the product-side replacement of LAMMPS comm->forward_comm / reverse_comm"""

    def __init__(self, device, nranks=1, rank=0, unique_id=None):
        self.lib = load_library()
        self.c = C.c_void_p()
        rc = self.lib.alg_comm_create(int(device), int(nranks), int(rank), unique_id, C.byref(self.c))
        if rc != 0:
            msg = self.lib.alg_comm_last_error(None).decode()
            self.c = None
            raise AllegroError(rc, msg)
        self.rank, self.nranks = rank, nranks

    @staticmethod
    def unique_id():
        lib = load_library()
        buf = C.create_string_buffer(128)
        rc = lib.alg_comm_unique_id(buf)
        if rc != 0:
            raise AllegroError(rc, lib.alg_comm_last_error(None).decode())
        return buf.raw

    def _check(self, rc):
        if rc != 0:
            raise AllegroError(rc, self.lib.alg_comm_last_error(self.c).decode())

    def set_plan(self, plan):
        """plan = dict(recv_slices={peer: (a, b)}, send_index={peer: int32[n]}, send_shift={peer: f64[n,3]}) as built by
        lmpshim/harness.py (decompose_rank / the single-rank self-image plan)"""
        peers = sorted(plan["send_index"].keys())
        n = len(peers)
        self._keep = []
        pr = (C.c_int * max(n, 1))(*peers)
        sc = (C.c_int * max(n, 1))(*[len(plan["send_index"][p]) for p in peers])
        rb = (C.c_int * max(n, 1))(*[int(plan["recv_slices"][p][0]) for p in peers])
        rcnt = (C.c_int * max(n, 1))(*[int(plan["recv_slices"][p][1] - plan["recv_slices"][p][0]) for p in peers])
        si = (C.POINTER(C.c_int) * max(n, 1))()
        ss = (C.POINTER(C.c_double) * max(n, 1))()
        for k, p in enumerate(peers):
            a = np.ascontiguousarray(plan["send_index"][p], dtype=np.int32)
            b = np.ascontiguousarray(plan["send_shift"][p], dtype=np.float64)
            self._keep += [a, b]
            si[k] = _iptr(a)
            ss[k] = _dptr(b)
        self._check(self.lib.alg_comm_set_plan(self.c, n, pr, sc, si, ss, rb, rcnt))

    def forward(self, d_x, stream=0):
        self._check(self.lib.alg_comm_forward(self.c, d_x, stream or None))

    def reverse(self, d_f, stream=0):
        self._check(self.lib.alg_comm_reverse(self.c, d_f, stream or None))

    def allreduce_sum(self, values, stream=0):
        v = np.ascontiguousarray(values, dtype=np.float64).copy()
        self._check(self.lib.alg_comm_allreduce_sum(self.c, _dptr(v), len(v), stream or None))
        return v

    def stats(self):
        t = np.zeros(4)
        self._check(self.lib.alg_comm_stats(self.c, _dptr(t)))
        return t

    def close(self):
        if getattr(self, "c", None):
            self.lib.alg_comm_destroy(self.c)
            self.c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NeighborBuilder:
    """binned FULL neighbour list on the device (alg_neigh_*): the product-side stand-in for LAMMPS' Neighbor class"""

    def __init__(self, device=0):
        self.lib = load_library()
        self.n = C.c_void_p()
        rc = self.lib.alg_neigh_create(int(device), C.byref(self.n))
        if rc != 0:
            self.n = None
            raise AllegroError(rc, "alg_neigh_create failed")

    def _check(self, rc):
        if rc != 0:
            raise AllegroError(rc, self.lib.alg_neigh_last_error(self.n).decode())

    def build(self, nlocal, nghost, d_x, lo, hi, rneigh, max_neigh, d_neighbors, d_numneigh, stride_i=None, stride_jj=1, want_max=True, stream=0):
        """d_* raw device addresses; the view is LayoutRight [nlocal][max_neigh] unless strides are given.  returns the largest count"""
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        mx = C.c_int(0)
        self._check(self.lib.alg_neigh_build(self.n, int(nlocal), int(nghost), d_x, _dptr(lo), _dptr(hi), float(rneigh), int(max_neigh),
                                              int(max_neigh if stride_i is None else stride_i), int(stride_jj), d_neighbors, d_numneigh,
                                              C.byref(mx) if want_max else None, stream or None))
        return mx.value

    def needs_rebuild(self, ntot, d_x, skin, stream=0):
        r = C.c_int(1)
        self._check(self.lib.alg_neigh_check(self.n, int(ntot), d_x, float(skin), C.byref(r), stream or None))
        return bool(r.value)

    def close(self):
        if getattr(self, "n", None):
            self.lib.alg_neigh_destroy(self.n)
            self.n = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


KERNEL_FAMILIES = ["F0", "FK", "T", "BK", "B0", "fixup", "fused"]


def version():
    return load_library().alg_version().decode()
