import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pair_allegro_b200 import capi
lib = capi.load_debug_library()
fn = lib.alg_debug_umma_gemm2
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]; fn.restype = C.c_int
np.set_printoptions(precision=3, linewidth=220, suppress=True)
for (K, N, passes, variant) in [(8, 32, 1, 8), (8, 32, 1, 9), (32, 64, 1, 0), (32, 64, 1, 1), (64, 64, 3, 0), (64, 64, 3, 1)]:
    rng = np.random.default_rng(1)
    A = rng.normal(size=(K, 128)).astype(np.float32)
    W = rng.normal(size=(K, N)).astype(np.float32)
    out = np.full((128, N), -77.0, dtype=np.float32)
    dump = np.zeros((128, 128), dtype=np.float32)
    rc = fn(A.ctypes.data, W.ctypes.data, out.ctypes.data, K, N, passes, variant, dump.ctypes.data)
    ref = A.astype(np.float64).T @ W.astype(np.float64)
    print("K,N,passes,variant", K, N, passes, variant, "rc", rc, "relerr", np.abs(out - ref).max() / np.abs(ref).max(), "nonzero in dump", int((dump != 0).sum()))
    print(" out[0,:8]", out[0, :8]); print(" ref[0,:8]", ref[0, :8])
    if np.abs(dump).max() > 0:
        nzr = np.nonzero(np.abs(dump).sum(1))[0]; nzc = np.nonzero(np.abs(dump).sum(0))[0]
        print(" dump nonzero rows", nzr[:10], "...", len(nzr), "cols", nzc[:10], "...", len(nzc))
        c = np.corrcoef(np.concatenate([dump[:, :N], ref], 0))[:128, 128:]
        print(" best ref row for dump rows 0..7:", np.argmax(np.abs(np.nan_to_num(c[:8])), 1), np.max(np.abs(np.nan_to_num(c[:8])), 1))
