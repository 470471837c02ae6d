"""tcgen05 tensor-core GEMM primitive (csrc/umma.cuh) checked on the GPU against numpy:
operand layout (K-major SWIZZLE_128B), descriptors, TMEM allocation / loads, and the
3xTF32 split that gives fp32-level accuracy."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(32, 32), (32, 64), (64, 64), (128, 64), (64, 128), (96, 64), (64, 96), (96, 32)])
@pytest.mark.parametrize("passes", [1, 3])
def test_umma_gemm(K, N, passes, ensure_built):
    from pair_allegro_b200 import capi
    lib = capi.load_debug_library()
    fn = lib.alg_debug_umma_gemm
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    fn.restype = C.c_int
    rng = np.random.default_rng(K * 131 + N)
    A = rng.normal(size=(K, 128)).astype(np.float32)       # [k][m]
    W = (rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)
    out = np.zeros((128, N), dtype=np.float32)
    rc = fn(A.ctypes.data, W.ctypes.data, out.ctypes.data, K, N, passes)
    assert rc == 0
    ref = A.astype(np.float64).T @ W.astype(np.float64)
    err = np.abs(out - ref).max() / np.abs(ref).max()
    print("K=%d N=%d passes=%d rel err %.3e" % (K, N, passes, err))
    assert err < (3e-3 if passes == 1 else 3e-6)
