#!/usr/bin/env python
"""tcgen05.mma with the A operand in tensor memory: correctness vs numpy and cycles per MMA
(alg_debug_umma_gemm_ta in csrc/alg_debug.cu)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "pair_allegro_b200", "liballegro_b200_debug.so"))
vp = C.c_void_p
lib.alg_debug_umma_gemm_ta.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
rng = np.random.default_rng(0)
for K, N, passes in ((64, 64, 3), (32, 64, 3), (64, 32, 3), (64, 64, 1)):
    A = rng.standard_normal((K, 128)).astype(np.float32)
    W = rng.standard_normal((K, N)).astype(np.float32)
    Cc = np.zeros((128, N), dtype=np.float32)
    cyc = np.zeros(1, dtype=np.int64)
    rc = lib.alg_debug_umma_gemm_ta(A.ctypes.data, W.ctypes.data, Cc.ctypes.data, K, N, passes, 0, 1, cyc.ctypes.data)
    ref = A.astype(np.float64).T @ W.astype(np.float64)
    err = np.abs(Cc - ref).max() / np.abs(ref).max()
    print("K=%d N=%d passes=%d rc=%d rel err %.2e" % (K, N, passes, rc, err))
for nb in (148, 296):
    for N in (32, 64):
        cyc = np.zeros(1, dtype=np.int64)
        A = np.ones((64, 128), dtype=np.float32); W = np.ones((64, N), dtype=np.float32); Cc = np.zeros((128, N), dtype=np.float32)
        it = 40
        rc = lib.alg_debug_umma_gemm_ta(A.ctypes.data, W.ctypes.data, Cc.ctypes.data, 64, N, 3, it, nb, cyc.ctypes.data)
        print("rate: N=%d blocks=%d rc=%d cycles/MMA %.1f" % (N, nb, rc, cyc[0] / (24.0 * it)))
