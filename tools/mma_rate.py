#!/usr/bin/env python
"""tcgen05.mma (kind::tf32, M=128) issue-rate microbenchmark: cycles per MMA vs N and the number of
independent TMEM accumulators (alg_debug_mma_rate in csrc/alg_debug.cu)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "pair_allegro_b200", "liballegro_b200_debug.so"))
lib.alg_debug_mma_rate.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_longlong)]
out = (C.c_longlong * 2)()
print("N nacc groups sync blocks : cycles/MMA (wall), cycles/MMA inside the issue loop, cycles per commit round trip")
iters = 50
for N in (32, 64, 128, 256):
    for nacc in (1, 2):
        if N * nacc > 256:
            continue
        for groups, sync in ((3, 1), (3, 0), (1, 1), (16, 1)):
            for nb in (148, 296):
                rc = lib.alg_debug_mma_rate(N, nacc, groups, iters, sync, nb, out)
                n = 8.0 * groups * iters
                print("%3d %d %2d %d %3d : %6.1f %6.1f %8.1f%s" % (N, nacc, groups, sync, nb, out[0] / n, out[1] / n, out[0] / iters, "" if rc == 0 else "  rc=%d" % rc))
