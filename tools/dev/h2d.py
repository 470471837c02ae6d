import ctypes, time, numpy as np, torch
n = 27_000_000
a = np.random.rand(n // 8)            # pageable numpy
d = torch.empty(n // 8, dtype=torch.float64, device="cuda")
rt = torch.cuda.cudart()
def bw(src_t, label, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): d.copy_(src_t, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    print("%-28s %.2f ms  %.1f GB/s" % (label, dt * 1e3, n / dt / 1e9))
t = torch.from_numpy(a)
bw(t, "pageable numpy")
rc = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
print("cudaHostRegister rc", rc)
bw(t, "registered numpy")
p = torch.empty(n // 8, dtype=torch.float64).pin_memory()
bw(p, "torch pinned")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): p.copy_(d, non_blocking=True)
torch.cuda.synchronize(); print("D2H pinned %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
