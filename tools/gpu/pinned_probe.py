"""How fast does the CPU read pinned host memory on this box?  (cudaHostAlloc vs cudaHostRegister'ed malloc vs pageable)"""
import ctypes as C, time
import numpy as np, torch
n = 64 << 20  # int32 elements = 256 MB
g = torch.arange(n, dtype=torch.int32, device="cuda")
def rate(a, label):
    t0 = time.perf_counter(); s = int(a.sum(dtype=np.int64)); dt = time.perf_counter() - t0
    print(f"{label}: CPU read {a.nbytes / dt / 1e9:.2f} GB/s (sum {s})", flush=True)
pinned = torch.empty(n, dtype=torch.int32, pin_memory=True)
pinned.copy_(g); torch.cuda.synchronize()
rate(pinned.numpy(), "cudaHostAlloc pinned, after a D2H copy")
rate(pinned.numpy(), "cudaHostAlloc pinned, second pass")
page = torch.empty(n, dtype=torch.int32)
page.copy_(g); torch.cuda.synchronize()
rate(page.numpy(), "pageable, after a D2H copy")
reg = np.empty(n, np.int32); reg[:] = 0
rt = torch.cuda.cudart()
rc = rt.cudaHostRegister(reg.ctypes.data, reg.nbytes, 0)
print("cudaHostRegister rc", rc)
t = torch.from_numpy(reg)
t0 = time.perf_counter(); t.copy_(g, non_blocking=True); torch.cuda.synchronize(); print(f"D2H into registered memory: {reg.nbytes/(time.perf_counter()-t0)/1e9:.1f} GB/s")
rate(reg, "cudaHostRegister'ed malloc, after a D2H copy")
t0 = time.perf_counter(); pinned.copy_(g, non_blocking=True); torch.cuda.synchronize(); print(f"D2H into cudaHostAlloc memory: {reg.nbytes/(time.perf_counter()-t0)/1e9:.1f} GB/s")
w = pinned.numpy()
t0 = time.perf_counter(); w[:] = 7; dt = time.perf_counter() - t0; print(f"CPU write to cudaHostAlloc pinned: {w.nbytes/dt/1e9:.2f} GB/s")
t0 = time.perf_counter(); reg[:] = 7; dt = time.perf_counter() - t0; print(f"CPU write to registered: {reg.nbytes/dt/1e9:.2f} GB/s")
