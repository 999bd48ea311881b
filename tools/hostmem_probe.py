"""Host-side costs of the staged (pageable caller) path on this box: memcpy bandwidth vs threads, first-touch cost
of fresh numpy arrays, MADV_POPULATE_WRITE, pageable cudaMemcpy by the driver."""
import ctypes as C, mmap, os, sys, threading, time
import numpy as np, torch
libc = C.CDLL("libc.so.6", use_errno=True)
libc.memcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]; libc.memcpy.restype = C.c_void_p
libc.madvise.argtypes = [C.c_void_p, C.c_size_t, C.c_int]; libc.madvise.restype = C.c_int
print("cpus:", os.cpu_count(), "affinity:", len(os.sched_getaffinity(0)))
try:
    print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
except Exception as e:
    print("THP: ?", e)
NB = 256 << 20
src = np.ones(NB // 8); dst = np.ones(NB // 8)
def par(dst_p, src_p, nbytes, nt):
    part = (nbytes // nt + 63) & ~63
    th = []
    for i in range(nt):
        off = i * part
        ln = min(part, nbytes - off)
        if ln <= 0: break
        t = threading.Thread(target=libc.memcpy, args=(dst_p + off, src_p + off, ln)); t.start(); th.append(t)
    for t in th: t.join()
for nt in (1, 2, 4, 8, 12, 16, 24):
    par(dst.ctypes.data, src.ctypes.data, NB, nt)
    t0 = time.perf_counter()
    for _ in range(3): par(dst.ctypes.data, src.ctypes.data, NB, nt)
    s = (time.perf_counter() - t0) / 3
    print("memcpy warm pages, %2d threads: %.1f GB/s" % (nt, NB / s / 1e9))
pin = torch.empty(NB // 8, dtype=torch.float64).pin_memory()
for nt in (1, 4, 8, 16):
    t0 = time.perf_counter()
    for _ in range(3): par(dst.ctypes.data, pin.data_ptr(), NB, nt)
    s = (time.perf_counter() - t0) / 3
    print("memcpy pinned -> warm pageable, %2d threads: %.1f GB/s" % (nt, NB / s / 1e9))
for nt in (1, 4, 8, 16):
    s = 0
    for _ in range(3):
        fresh = np.empty(NB // 8)
        t0 = time.perf_counter(); par(fresh.ctypes.data, pin.data_ptr(), NB, nt); s += time.perf_counter() - t0
        del fresh
    print("memcpy pinned -> FRESH np.empty, %2d threads: %.1f GB/s" % (nt, NB / (s / 3) / 1e9))
MADV_POPULATE_WRITE = 23
s = 0
for _ in range(3):
    fresh = np.empty(NB // 8)
    a = fresh.ctypes.data; a0 = (a + 4095) & ~4095
    t0 = time.perf_counter(); rc = libc.madvise(a0, (NB - (a0 - a)) & ~4095, MADV_POPULATE_WRITE); s += time.perf_counter() - t0
    del fresh
print("MADV_POPULATE_WRITE of a fresh 256 MB array: rc=%d  %.1f GB/s" % (rc, NB / (s / 3) / 1e9))
d = torch.empty(NB // 8, dtype=torch.float64, device="cuda")
t_src = torch.from_numpy(src)
d.copy_(t_src); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3): d.copy_(t_src)
torch.cuda.synchronize(); s = (time.perf_counter() - t0) / 3
print("driver pageable H2D: %.1f GB/s" % (NB / s / 1e9))
t_dst = torch.from_numpy(dst)
t_dst.copy_(d); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3): t_dst.copy_(d)
torch.cuda.synchronize(); s = (time.perf_counter() - t0) / 3
print("driver pageable D2H: %.1f GB/s" % (NB / s / 1e9))
