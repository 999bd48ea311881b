"""Where does the host-resident path spend its time?  PCIe copy rates, library call with preallocated pinned
buffers, and the public API call."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from gp_emulator_b200 import _lib
from oracle import gp_oracle as orc

N = int(float(os.environ.get("N", 2e7))); D = 10
def t(f, reps=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

hin = torch.rand(N, D, dtype=torch.float64).pin_memory()
din = torch.empty(N, D, dtype=torch.float64, device="cuda")
hout = torch.empty(N, D + 2, dtype=torch.float64).pin_memory()
dout = torch.empty(N, D + 2, dtype=torch.float64, device="cuda")
s = t(lambda: din.copy_(hin, non_blocking=True)); print("H2D pinned  %.1f GB/s" % (hin.numel() * 8 / s / 1e9))
s = t(lambda: hout.copy_(dout, non_blocking=True)); print("D2H pinned  %.1f GB/s" % (hout.numel() * 8 / s / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
    with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
s = t(both); print("H2D+D2H concurrent: %.1f GB/s total" % ((hin.numel() + hout.numel()) * 8 / s / 1e9))
pg = np.random.rand(N // 4, D)
s = t(lambda: din[: N // 4].copy_(torch.from_numpy(pg))); print("H2D pageable %.1f GB/s" % (pg.size * 8 / s / 1e9))

inputs, theta, invQ, invQt, _ = orc.make_S_model(250, 10, 1, seed=0)
m = g.DeviceModel(inputs, theta, invQt, invQ)
lib = _lib.load()
mu = torch.empty(N, dtype=torch.float64).pin_memory(); var = torch.empty(N, dtype=torch.float64).pin_memory()
der = torch.empty(N, D, dtype=torch.float64).pin_memory()
def libcall():
    rc = lib.gpe_predict(m._h, hin.data_ptr(), N, mu.data_ptr(), var.data_ptr(), der.data_ptr(), None, 0x107, None)
    assert rc == 0, lib.gpe_last_error()
s = t(libcall); print("gpe_predict host ptrs, preallocated pinned: %.3f s  %.3e pts/s" % (s, N / s))
dev = t(lambda: m.predict(din)); print("device-resident: %.3f s %.3e pts/s" % (dev, N / dev))
hnp = hin.numpy()
s = t(lambda: m.predict(hnp)); print("DeviceModel.predict(numpy pinned): %.3f s  %.3e pts/s" % (s, N / s))
gp = g.GaussianProcess(inputs, []); gp.theta, gp.invQ, gp.invQt = theta, invQ, invQt
s = t(lambda: gp.predict(hnp)); print("GaussianProcess.predict(numpy pinned): %.3f s  %.3e pts/s" % (s, N / s))
t0 = time.perf_counter(); x = torch.empty(N, D, dtype=torch.float64, pin_memory=True); print("pinned alloc 1.6GB first: %.3f s" % (time.perf_counter() - t0))
del x
t0 = time.perf_counter(); x = torch.empty(N, D, dtype=torch.float64, pin_memory=True); print("pinned alloc 1.6GB again: %.3f s" % (time.perf_counter() - t0))
