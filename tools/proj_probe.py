import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from gp_emulator_b200 import _lib
from gp_emulator_b200._lib import addr, check
M, D, P, W, N = 250, 10, 20, 2101, 200_000
rs = np.random.RandomState(4)
inputs = rs.random_sample((M, D)); thetas = rs.random_sample((P, D + 2)); invQts = rs.random_sample((P, M))
basis = np.linalg.qr(rs.standard_normal((W, P)))[0].T.copy()
bank = g.DeviceBank(inputs, thetas, invQts, None, basis=basis)
lib = _lib.load()
mu = torch.rand(N, P, dtype=torch.float64, device="cuda")
der = torch.rand(N, P, D, dtype=torch.float64, device="cuda")
fwd = torch.empty(N, W, dtype=torch.float64, device="cuda")
def ev(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
s = ev(lambda: fwd.fill_(1.0)); print("fill_ 3.36 GB: %.3f ms  %.0f GB/s" % (s * 1e3, fwd.numel() * 8 / s / 1e9))
s = ev(lambda: check(lib.gpe_bank_project(bank._h, addr(mu), None, N, addr(fwd), None, None)))
print("project fwd: %.3f ms  %.0f GB/s  %.1f TFLOP/s" % (s * 1e3, fwd.numel() * 8 / s / 1e9, N * 2 * P * W / s / 1e12))
ref = mu[:512].cpu().numpy() @ basis
print("err", float(np.max(np.abs(fwd[:512].cpu().numpy() - ref)) / np.max(np.abs(ref))))
Nd = 20000
dfull = torch.empty(Nd, D, W, dtype=torch.float64, device="cuda")
s = ev(lambda: check(lib.gpe_bank_project(bank._h, addr(mu), addr(der), Nd, None, addr(dfull), None)))
print("project deriv_full (N=%d): %.3f ms  %.0f GB/s" % (Nd, s * 1e3, dfull.numel() * 8 / s / 1e9))
ref = np.einsum("npd,pw->ndw", der[:64].cpu().numpy(), basis)
print("err", float(np.max(np.abs(dfull[:64].cpu().numpy() - ref)) / np.max(np.abs(ref))))
