"""Where does the TMA back-projection differ from numpy?  Error pattern by row parity / column inside the 128-column group."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from gp_emulator_b200 import _lib
from gp_emulator_b200._lib import addr, check
W = int(os.environ.get("W", 2101)); P = int(os.environ.get("P", 20)); N = int(os.environ.get("N", 1000))
M, D = 50, 3
rs = np.random.RandomState(4)
inputs = rs.random_sample((M, D)); thetas = rs.random_sample((P, D + 2)); invQts = rs.random_sample((P, M))
basis = rs.standard_normal((P, W))
bank = g.DeviceBank(inputs, thetas, invQts, None, basis=basis)
lib = _lib.load()
mu = torch.rand(N, P, dtype=torch.float64, device="cuda")
fwd = torch.full((N, W), float("nan"), dtype=torch.float64, device="cuda")
check(lib.gpe_bank_project(bank._h, addr(mu), None, N, addr(fwd), None, None))
torch.cuda.synchronize()
got = fwd.cpu().numpy(); ref = mu.cpu().numpy() @ basis
bad = ~(np.abs(got - ref) <= 1e-9 * np.abs(ref).max())
print("W=%d P=%d N=%d: %d wrong of %d (%d NaN)" % (W, P, N, bad.sum(), bad.size, np.isnan(got).sum()))
if bad.any():
    r, c = np.nonzero(bad)
    print(" wrong rows by parity: even %d odd %d" % ((r % 2 == 0).sum(), (r % 2 == 1).sum()))
    print(" wrong by (column %% 128):", sorted(set((c % 128).tolist()))[:40])
    print(" wrong by column group:", sorted(set((c // 128).tolist())))
    print(" wrong by (row %% 64):", sorted(set((r % 64).tolist())))
    print(" first few:", list(zip(r[:8].tolist(), c[:8].tolist())), got[r[0], c[0]], ref[r[0], c[0]])
    # is a wrong value some other entry of the reference?
    for k in range(min(5, len(r))):
        hit = np.argwhere(np.abs(ref - got[r[k], c[k]]) < 1e-9)
        print("  got[%d,%d] equals ref at" % (r[k], c[k]), hit[:3].tolist())
