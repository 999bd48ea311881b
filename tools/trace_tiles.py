"""Per-tile phase timing of CTA 0 of the fused kernel (clock64 at phase boundaries)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from gp_emulator_b200 import _lib
from oracle import gp_oracle as orc
M = int(os.environ.get("M", 250))
inputs, theta, invQ, invQt, _ = orc.make_S_model(M, 10, 1, seed=0)
m = g.DeviceModel(inputs, theta, invQt, invQ, symmetric_variance=bool(int(os.environ.get("SYM", "0"))))
t = torch.rand(int(float(os.environ.get("N", 2e6))), 10, dtype=torch.float64, device="cuda")
HESS = bool(int(os.environ.get("HESS", "0")))
run = lambda: m.predict(t, want_hess=HESS)
run(); torch.cuda.synchronize()
buf = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.gpe_debug_trace.argtypes = [C.c_void_p]
lib.gpe_debug_trace(buf.data_ptr())
run(); torch.cuda.synchronize()
lib.gpe_debug_trace(None)
NP = 8 if HESS else 6
tr = buf.cpu().numpy().reshape(64, 8)[4:60, :NP]
d = np.diff(tr, axis=1)
names = ["rows->regs+sync", "phase A j-loop", "reduce+outs+writes", "phase B DMMA loop", "epilogue+syncs",
         "phase C DMMA loop", "S2 staging + Hessian out"]
tot = tr[:, NP - 1] - tr[:, 0]
print("cycles per tile (median over tiles 4..59): total %d, tile-to-tile period %d" % (np.median(tot), np.median(np.diff(tr[:, 0]))))
for i, n in enumerate(names[:NP - 1]):
    print("  %-22s %7d  (%4.1f%%)" % (n, np.median(d[:, i]), 100 * np.median(d[:, i]) / np.median(tot)))
