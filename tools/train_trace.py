"""Phase timing of CTA 0 of k_train_eval (clock64 durations, dev aid gpe_debug_train_trace): python tools/train_trace.py [M] [D] [B]"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gp_emulator_b200 import _lib
from gp_emulator_b200.training import DeviceTrainer
M, D, B = (int(v) for v in (sys.argv[1:4] + ["250", "10", "148"][len(sys.argv) - 1:]))
rs = np.random.RandomState(0)
x = rs.random_sample((M, D))
tr = DeviceTrainer(x, np.sin(x.sum(axis=1)))
th = rs.random_sample((B, D + 2)) - 0.5
tr.evaluate(th)
buf = torch.zeros(8, dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.gpe_debug_train_trace.argtypes = [C.c_void_p]
lib.gpe_debug_train_trace(buf.data_ptr())
ll, g, st = tr.evaluate(th)
torch.cuda.synchronize()
lib.gpe_debug_train_trace(None)
v = buf.cpu().numpy()
names = ["inputs + covariance", "staging rows/cols", "pivot-block inverse", "coefficients", "rank-NB update", "alpha + sums", "gradient sums"]
print("M=%d D=%d B=%d: %d cycles in CTA 0 (%d failed problems)" % (M, D, B, v[:7].sum(), int(st.sum())))
for n, c in zip(names, v):
    print("  %-22s %8d  (%4.1f%%)" % (n, c, 100.0 * c / max(1, v[:7].sum())))
