"""First light for the tcgen05 / TMEM single-precision kernel: parity vs the FP64 oracle + throughput."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
M = int(os.environ.get("M", 250)); D = int(os.environ.get("D", 10))
inputs, theta, invQ, invQt, testing = orc.make_S_model(M, D, 1000, seed=0)
m = g.DeviceModel(inputs, theta, invQt, invQ)
mu, var, deriv = orc.predict(inputs, theta, invQ, invQt, testing)
o = m.predict_f32(testing.astype(np.float32), want_var=False)
print("mean-only: mu %.2e deriv %.2e" % (orc.ref_err(o["mu"], mu), orc.ref_err(o["deriv"], deriv)), flush=True)
t32 = testing.astype(np.float32)
mu, var, deriv = orc.predict(inputs, theta, invQ, invQt, t32.astype(np.float64))
for fast in (True, False):
    o = m.predict_f32(t32, fast=fast)
    print("fast=%s: mu %.2e var %.2e deriv %.2e" % (fast, orc.ref_err(o["mu"], mu), orc.ref_err(o["var"], var), orc.ref_err(o["deriv"], deriv)), flush=True)
N = int(float(os.environ.get("N", 2e7)))
t = torch.rand(N, D, dtype=torch.float32, device="cuda")
for wv, fast in ((True, False), (True, True), (False, False)):
    for _ in range(2): m.predict_f32(t, want_var=wv, fast=fast)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = m.predict_f32(t, want_var=wv, fast=fast)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"want_var": wv, "fast_tf32": fast, "M": M, "N": N, "ms": ms, "pts_per_s": N / ms * 1e3}), flush=True)
