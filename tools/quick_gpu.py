"""Developer loop on the GPU box: one parity check + device-timed throughput of the headline config."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc

M = int(os.environ.get("M", 250)); D = int(os.environ.get("D", 10)); N = int(float(os.environ.get("N", 4e6)))
inputs, theta, invQ, invQt, testing = orc.make_S_model(M, D, 2000, seed=0)
m = g.DeviceModel(inputs, theta, invQt, invQ, symmetric_variance=bool(int(os.environ.get("SYM", "0"))))
out = m.predict(testing)
mu, var, deriv = orc.predict(inputs, theta, invQ, invQt, testing)
errs = {k: orc.ref_err(out[k], r) for k, r in (("mu", mu), ("var", var), ("deriv", deriv))}
print("parity", errs, flush=True)

t = torch.rand(N, D, dtype=torch.float64, device="cuda")
for want_var in (True, False):
    for _ in range(2):
        m.predict(t, want_var=want_var)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        m.predict(t, want_var=want_var)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    F = 2 * M * M + M * (5 * D + 6) + D + 1 if want_var else M * (5 * D + 4) + D + 1
    print(json.dumps({"want_var": want_var, "M": M, "D": D, "N": N, "ms": ms, "pts_per_s": N / ms * 1e3,
                      "alg_tflops": N * F / ms * 1e3 / 1e12}), flush=True)
