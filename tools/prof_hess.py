"""ncu target: mean + gradient kernel, then the fused kernel with the Hessian phase (mu + var + deriv + hess)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
inputs, theta, invQ, invQt, _ = orc.make_S_model(250, 10, 1, seed=0)
m = g.DeviceModel(inputs, theta, invQt, invQ)
N = int(float(os.environ.get("N", 2e6)))
t = torch.rand(N, 10, dtype=torch.float64, device="cuda")
out = {k: torch.empty((N,) + s, dtype=torch.float64, device="cuda") for k, s in
       [("mu", ()), ("var", ()), ("deriv", (10,)), ("hess", (10, 10))]}
for _ in range(3):
    m.predict(t, want_var=False, out=out)
    m.predict(t, want_var=True, want_hess=True, out=out)
torch.cuda.synchronize()
print("ok")
