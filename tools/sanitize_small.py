"""Small run of every kernel for compute-sanitizer (memcheck / racecheck): tiny sizes, all code paths."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
for M, D, N in ((250, 10, 200), (300, 5, 70), (600, 3, 40), (37, 3, 33)):
    inputs, theta, invQ, invQt, testing = orc.make_S_model(M, D, N, seed=1)
    m = g.DeviceModel(inputs, theta, invQt, invQ)
    o = m.predict(testing, want_hess=True)
    mu, var, deriv = orc.predict(inputs, theta, invQ, invQt, testing)
    assert orc.ref_err(o["var"], var) < 1e-10 and orc.ref_err(o["mu"], mu) < 1e-10
    m.predict(testing, want_var=False)
    if M <= 256:
        o32 = m.predict_f32(testing.astype(np.float32))
        assert orc.ref_err(o32["var"], var) < 1e-3
rs = np.random.RandomState(0)
E, M, D, W = 3, 40, 4, 300
inputs = rs.random_sample((M, D))
bank = g.DeviceBank(inputs, rs.random_sample((E, D + 2)), rs.random_sample((E, M)), rs.random_sample((E, M, M)),
                    basis=rs.random_sample((E, W)))
bank.predict(rs.random_sample((50, D)), want_var=True, want_deriv=True, want_hess=True, project=True, project_deriv=True)
torch.cuda.synchronize()
print("sanitize run ok")
