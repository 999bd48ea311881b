"""ncu targets of round 2: WHAT=proj (k_project), sym (symmetric-folded fused kernel), small (M = 100 fused kernel),
train (k_train_eval), full (headline fused kernel)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from gp_emulator_b200 import _lib
from gp_emulator_b200._lib import addr, check
from oracle import gp_oracle as orc
what = os.environ.get("WHAT", "proj")
if what == "proj":
    M, D, P, W, N = 250, 10, 20, 2101, 200_000
    rs = np.random.RandomState(4)
    inputs = rs.random_sample((M, D)); thetas = rs.random_sample((P, D + 2)); invQts = rs.random_sample((P, M))
    basis = np.linalg.qr(rs.standard_normal((W, P)))[0].T.copy()
    bank = g.DeviceBank(inputs, thetas, invQts, None, basis=basis)
    lib = _lib.load()
    mu = torch.rand(N, P, dtype=torch.float64, device="cuda")
    fwd = torch.empty(N, W, dtype=torch.float64, device="cuda")
    for _ in range(3):
        check(lib.gpe_bank_project(bank._h, addr(mu), None, N, addr(fwd), None, None))
elif what in ("sym", "small", "full"):
    M = 100 if what == "small" else 250
    inputs, theta, invQ, invQt, _ = orc.make_S_model(M, 10, 1, seed=0)
    m = g.DeviceModel(inputs, theta, invQt, invQ, symmetric_variance=(what == "sym"))
    t = torch.rand(int(float(os.environ.get("N", 2e6))), 10, dtype=torch.float64, device="cuda")
    for _ in range(3):
        m.predict(t)
elif what == "train":
    from gp_emulator_b200.training import DeviceTrainer
    inputs, theta, invQ, invQt, _ = orc.make_S_model(250, 10, 1, seed=0)
    tr = DeviceTrainer(inputs, np.sin(inputs.sum(axis=1)), device=0)
    th = 5.0 * (np.random.RandomState(1).random_sample((148, 12)) - 0.5)
    for _ in range(3):
        tr.evaluate(th)
torch.cuda.synchronize()
print("ok")
