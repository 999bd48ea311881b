"""Dense vs symmetric-folded variance contraction: points/s on device-resident points (one GPU)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
def rate(fn, n, reps=3):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return n * reps / (a.elapsed_time(b) * 1e-3)
for M, N in ((250, 20_000_000), (200, 20_000_000), (128, 20_000_000), (500, 4_000_000), (1000, 2_000_000)):
    inputs, theta, invQ, invQt, tt = orc.make_S_model(M, 10, 500, seed=0)
    t = torch.rand(N, 10, dtype=torch.float64, device="cuda")
    ref = orc.predict(inputs, theta, invQ, invQt, tt)
    res = []
    for sym in (False, True):
        m = g.DeviceModel(inputs, theta, invQt, invQ, symmetric_variance=sym)
        err = orc.ref_err(m.predict(tt)["var"], ref[1])
        res.append((rate(lambda: m.predict(t), N), err))
        m.close()
    F = 2 * M * M + M * 56 + 11
    print("M=%4d dense %.3e pts/s (%.1f TF, err %.1e) | symmetric %.3e pts/s (x%.2f, err %.1e)" %
          (M, res[0][0], res[0][0] * F / 1e12, res[0][1], res[1][0], res[1][0] / res[0][0], res[1][1]), flush=True)
    del t
