"""The drop-in call as a reference user makes it: plain (pageable) numpy in, fresh numpy arrays out, for a sweep of
N.  Reports points/s against the device-resident kernel rate."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc

M, D = 250, 10
inputs, theta, invQ, invQt, _ = orc.make_S_model(M, D, 1, seed=0)
gp = g.GaussianProcess(inputs, []); gp.theta, gp.invQ, gp.invQt = theta, invQ, invQt
rs = np.random.RandomState(1)
for N in [1, 100, 1000, 10_000, 100_000, 1_000_000, 10_000_000]:
    x = rs.random_sample((N, D))
    for _ in range(4):
        mu, var, der = gp.predict(x)
        mu2, der2 = gp.predict(x, do_unc=False)
    reps = 200 if N <= 10_000 else (20 if N <= 1_000_000 else 3)
    t0 = time.perf_counter()
    for _ in range(reps):
        mu, var, der = gp.predict(x)
    s = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        mu2, der2 = gp.predict(x, do_unc=False)
    s2 = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        mu3, var3, der3 = gp.predict(x, pinned=False)
    s3 = (time.perf_counter() - t0) / reps
    print("N=%9d  predict %10.1f us  %.3e pts/s (%.1f GB/s of host traffic) | do_unc=False %10.1f us %.3e pts/s | pinned=False %10.1f us %.3e pts/s"
          % (N, s * 1e6, N / s, N * 176 / s / 1e9, s2 * 1e6, N / s2, s3 * 1e6, N / s3), flush=True)
