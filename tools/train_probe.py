#!/usr/bin/env python
"""Batched training objective on the GPU: evaluations/s against the host numpy path, and one end-to-end fit.

    python tools/train_probe.py [--fit]      (GPU box)
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gp_emulator_b200 import GaussianProcess, MultivariateEmulator  # noqa: E402
from gp_emulator_b200.training import DeviceTrainer  # noqa: E402


def main():
    out = {}
    rs = np.random.RandomState(0)
    for M, D in ((250, 10), (100, 10), (500, 10), (1000, 10)):
        x = rs.random_sample((M, D))
        T = 12
        targets = np.sin(np.arange(1, T + 1)[:, None] * x.sum(axis=1)[None, :])
        tr = DeviceTrainer(x, targets)
        row = {}
        for B in (1, 15, 148, 180, 296, 1184):
            if M >= 1000 and B > 296:
                continue
            th = 5.0 * (rs.random_sample((B, D + 2)) - 0.5)
            ti = (np.arange(B) % T).astype(np.int32)
            tr.evaluate(th, ti)
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                ll, g, st = tr.evaluate(th, ti)
            dt = (time.perf_counter() - t0) / reps
            row[B] = {"ms": dt * 1e3, "evals_per_s": B / dt, "gfma_per_s": B * M ** 3 / dt / 1e9, "failed": int(st.sum())}
        tr.close()
        gp = GaussianProcess(x, targets[0])
        th = 5.0 * (rs.random_sample(D + 2) - 0.5)
        gp.loglikelihood(th)
        t0 = time.perf_counter()
        for _ in range(2):
            gp.loglikelihood(th); gp.partial_devs(th)
        host_ms = (time.perf_counter() - t0) / 2 * 1e3
        out["M%d_D%d" % (M, D)] = {"gpu": row, "host_numpy_ms_per_eval": host_ms}
        print("M=%d D=%d  host %.1f ms/eval;" % (M, D, host_ms),
              "  ".join("B=%d: %.2f ms (%.0f/s)" % (b, r["ms"], r["evals_per_s"]) for b, r in row.items()), flush=True)
    if "--fit" in sys.argv:
        # a PROSAIL-sized MultivariateEmulator: 250 training spectra, 10 parameters, n_tries = 5 per PC
        M, D, W = 250, 10, 2101
        y = rs.random_sample((M, D))
        wl = np.linspace(0.0, 1.0, W)
        X = sum(np.sin((k + 1) * 2.0 * wl[None, :] * y[:, k:k + 1] + k) / (k + 1) for k in range(D))
        np.random.seed(1)
        t0 = time.perf_counter()
        mv = MultivariateEmulator(X=X, y=y, thresh=0.97, n_tries=5, batched_training=True)
        dt = time.perf_counter() - t0
        out["fit"] = {"n_pcs": mv.n_pcs, "n_tries": 5, "seconds": dt, **mv.training_stats}
        print("fit: %d PCs x 5 starts in %.1f s (%d evaluations in %d rounds)" % (mv.n_pcs, dt, mv.training_stats["evaluations"],
                                                                               mv.training_stats["rounds"]), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/train_probe.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
