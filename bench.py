#!/usr/bin/env python
"""Headline benchmark: test-point predictions/sec (mean + variance + gradient, M=250, D=10, FP64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

* workload = BASELINE.json configs[1]: the tests/benchmark.py-style synthetic GaussianProcess
  (M=250, D=10, all-U(0,1) model, seed 0), 1e8 FP64 test points PER GPU generated on the device
  (weak scaling: test points shard trivially, no data-path collective; rank 0 broadcasts the model once).
* a step = one pass of the fused predict kernel over the rank's 1e8 points (one kernel launch);
  `value` = points of all ranks / max-over-ranks device time (CUDA events on the launching stream).
* `e2e` = the same metric through the public drop-in API GaussianProcess.predict() with HOST buffers
  (pinned input, H2D + kernels + D2H inside the timed region) on --e2e-points points per GPU per step.
* `roofline`: the kernel is FP64-arithmetic bound; achieved = N * F(M, D) flop per
  launch / launch duration with F = 2M^2 + M(5D+6) + D + 1 = 139,011; peak = the DMMA (FP64 tensor)
  throughput of THIS GPU measured live by the library's micro-benchmark, because the driver-written
  MEASURED_PEAKS.json holds no FP64 figure.
* `cpu_baseline` / `--impl reference`: the reference's numpy/scipy cpu_predict (restated in oracle/, the
  reference itself is Python 2 and cannot be imported) on the box's host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M, D = 250, 10
F_PER_POINT = 2 * M * M + M * (5 * D + 6) + D + 1          # 139,011 flop (+ M exp), SURVEY.md section 8d
BYTES_PER_POINT = 8 * D + 8 * (2 + D)                      # 176 B
METRIC = "test-point predictions/sec (mean+var+grad, M=250, D=10, FP64)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=float, default=1e8, help="device-resident test points per GPU per step")
    ap.add_argument("--e2e-points", type=float, default=2e7, help="host-resident test points per GPU per step")
    ap.add_argument("--cpu-points", type=float, default=1e5, help="sample size of the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(n_points):
    return {"workload": "cfg2 GaussianProcess S-model (tests/benchmark.py recipe, seed 0) M=250 D=10 FP64 "
                        "mu+var+grad, test points per GPU per step below, generated on device",
            "points_per_gpu_per_step": int(n_points),
            "l2": "inputs (%.1f GB per step) larger than L2" % (n_points * D * 8 / 1e9),
            "sharding": "contiguous test-point ranges per rank, model broadcast once, no data-path collective"}


def synthetic_model():
    # the reference benchmark's recipe (tests/benchmark.py:11-15, 28-29): everything U(0, 1), drawn in the order
    # inputs, testing, theta, invQ, invQt; legacy RandomState so the stream is the same on every numpy
    rs = np.random.RandomState(0)
    inputs = rs.random_sample((M, D))
    rs.random_sample((1, D))
    theta = rs.random_sample(D + 2)
    invQ = rs.random_sample((M, M))
    invQt = rs.random_sample(M)
    return {"inputs": inputs, "theta": theta, "invQ": invQ, "invQt": invQt}


# --------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core BLAS can take."""
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=os.cpu_count())
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count()


def cpu_baseline(model, npts, repeats=3):
    """Reference cpu_predict semantics (oracle port) on the host cores; best of `repeats`."""
    from oracle import gp_oracle as orc
    threads = _use_all_host_threads()
    testing = np.random.RandomState(1).random_sample((int(npts), D))
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        orc.predict(model["inputs"], model["theta"], model["invQ"], model["invQt"], testing, chunk=100000)
        best = min(best, time.perf_counter() - t0)
    return {"value": npts / best, "unit": "points/s", "cores": int(threads), "kind": "port",
            "sample": "%d points of the same workload (one 1e5-point chunk, as the reference materialises (M,N) "
                      "matrices), best of %d; numpy/scipy cpu_predict restated from GaussianProcess.py:211-251" %
                      (int(npts), repeats)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    model = synthetic_model()
    from oracle import gp_oracle as orc
    npts = int(args.cpu_points)
    threads = _use_all_host_threads()
    testing = np.random.RandomState(1).random_sample((npts, D))
    for _ in range(args.warmup):
        orc.predict(model["inputs"], model["theta"], model["invQ"], model["invQt"], testing[:20000])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.predict(model["inputs"], model["theta"], model["invQ"], model["invQt"], testing, chunk=100000)
    dt = time.perf_counter() - t0
    value = npts * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.points),
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": int(threads), "kind": "port",
                         "sample": "each step = %d points of the workload on the host CPU (the reference materialises "
                                   "(M,N) matrices, 1e8 points would need 200 GB); numpy/scipy cpu_predict restated "
                                   "in oracle/ (the reference is Python 2 and cannot be imported)" % npts},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch ourselves under torchrun the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"),
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import gp_emulator_b200 as gpe
    from gp_emulator_b200 import _lib, sharding

    # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION prints it there) off it
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    # one process per GPU: keep each rank (and the page-locked buffers it touches first) on its GPU's NUMA node
    numa_cpus = sharding.bind_host_to_gpu(local_rank) if (world > 1 and not os.environ.get("GPE_NO_NUMA_BIND")) else None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    model = synthetic_model() if rank == 0 else None
    if world > 1:
        # NCCL initialises lazily and may print (version banner, NCCL_DEBUG output) on fd 1: point fd 1 at stderr
        # until the communicator exists, so that stdout carries exactly one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            model = sharding.broadcast_model(model, src=0, device=dev)   # the only collective on the path
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    gp = gpe.GaussianProcess(model["inputs"], [], device=local_rank)
    gp.theta, gp.invQ, gp.invQt = model["theta"], model["invQ"], model["invQt"]   # tests/benchmark.py:11-15
    dm = gp._device_model()
    lib = _lib.load()

    N = int(args.points)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    testing = torch.rand(N, D, dtype=torch.float64, device=dev, generator=gen)   # 8 GB >> 126 MB L2
    peaks = gpe.measure_fp64_peaks(local_rank) if rank == 0 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident: `value` and the roofline ------------------------------------------------
    out = None
    for _ in range(args.warmup):
        out = dm.predict(testing)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.gpe_launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for i in range(args.steps):
        out = dm.predict(testing)
        evs[i + 1].record()
    barrier()
    launches = lib.gpe_launch_count() - launches0
    sampler.stop_flag = True
    sampler.join()
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    total_ms = evs[0].elapsed_time(evs[-1])
    if world > 1:
        total_ms = sharding.max_over_ranks(total_ms, device=dev)
    # cheap correctness tripwire inside the bench: the prefix against the oracle
    mu_head = out["mu"][:512].cpu().numpy(); var_head = out["var"][:512].cpu().numpy()
    t_head = testing[:512].cpu().numpy()
    del out
    torch.cuda.empty_cache()

    # ---- end to end through the public API with host buffers ------------------------------------
    Ne = int(args.e2e_points)
    host_in = torch.rand(Ne, D, dtype=torch.float64).pin_memory().numpy()
    # result buffers are allocated once, page-locked, outside the timed region (a fresh 2 GB pinned allocation
    # costs more than the whole step); every step still moves all inputs H2D and all results D2H
    host_out = {"mu": torch.empty(Ne, dtype=torch.float64).pin_memory().numpy(),
                "var": torch.empty(Ne, dtype=torch.float64).pin_memory().numpy(),
                "deriv": torch.empty(Ne, D, dtype=torch.float64).pin_memory().numpy()}
    for _ in range(max(1, min(args.warmup, 2))):
        gp.predict(host_in, out=host_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        mu_h, var_h, der_h = gp.predict(host_in, out=host_out)   # H2D + kernels + D2H, returns numpy arrays
        _ = float(mu_h[-1]) + float(der_h[-1, -1])               # the results are on the host
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        e2e_s = sharding.max_over_ranks(e2e_s, device=dev)
    # the same call the way a reference user makes it: plain (pageable) numpy in, fresh numpy arrays out
    pageable_rate = None
    if world == 1:
        Np = min(Ne, 2_000_000)
        plain_in = np.array(host_in[:Np])
        for _ in range(4):   # repeated calls of one size get page-locked result arrays from the 2nd call on (engine._auto_pin);
            mu_p, var_p, der_p = gp.predict(plain_in)   # let torch's pinned cache fill before timing the steady state
        t0 = time.perf_counter()
        for _ in range(3):
            mu_p, var_p, der_p = gp.predict(plain_in)
        pageable_rate = 3 * Np / (time.perf_counter() - t0)
        del plain_in, mu_p, var_p, der_p

    # ---- opt-in variants of the same workload, reported beside the headline (short, outside every timed region) --
    variants = None
    if rank == 0:
        def _rate(fn, n_pts, reps=3):
            fn(); fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return n_pts * reps / (a.elapsed_time(b) * 1e-3)
        nv = min(N, 20_000_000)
        tv = testing[:nv]
        sym = gpe.DeviceModel(model["inputs"], model["theta"], model["invQt"], model["invQ"], device=local_rank,
                              symmetric_variance=True)
        t32 = tv.to(torch.float32)
        nh = min(nv, 4_000_000)   # the (N, D, D) Hessian output is 800 B per point
        hout = {"hess": torch.empty(nh, D, D, dtype=torch.float64, device=tv.device)}
        variants = {
            "points": nv,
            "fp64_symmetric_variance_points_per_s": _rate(lambda: sym.predict(tv), nv),
            "fp32_3xtf32_tcgen05_points_per_s": _rate(lambda: dm.predict_f32(t32), nv),
            "fp32_1xtf32_tcgen05_points_per_s": _rate(lambda: dm.predict_f32(t32, fast=True), nv),
            "fp64_mean_gradient_only_points_per_s": _rate(lambda: dm.predict(tv, want_var=False), nv),
            "fp64_with_hessian_points_per_s": _rate(lambda: dm.predict(tv[:nh], want_hess=True, out=hout), nh),
            "note": "same model and test points; symmetric = opt-in upper-triangular fold of invQ (exact identity, "
                    "half the DMMAs); with_hessian = mean + variance + gradient + (N, D, D) Hessian in one fused launch; tf32 = single precision with the variance contraction on tcgen05/TMEM (3x split: var error "
                    "~3e-6, meets the reference FP32 bar 1e-5; 1x: ~6e-5)",
        }
        del sym, t32, tv, hout
        # the caller on the input side of the path: cost + gradient of the hyper-parameter fit, one
        # (theta, target) problem per SM, timed through the C ABI with host buffers (wall clock, synchronous call)
        from gp_emulator_b200.training import DeviceTrainer
        rs_t = np.random.RandomState(1)
        nprob = torch.cuda.get_device_properties(local_rank).multi_processor_count
        trainer = DeviceTrainer(model["inputs"], np.sin(model["inputs"].sum(axis=1)), device=local_rank)
        thetas = 5.0 * (rs_t.random_sample((nprob, D + 2)) - 0.5)
        trainer.evaluate(thetas)
        t0 = time.perf_counter()
        for _ in range(5):
            trainer.evaluate(thetas)
        variants["training_objective_evaluations_per_s"] = 5 * nprob / (time.perf_counter() - t0)
        variants["training_note"] = ("loglikelihood + partial_devs (GaussianProcess.py:78-125) for %d thetas per launch, "
                                     "M=%d D=%d: block Gauss-Jordan on DMMA, k_train_eval" % (nprob, M, D))
        trainer.close()
    if rank == 0:
        from oracle import gp_oracle as orc
        mu_o, var_o, _ = orc.predict(model["inputs"], model["theta"], model["invQ"], model["invQt"], t_head)
        parity = {"mu": orc.ref_err(mu_head, mu_o), "var": orc.ref_err(var_head, var_o)}
        value = N * world * args.steps / (total_ms * 1e-3)
        kern_ms = float(np.mean(step_ms))
        achieved = N * F_PER_POINT / (kern_ms * 1e-3) / 1e12
        peak = peaks["dmma_tflops"]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                traffic = tj["dram_bytes_per_point"] * N
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(N),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "note": "FP64: achieved = N*139011 flop / mean launch time (CUDA events, one launch per step); "
                                 "peak = DMMA.8x8x4 rate measured live on this GPU (MEASURED_PEAKS.json has no FP64 "
                                 "figure); DFMA peak %.1f, SM %.0f MHz under FP64 load" %
                                 (peaks["dfma_tflops"], peaks["sm_mhz_fp64_load"]),
                         "kernel": "k_predict_full<4,8,2,4,10>", "kernel_ms": kern_ms,
                         "hbm_GBps": N * BYTES_PER_POINT / (kern_ms * 1e-3) / 1e9},
            "e2e": {"value": Ne * world * args.steps / e2e_s, "unit": "points/s",
                    "h2d_bytes_per_step": Ne * D * 8, "d2h_bytes_per_step": Ne * (2 + D) * 8,
                    "points_per_gpu_per_step": Ne,
                    "pageable_numpy_points_per_s": pageable_rate,   # pageable inputs, fresh result arrays per call (steady state)
                    "host_numa_binding": ("rank pinned to the %d CPUs local to its GPU" % len(numa_cpus)) if numa_cpus else None,
                    "api": "GaussianProcess.predict(numpy pinned in, preallocated pinned out) -> libgpemu two-slot stream pipeline"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "parity_vs_oracle": parity,
            "variants": variants,
        }
        if not args.no_cpu_baseline and world == 1:   # reported on rank 0 at N=1 only
            line["cpu_baseline"] = cpu_baseline(model, int(args.cpu_points))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
