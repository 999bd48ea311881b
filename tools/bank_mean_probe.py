"""Bank mean + gradient throughput: k_bank_mean (shared input differences) against the one-emulator kernel
(GPE_BANK_SHARED=off).  Usage: python tools/bank_mean_probe.py [E] [M] [D] [N]; prints emulator-points/s."""
import os, sys, subprocess
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

def run():
    import numpy as np, torch
    import gp_emulator_b200 as g
    E, M, D, N = [int(a) for a in sys.argv[2:6]]
    rs = np.random.RandomState(0)
    inputs = rs.random_sample((M, D)); thetas = rs.random_sample((E, D + 2)); invQts = rs.randn(E, M)
    bank = g.DeviceBank(inputs, thetas, invQts, None)
    t = torch.rand(N, D, dtype=torch.float64, device="cuda")
    for _ in range(3): bank.predict(t, want_var=False, want_deriv=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps): bank.predict(t, want_var=False, want_deriv=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-4s E=%d M=%d D=%d N=%d: %.3f ms, %.3e points/s, %.3e emulator-points/s" % (
        os.environ.get("GPE_BANK_SHARED", "on"), E, M, D, N, ms, N / ms * 1e3, N * E / ms * 1e3), flush=True)
    for _ in range(3): bank.predict(t, want_var=False, want_deriv=False)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps): bank.predict(t, want_var=False, want_deriv=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-4s means only:                  %.3f ms, %.3e points/s, %.3e emulator-points/s" % (
        os.environ.get("GPE_BANK_SHARED", "on"), ms, N / ms * 1e3, N * E / ms * 1e3), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        run()
    else:
        args = sys.argv[1:5] if len(sys.argv) >= 5 else ["20", "250", "10", "200000"]
        for mode in ("off", "on"):
            env = dict(os.environ, GPE_BANK_SHARED=mode)
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"] + args, env=env, check=True)
