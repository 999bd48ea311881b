"""Secondary BASELINE.json configurations (3, 4, 5) on one GPU: device-resident throughput + parity spot checks.
Writes one JSON object per configuration (profiles/r01_configs.json is a saved copy of this output)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc

peaks = g.measure_fp64_peaks(0)
P64 = peaks["dmma_tflops"]


def timeit(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3, r


def F(M, D):
    return 2 * M * M + M * (5 * D + 6) + D + 1


out = []
# ---- config 3: M = 1000, variance dominated -----------------------------------------------------------------
M, D, N = 1000, 10, 1_000_000
inputs, theta, invQ, invQt, testing = orc.make_S_model(M, D, 2000, seed=3)
m = g.DeviceModel(inputs, theta, invQt, invQ)
o = m.predict(testing)
mu, var, deriv = orc.predict(inputs, theta, invQ, invQt, testing)
t = torch.rand(N, D, dtype=torch.float64, device="cuda")
s, _ = timeit(lambda: m.predict(t))
out.append({"config": "3: M=1000 D=10 FP64 mu+var+grad", "N": N, "points_per_s": N / s,
            "alg_tflops": N * F(M, D) / s / 1e12, "frac_of_dmma_peak": N * F(M, D) / s / 1e12 / P64,
            "parity": {"mu": orc.ref_err(o["mu"], mu), "var": orc.ref_err(o["var"], var), "deriv": orc.ref_err(o["deriv"], deriv)},
            })
t32d = torch.rand(N, D, dtype=torch.float32, device="cuda")
o32 = m.predict_f32(testing.astype(np.float32), fast=False)
mu32, var32, deriv32 = orc.predict(inputs, theta, invQ, invQt, testing.astype(np.float32).astype(np.float64))
s32, _ = timeit(lambda: m.predict_f32(t32d, fast=False))
s32f, _ = timeit(lambda: m.predict_f32(t32d, fast=True))
o32f = m.predict_f32(testing.astype(np.float32), fast=True)
out.append({"config": "3T: M=1000 D=10 FP32 on tcgen05 + TMEM (column passes): points_per_s = 3xTF32 (opt-in at this M), fast_points_per_s = 1xTF32 (default for M > 256)", "N": N,
            "points_per_s": N / s32, "fast_points_per_s": N / s32f, "fast_var_err": orc.ref_err(o32f["var"], var32), "tf32_tflops": N * 2 * 1024 * 1024 / s32 / 1e12,
            "parity_vs_fp64_oracle": {"mu": orc.ref_err(o32["mu"], mu32), "var": orc.ref_err(o32["var"], var32),
                                      "deriv": orc.ref_err(o32["deriv"], deriv32)}})
del t, t32d

# ---- config 1T: headline shape in single precision on tcgen05 -----------------------------------------------
M, D, N = 250, 10, 40_000_000
inputs, theta, invQ, invQt, testing = orc.make_S_model(M, D, 4000, seed=0)
m = g.DeviceModel(inputs, theta, invQt, invQ)
t32 = testing.astype(np.float32)
o = m.predict_f32(t32)
mu, var, deriv = orc.predict(inputs, theta, invQ, invQt, t32.astype(np.float64))
t = torch.rand(N, D, dtype=torch.float32, device="cuda")
s, _ = timeit(lambda: m.predict_f32(t))
sf, _ = timeit(lambda: m.predict_f32(t, fast=True))
of = m.predict_f32(t32, fast=True)
out.append({"config": "1T: M=250 D=10 FP32 on tcgen05 + TMEM: points_per_s = 3xTF32 (default for M <= 256), fast_points_per_s = 1xTF32", "N": N, "points_per_s": N / s,
            "fast_points_per_s": N / sf, "fast_var_err": orc.ref_err(of["var"], var),
            "tf32_tflops": N * 2 * 256 * 256 / s / 1e12,
            "parity_vs_fp64_oracle": {"mu": orc.ref_err(o["mu"], mu), "var": orc.ref_err(o["var"], var), "deriv": orc.ref_err(o["deriv"], deriv)}})
del t

# ---- config 4: MultivariateEmulator, P = 20 PCs, W = 2101 wavelengths ---------------------------------------
M, D, P, W, N = 250, 10, 20, 2101, 200_000
rs = np.random.RandomState(4)
inputs = rs.random_sample((M, D))
thetas = rs.random_sample((P, D + 2)); invQts = rs.random_sample((P, M)); invQs = rs.random_sample((P, M, M))
basis = np.linalg.qr(rs.standard_normal((W, P)))[0].T.copy()
bank = g.DeviceBank(inputs, thetas, invQts, invQs, basis=basis)
tt = rs.random_sample((64, D))
ob = bank.predict(tt, want_var=True, want_deriv=True, project=True, project_deriv=True)
models = [(inputs, thetas[i], invQs[i], invQts[i]) for i in range(P)]
fwd, mu_o, var_o, grad_o, dfull = orc.mv_predict_batch(models, basis, tt, want_deriv_full=True)
par4 = {"fwd": orc.ref_err(ob["fwd"], fwd), "pc_mu": orc.ref_err(ob["mu"], mu_o), "pc_var": orc.ref_err(ob["var"], var_o),
        "pc_grad": orc.ref_err(ob["deriv"], grad_o), "deriv_full": orc.ref_err(ob["deriv_full"], dfull)}
t = torch.rand(N, D, dtype=torch.float64, device="cuda")
s_pc, r = timeit(lambda: bank.predict(t, want_var=False, want_deriv=True))
s_all, r = timeit(lambda: bank.predict(t, want_var=False, want_deriv=True, project=True))
s_var, r = timeit(lambda: bank.predict(t, want_var=True, want_deriv=True))
proj_s = s_all - s_pc
# the BASELINE size: 1e7 test inputs, walked in chunks of N points (the 168 GB of spectra are produced chunk-wise and
# dropped, as SURVEY 8d prescribes); fresh random inputs are generated per chunk outside the timed kernels' stream order
def full_size(fn, n_chunk, n_total):
    reps = n_total // n_chunk
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
        del r
    e1.record(); torch.cuda.synchronize()
    return reps * n_chunk / (e0.elapsed_time(e1) * 1e-3)
full4 = full_size(lambda: bank.predict(t, want_var=False, want_deriv=True, project=True), N, 10_000_000)
out.append({"config": "4: MultivariateEmulator P=20 W=2101 M=250 D=10 FP64", "N": N,
            "full_size_1e7_points_chunked_points_per_s": full4,
            "pc_mean_grad_points_per_s": N / s_pc, "pc_mean_var_grad_points_per_s": N / s_var,
            "mean_grad_plus_backprojection_points_per_s": N / s_all,
            "backprojection_only": {"seconds": proj_s, "output_GBps": N * W * 8 / proj_s / 1e9,
                                    "fp64_tflops": N * 2 * P * W / proj_s / 1e12},
            "parity": par4})
del t, r

# ---- config 5: bank of 64 per-band GPs with gradient and Hessian ---------------------------------------------
M, D, E, N = 250, 10, 64, 100_000
thetas = rs.random_sample((E, D + 2)); invQts = rs.random_sample((E, M)); invQs = rs.random_sample((E, M, M))
bank = g.DeviceBank(inputs, thetas, invQts, invQs)
tt = rs.random_sample((40, D))
ob = bank.predict(tt, want_var=True, want_deriv=True, want_hess=True)
models = [(inputs, thetas[i], invQs[i], invQts[i]) for i in range(E)]
mu_o, var_o, grad_o, hess_o = orc.bank_predict(models, tt, do_hess=True)
par5 = {"mu": orc.ref_err(ob["mu"], mu_o), "var": orc.ref_err(ob["var"], var_o), "deriv": orc.ref_err(ob["deriv"], grad_o),
        "hess": orc.ref_err(ob["hess"], hess_o)}
t = torch.rand(N, D, dtype=torch.float64, device="cuda")
s5, r = timeit(lambda: bank.predict(t, want_var=True, want_deriv=True, want_hess=True), reps=2, warm=1)
Fh = F(M, D) + M * (D * D + 2 * D) + D
full5 = full_size(lambda: bank.predict(t, want_var=True, want_deriv=True, want_hess=True), N, 1_000_000)
out.append({"config": "5: bank of 64 GPs, shared test points, mu+var+grad+Hessian, M=250 D=10 FP64", "N": N,
            "chunked_1e6_points_points_per_s": full5,
            "points_per_s": N / s5, "emulator_points_per_s": N * E / s5, "alg_tflops": N * E * Fh / s5 / 1e12,
            "frac_of_dmma_peak": N * E * Fh / s5 / 1e12 / P64, "output_GBps": N * E * 112 * 8 / s5 / 1e9, "parity": par5})
# the same bank consumed on the fly: least-squares misfit against observed bands + its gradient (gpe_bank_cost)
Nc = 1_000_000
tc = torch.rand(Nc, D, dtype=torch.float64, device="cuda")
obs = torch.from_numpy(mu_o.mean(axis=0)).cuda()
c_o, g_o = orc.bank_cost(models, tt, mu_o.mean(axis=0))
oc = bank.cost(tt, mu_o.mean(axis=0))
s5c, r = timeit(lambda: bank.cost(tc, obs), reps=2, warm=1)
s5m, r = timeit(lambda: bank.predict(tc[:200_000], want_var=False, want_deriv=True), reps=2, warm=1)
out.append({"config": "5b: the same bank reduced on the fly to cost + gradient per point (gpe_bank_cost)", "N": Nc,
            "points_per_s": Nc / s5c, "emulator_points_per_s": Nc * E / s5c,
            "materialising_mean_gradient_points_per_s": 200_000 / s5m,
            "output_bytes_per_point": 8 * (1 + D), "instead_of_bytes_per_point": 8 * E * (1 + D),
            "parity": {"cost": orc.ref_err(oc["cost"], c_o), "grad": orc.ref_err(oc["grad"], g_o)}})
print(json.dumps({"fp64_peaks": peaks, "configs": out}, indent=1))
