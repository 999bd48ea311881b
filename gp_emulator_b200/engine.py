"""Handles on device-resident models: thin, typed wrappers over the libgpemu C ABI.

``DeviceModel`` is one trained GP on one GPU, ``DeviceBank`` is E GPs that share training inputs and test
points (the per-PC emulators of a MultivariateEmulator, or a per-band bank).  Test points may be

* numpy arrays (host): the library streams them through its chunked host pipeline and the results
  come back as numpy arrays; or
* torch CUDA tensors (device): the call is asynchronous on torch's current stream and the results are
  torch CUDA tensors -- PyTorch is used only to own the device buffers.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import WANT_DERIV, WANT_HESS, WANT_MU, WANT_VAR, HOST_PTRS, GpemuError, addr, check, f64c


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _pinned_empty(shape):
    """numpy array backed by torch's caching pinned-host allocator (the view keeps the block alive)."""
    import torch
    return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()


# Result arrays of host calls.  A fresh pageable numpy array costs page faults + a CPU copy out of the staging buffer
# every call (the kernel zero-fills its pages: ~26 GB/s on the GPU box against 52 GB/s for warm pages); a page-locked
# one costs ~0.7 s/GB to create but is then recycled by torch's caching host allocator and lets the D2H copy land in
# it directly.  ``pinned=None`` (the default) therefore switches to page-locked results from the second request of
# the same total size on, for sizes where it pays and stays bounded.
_AUTO_PIN_MIN, _AUTO_PIN_MAX = 1 << 20, 1 << 30
_pin_requests = {}


def _auto_pin(nbytes):
    if not (_AUTO_PIN_MIN <= nbytes <= _AUTO_PIN_MAX):
        return False
    if len(_pin_requests) > 512:
        _pin_requests.clear()
    seen = _pin_requests.get(nbytes, 0)
    _pin_requests[nbytes] = seen + 1
    return seen >= 1


def _current_stream_ptr(device_index):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device_index).cuda_stream)


class DeviceModel:
    """One GP resident on one device.  Mirrors the state ``GaussianProcess.predict`` reads
    (reference gp_emulator/GaussianProcess.py:228-249): inputs (M, D), theta (D+2), invQ (M, M), invQt (M)."""

    def __init__(self, inputs, theta, invQt, invQ=None, device=0, symmetric_variance=False):
        """``symmetric_variance=True`` opts into evaluating k^T invQ k through the upper-triangular fold of invQ
        (exact identity for any invQ, half the tensor-core work, rounding differs at the 1e-16 level);
        ``"auto"`` folds only when ``invQ`` is symmetric to 1e-6 of its largest entry -- what the inverse of a
        covariance matrix is (``_prepare_likelihood``: ~3e-10), unlike the random ``invQ`` of the reference benchmark."""
        inputs = f64c(inputs)
        if inputs.ndim != 2:
            raise ValueError("inputs must be (M, D)")
        self.M, self.D = inputs.shape
        theta = f64c(theta).ravel()
        if theta.size < self.D + 1:
            raise ValueError(f"theta must hold at least D+1 = {self.D + 1} entries, got {theta.size}")
        expx = np.exp(theta[: self.D + 1])
        invQt = f64c(invQt).ravel()
        if invQt.size != self.M:
            raise ValueError("invQt must have M entries")
        if invQ is not None:
            invQ = f64c(invQ)
            if invQ.shape != (self.M, self.M):
                raise ValueError("invQ must be (M, M)")
        self.device = int(device)
        self.has_var = invQ is not None
        h = C.c_void_p()
        if isinstance(symmetric_variance, str):
            if symmetric_variance != "auto":
                raise ValueError('symmetric_variance must be True, False or "auto"')
            symmetric_variance = (invQ is not None and self.M > 1
                                  and float(np.max(np.abs(invQ - invQ.T))) <= 1e-6 * float(np.max(np.abs(invQ))))
        self.symmetric_variance = bool(symmetric_variance)
        check(_lib.load().gpe_model_create_ex(self.device, self.M, self.D, addr(inputs), addr(expx), addr(invQt),
                                              addr(invQ), _lib.OPT_SYMMETRIC_VARIANCE if self.symmetric_variance else 0,
                                              C.byref(h)))
        self._h = h
        self._fin = weakref.finalize(self, _lib.load().gpe_model_destroy, h)

    def close(self):
        self._fin()

    def predict(self, testing, want_var=True, want_deriv=True, want_hess=False, want_mu=True, out=None,
                pinned=None):
        """Returns a dict with the requested arrays among mu (N,), var (N,), deriv (N, D), hess (N, D, D).

        ``out`` may hold preallocated result arrays under the same keys (float64, C-contiguous, right shape;
        numpy for host calls, CUDA tensors for device calls): they are filled in place and returned.  Fresh host
        results are ordinary numpy arrays; ``pinned`` chooses their memory: False pageable (staged through the
        library's page-locked buffers), True page-locked (the D2H copy lands in them directly; ~0.7 s per GB the
        first time a size is seen, then recycled by torch's caching allocator), None (default) page-locked from the
        second request of the same size on -- repeated calls, as in an optimisation loop, then run at the pinned rate.
        """
        lib = _lib.load()
        D = self.D
        if want_var and not self.has_var:
            raise GpemuError("variance requested but the model was uploaded without invQ")
        flags = (WANT_MU if want_mu else 0) | (WANT_VAR if want_var else 0) | (WANT_DERIV if want_deriv else 0) | (
            WANT_HESS if want_hess else 0)
        wanted = [k for k, w in (("mu", want_mu), ("var", want_var), ("deriv", want_deriv), ("hess", want_hess)) if w]
        out = dict(out) if out else {}
        if _is_torch(testing):
            import torch
            t = testing
            if not t.is_cuda or t.device.index != self.device:
                raise ValueError(f"torch test points must live on cuda:{self.device}")
            if t.dtype != torch.float64 or t.dim() != 2 or t.shape[1] != D:
                raise ValueError(f"testing must be float64 (N, {D})")
            t = t.contiguous()
            N = t.shape[0]
            shapes = {"mu": (N,), "var": (N,), "deriv": (N, D), "hess": (N, D, D)}
            for k in wanted:
                if k not in out:
                    out[k] = torch.empty(shapes[k], dtype=torch.float64, device=t.device)
                elif (tuple(out[k].shape) != shapes[k] or out[k].dtype != torch.float64 or not out[k].is_cuda
                      or not out[k].is_contiguous()):
                    raise ValueError(f"out[{k!r}] must be a contiguous float64 CUDA tensor of shape {shapes[k]}")
            check(lib.gpe_predict(self._h, addr(t), N, addr(out.get("mu")), addr(out.get("var")),
                                  addr(out.get("deriv")), addr(out.get("hess")), flags,
                                  _current_stream_ptr(self.device)))
            return {k: out[k] for k in wanted}
        t = f64c(testing)
        if t.ndim != 2 or t.shape[1] != D:
            raise ValueError(f"testing must be (N, {D})")
        N = t.shape[0]
        shapes = {"mu": (N,), "var": (N,), "deriv": (N, D), "hess": (N, D, D)}
        missing = [k for k in wanted if k not in out]
        if pinned is None:
            per_point = {"mu": 1, "var": 1, "deriv": D, "hess": D * D}
            pinned = bool(missing) and _auto_pin(8 * N * sum(per_point[k] for k in missing))
        mk = _pinned_empty if pinned else np.empty
        for k in wanted:
            if k not in out:
                out[k] = mk(shapes[k])
            elif (not isinstance(out[k], np.ndarray) or out[k].shape != shapes[k] or out[k].dtype != np.float64
                  or not out[k].flags.c_contiguous):
                raise ValueError(f"out[{k!r}] must be a C-contiguous float64 numpy array of shape {shapes[k]}")
        check(lib.gpe_predict(self._h, addr(t), N, addr(out.get("mu")), addr(out.get("var")),
                              addr(out.get("deriv")), addr(out.get("hess")), flags | HOST_PTRS, None))
        return {k: out[k] for k in wanted}


    def predict_f32(self, testing, want_var=True, want_deriv=True, fast=None):
        """Single-precision prediction on the tcgen05 / TMEM path (M <= 256): float32 in, float32 out.

        numpy float32 (N, D) -> numpy results; torch float32 CUDA tensor -> torch results (asynchronous).
        The variance contraction runs on the tensor cores: 3xTF32 split by default (FP32-grade, meets the
        reference's 1e-5 FP32 criterion); ``fast=True`` uses a single TF32 pass (error ~1e-4, ~1.7x faster).
        ``fast=None`` (default) = split for M <= 256, single pass above (where FP32 accumulation dominates the
        error anyway); ``fast=False`` forces the split for any M.
        """
        lib = _lib.load()
        D = self.D
        if want_var and not self.has_var:
            raise GpemuError("variance requested but the model was uploaded without invQ")
        flags = WANT_MU | (WANT_VAR if want_var else 0) | (WANT_DERIV if want_deriv else 0) | (
            _lib.F32_FAST_TF32 if fast else (_lib.F32_FORCE_3X if fast is False else 0))
        if _is_torch(testing):
            import torch
            t = testing
            if not t.is_cuda or t.device.index != self.device or t.dtype != torch.float32 or t.dim() != 2 \
                    or t.shape[1] != D:
                raise ValueError(f"testing must be a float32 (N, {D}) tensor on cuda:{self.device}")
            t = t.contiguous()
            N = t.shape[0]
            mk = lambda *s: torch.empty(*s, dtype=torch.float32, device=t.device)
            out = {"mu": mk(N)}
            if want_var: out["var"] = mk(N)
            if want_deriv: out["deriv"] = mk(N, D)
            check(lib.gpe_predict_f32(self._h, addr(t), N, addr(out["mu"]), addr(out.get("var")),
                                      addr(out.get("deriv")), flags, _current_stream_ptr(self.device)))
            return out
        t = np.ascontiguousarray(testing, dtype=np.float32)
        if t.ndim != 2 or t.shape[1] != D:
            raise ValueError(f"testing must be (N, {D})")
        N = t.shape[0]
        out = {"mu": np.empty(N, dtype=np.float32)}
        if want_var: out["var"] = np.empty(N, dtype=np.float32)
        if want_deriv: out["deriv"] = np.empty((N, D), dtype=np.float32)
        check(lib.gpe_predict_f32(self._h, addr(t), N, addr(out["mu"]), addr(out.get("var")),
                                  addr(out.get("deriv")), flags | HOST_PTRS, None))
        return out


class MultiDeviceModel:
    """One GP resident on several GPUs of the box; ``predict`` splits a host batch into contiguous ranges, one
    per device, driven by one host thread each inside the library (no Python threads, no collective)."""

    def __init__(self, inputs, theta, invQt, invQ=None, devices=None, symmetric_variance=False):
        inputs = f64c(inputs)
        self.M, self.D = inputs.shape
        theta = f64c(theta).ravel()
        expx = np.exp(theta[: self.D + 1])
        invQt = f64c(invQt).ravel()
        invQ = None if invQ is None else f64c(invQ)
        lib = _lib.load()
        if devices is None:
            devices = list(range(lib.gpe_device_count()))
        self.devices = [int(d) for d in devices]
        if not self.devices:
            raise GpemuError("no CUDA device available; gp_emulator_b200 has no CPU fallback")
        self.has_var = invQ is not None
        arr = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        check(lib.gpe_multi_create(len(self.devices), arr, self.M, self.D, addr(inputs), addr(expx), addr(invQt),
                                   addr(invQ), _lib.OPT_SYMMETRIC_VARIANCE if symmetric_variance else 0, C.byref(h)))
        self._h = h
        self._fin = weakref.finalize(self, lib.gpe_multi_destroy, h)

    def close(self):
        self._fin()

    def predict(self, testing, want_var=True, want_deriv=True, want_hess=False, out=None, pinned=False):
        t = f64c(testing)
        if t.ndim != 2 or t.shape[1] != self.D:
            raise ValueError(f"testing must be (N, {self.D})")
        if want_var and not self.has_var:
            raise GpemuError("variance requested but the model was uploaded without invQ")
        N, D = t.shape
        shapes = {"mu": (N,), "var": (N,), "deriv": (N, D), "hess": (N, D, D)}
        wanted = ["mu"] + [k for k, w in (("var", want_var), ("deriv", want_deriv), ("hess", want_hess)) if w]
        out = dict(out) if out else {}
        mk = _pinned_empty if pinned else np.empty
        for k in wanted:
            if k not in out:
                out[k] = mk(shapes[k])
        flags = WANT_MU | (WANT_VAR if want_var else 0) | (WANT_DERIV if want_deriv else 0) | (
            WANT_HESS if want_hess else 0)
        check(_lib.load().gpe_multi_predict(self._h, addr(t), N, addr(out.get("mu")), addr(out.get("var")),
                                            addr(out.get("deriv")), addr(out.get("hess")), flags))
        return {k: out[k] for k in wanted}


class DeviceBank:
    """E GPs sharing training inputs and test points; optional PCA basis (E, W) for back-projection."""

    def __init__(self, inputs, thetas, invQts, invQs=None, basis=None, device=0):
        inputs = f64c(inputs)
        self.M, self.D = inputs.shape
        thetas = f64c(thetas)
        if thetas.ndim != 2 or thetas.shape[1] < self.D + 1:
            raise ValueError("thetas must be (E, >= D+1)")
        self.E = thetas.shape[0]
        expx = f64c(np.exp(thetas[:, : self.D + 1]))
        invQts = f64c(invQts)
        if invQts.shape != (self.E, self.M):
            raise ValueError("invQts must be (E, M)")
        if invQs is not None:
            invQs = f64c(invQs)
            if invQs.shape != (self.E, self.M, self.M):
                raise ValueError("invQs must be (E, M, M)")
        self.W = 0
        if basis is not None:
            basis = f64c(basis)
            if basis.ndim != 2 or basis.shape[0] != self.E:
                raise ValueError("basis must be (E, W)")
            self.W = basis.shape[1]
        self.device = int(device)
        self.has_var = invQs is not None
        h = C.c_void_p()
        check(_lib.load().gpe_bank_create(self.device, self.E, self.M, self.D, addr(inputs), addr(expx),
                                          addr(invQts), addr(invQs), addr(basis), self.W, C.byref(h)))
        self._h = h
        self._fin = weakref.finalize(self, _lib.load().gpe_bank_destroy, h)

    def close(self):
        self._fin()

    def predict(self, testing, want_var=True, want_deriv=True, want_hess=False, project=False,
                project_deriv=False):
        """Point-major outputs: mu (N, E), var (N, E), deriv (N, E, D), hess (N, E, D, D);
        with ``project``: fwd (N, W) = mu @ basis; with ``project_deriv``: deriv_full (N, D, W).
        numpy in -> numpy out (copied through torch device buffers); torch CUDA in -> torch CUDA out."""
        import torch
        lib = _lib.load()
        D, E = self.D, self.E
        if want_var and not self.has_var:
            raise GpemuError("variance requested but the bank was uploaded without invQ")
        as_numpy = not _is_torch(testing)
        dev = torch.device("cuda", self.device)
        if as_numpy:
            th = f64c(testing)
            if th.ndim != 2 or th.shape[1] != D:
                raise ValueError(f"testing must be float64 (N, {D})")
            # host caller: bound the device footprint (fwd / deriv_full are W and D*W doubles per point) by walking
            # the batch in chunks; results land directly in the numpy arrays that are returned
            N = th.shape[0]
            per_point = 8 * (E * (1 + (1 if want_var else 0) + (D if (want_deriv or project_deriv) else 0)
                                  + (D * D if want_hess else 0))
                             + (self.W if project else 0) + (D * self.W if project_deriv else 0))
            chunk = max(1, min(N, self.host_chunk_bytes // max(per_point, 1)))
            if N > chunk:
                res = None
                for s0 in range(0, N, chunk):
                    part = self._predict_device(torch.from_numpy(th[s0:s0 + chunk]).to(dev), want_var, want_deriv,
                                                want_hess, project, project_deriv)
                    if res is None:
                        res = {k: np.empty((N,) + tuple(v.shape[1:])) for k, v in part.items()}
                    for k, v in part.items():
                        torch.from_numpy(res[k][s0:s0 + chunk]).copy_(v)
                    del part
                return res
            out = self._predict_device(torch.from_numpy(th).to(dev), want_var, want_deriv, want_hess, project,
                                       project_deriv)
            return {k: v.cpu().numpy() for k, v in out.items()}
        t = testing.contiguous()
        if t.dim() != 2 or t.shape[1] != D or t.dtype != torch.float64:
            raise ValueError(f"testing must be float64 (N, {D})")
        return self._predict_device(t, want_var, want_deriv, want_hess, project, project_deriv)

    def cost(self, testing, obs, weights=None, want_grad=True):
        """Least-squares misfit of the bank's means against observations, reduced over the emulators on the device:
        ``cost (N,) = 1/2 sum_e w_e (mu_ne - obs_ne)^2`` and ``grad (N, D) = sum_e w_e (mu_ne - obs_ne) deriv_ned``
        (``gpe_bank_cost``).  ``obs`` is (E,) -- one observation for every point -- or (N, E); ``weights`` (E,) or None.
        numpy in -> numpy out; torch CUDA in -> torch CUDA out (asynchronous on the current stream)."""
        import torch
        D, E = self.D, self.E
        as_numpy = not _is_torch(testing)
        dev = torch.device("cuda", self.device)
        if as_numpy:
            th = f64c(testing)
            if th.ndim != 2 or th.shape[1] != D:
                raise ValueError(f"testing must be float64 (N, {D})")
            t = torch.from_numpy(th).to(dev)
        else:
            t = testing.contiguous()
            if t.dim() != 2 or t.shape[1] != D or t.dtype != torch.float64:
                raise ValueError(f"testing must be float64 (N, {D})")
        N = t.shape[0]
        o = torch.as_tensor(np.asarray(obs, dtype=np.float64) if not _is_torch(obs) else obs, dtype=torch.float64,
                            device=dev).contiguous()
        if tuple(o.shape) == (E,):
            obs_ld = 0
        elif tuple(o.shape) == (N, E):
            obs_ld = E
        else:
            raise ValueError(f"obs must be ({E},) or ({N}, {E})")
        wt = None
        if weights is not None:
            wt = torch.as_tensor(np.asarray(weights, dtype=np.float64) if not _is_torch(weights) else weights,
                                 dtype=torch.float64, device=dev).contiguous()
            if tuple(wt.shape) != (E,):
                raise ValueError(f"weights must be ({E},)")
        out = {"cost": torch.empty(N, dtype=torch.float64, device=dev)}
        if want_grad:
            out["grad"] = torch.empty(N, D, dtype=torch.float64, device=dev)
        check(_lib.load().gpe_bank_cost(self._h, addr(t), N, addr(o), obs_ld, addr(wt), addr(out["cost"]),
                                        addr(out.get("grad")), _current_stream_ptr(self.device)))
        if as_numpy:
            return {k: v.cpu().numpy() for k, v in out.items()}
        return out

    host_chunk_bytes = 1 << 30   # device bytes of results per chunk when the caller passes numpy arrays

    def forward(self, testing, want_deriv=True):
        """numpy (N, D) -> fwd (N, W) [, deriv_full (N, D, W)] through ``gpe_bank_forward``: one call, one
        synchronisation, nothing but the spectra (and Jacobians) crosses PCIe.  This is what
        ``MultivariateEmulator.predict`` (reference multivariate_gp.py:195-222) runs on."""
        if self.W == 0:
            raise GpemuError("bank has no basis functions to project onto")
        t = f64c(testing)
        if t.ndim != 2 or t.shape[1] != self.D:
            raise ValueError(f"testing must be (N, {self.D})")
        N = t.shape[0]
        fwd = np.empty((N, self.W))
        dfull = np.empty((N, self.D, self.W)) if want_deriv else None
        check(_lib.load().gpe_bank_forward(self._h, addr(t), N, addr(fwd), addr(dfull)))
        return (fwd, dfull) if want_deriv else fwd

    def _predict_device(self, t, want_var, want_deriv, want_hess, project, project_deriv):
        import torch
        lib = _lib.load()
        D, E = self.D, self.E
        dev = t.device
        N = t.shape[0]
        mk = lambda *s: torch.empty(*s, dtype=torch.float64, device=dev)
        out = {"mu": mk(N, E)}
        flags = WANT_MU
        if want_var: out["var"] = mk(N, E); flags |= WANT_VAR
        if want_deriv or project_deriv: out["deriv"] = mk(N, E, D); flags |= WANT_DERIV
        if want_hess: out["hess"] = mk(N, E, D, D); flags |= WANT_HESS
        st = _current_stream_ptr(self.device)
        check(lib.gpe_bank_predict(self._h, addr(t), N, addr(out["mu"]), addr(out.get("var")),
                                   addr(out.get("deriv")), addr(out.get("hess")), flags, st))
        if project or project_deriv:
            if self.W == 0:
                raise GpemuError("bank has no basis functions to project onto")
            if project: out["fwd"] = mk(N, self.W)
            if project_deriv: out["deriv_full"] = mk(N, D, self.W)
            check(lib.gpe_bank_project(self._h, addr(out["mu"]), addr(out.get("deriv")), N, addr(out.get("fwd")),
                                       addr(out.get("deriv_full")), st))
        return out
