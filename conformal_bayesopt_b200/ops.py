"""Thin functional layer over the C ABI: tensors in, tensors out, current CUDA stream.

No arithmetic of the hot path happens in Python/torch here -- every call lands in libcbo.so.
A CPU tensor is an error (there is no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import torch

from . import _lib
from .model_state import GPModelState

ACQ_KINDS = {"ucb": 0, "ei": 1, "nei": 2, "joint": 3}
_WORKSPACES: Dict[tuple, torch.Tensor] = {}
LAUNCH_COUNT = 0   # kernels launched through this module (bench.py reports it as gpu_launches)


@dataclass
class RatioWeights:
    """Device copies of RatioEstimator._target_network weights (ratio_estimation.py:73-86)."""
    W1: torch.Tensor
    b1: torch.Tensor
    W2: torch.Tensor
    b2: torch.Tensor
    W3: torch.Tensor
    b3: torch.Tensor

    @classmethod
    def from_sequential(cls, net, device) -> "RatioWeights":
        lin = [m for m in net if isinstance(m, torch.nn.Linear)]
        assert len(lin) == 3, "RatioEstimator network is d->H->H->1"
        f = lambda t: t.detach().to(device=device, dtype=torch.float64).contiguous()
        return cls(f(lin[0].weight), f(lin[0].bias), f(lin[1].weight), f(lin[1].bias), f(lin[2].weight), f(lin[2].bias))

    @classmethod
    def from_flat(cls, flat: torch.Tensor, d: int, hidden: int = 64) -> "RatioWeights":
        """Views into a flat parameter vector in torch's parameter order (the layout cbo_ratio_train_step_f64 updates in
        place): the acquisition kernels then always read the current target network, with no copies."""
        o, out = 0, []
        for n, shape in ((hidden * d, (hidden, d)), (hidden, (hidden,)), (hidden * hidden, (hidden, hidden)),
                         (hidden, (hidden,)), (hidden, (1, hidden)), (1, (1,))):
            out.append(flat[o:o + n].view(shape))
            o += n
        assert o == flat.numel()
        return cls(*out)

    def c_struct(self) -> "_lib.RatioMlpF64":
        return _lib.RatioMlpF64(self.W1.shape[0], self.W1.data_ptr(), self.b1.data_ptr(), self.W2.data_ptr(),
                                self.b2.data_ptr(), self.W3.data_ptr(), self.b3.data_ptr())


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.CboError("libcbo has no CPU path: tensors must live on a CUDA device")


def _f64(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None or (t.dtype == torch.float64 and t.is_contiguous()):
        return t
    return t.to(torch.float64).contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


@dataclass
class NeiBaseline:
    """Device arrays of the cNEI baseline (acquisitions.py:117-122), built once per acquisition object (see `build`)."""
    Xb: torch.Tensor     # (nb, d)
    KbX: torch.Tensor    # (nb, Np)
    Sbb: torch.Tensor    # (nb, nb)
    mub: torch.Tensor    # (nb,)

    @classmethod
    def build(cls, state: GPModelState, X_baseline: torch.Tensor) -> "NeiBaseline":
        """All of it on the library's own kernels (no cuBLAS / torch.linalg on the product path): k(X_b, X) and the posterior
        mean by cbo_kernel_rows_f64, V_b = k(X_b, X) L^-T and V_b V_b^T by cbo_gemm_tn_f64, k(X_b, X_b) by a second
        kernel-row call against a model state whose "training inputs" are the baseline points."""
        Xb = X_baseline.detach().to(state.device, torch.float64).reshape(-1, state.d).contiguous()
        nb = Xb.shape[0]
        KbX, mub = kernel_rows(state, Xb)                                   # (nb, Np) zero-padded, mean + k alpha
        Vb = gemm_tn(KbX, state.linv, tri=1)                                # (nb, Np): rows L^-1 k(X, x_b)
        nbp = (nb + 127) // 128 * 128
        Vpad = torch.zeros(nbp, state.Np, dtype=torch.float64, device=state.device)
        Vpad[:nb] = Vb
        VVt = gemm_tn(Vb, Vpad, tri=0)[:, :nb]                              # V_b V_b^T
        zeros = torch.zeros(nbp, dtype=torch.float64, device=state.device)
        pseudo = GPModelState(X=Xb, y=mub, mean=0.0, lengthscale=state.lengthscale, outputscale=state.outputscale,
                              noise=state.noise, alpha=zeros, linv=zeros, linv_t=zeros,       # factors: never read by kernel_rows
                              inv_lengthscale=state.inv_lengthscale)
        Kbb = kernel_rows(pseudo, Xb)[0][:, :nb]
        Kbb.diagonal().fill_(state.outputscale)
        return cls(Xb, KbX, (Kbb - VVt).contiguous(), mub.contiguous())

    def c_struct(self) -> "_lib.NeiBaselineF64":
        return _lib.NeiBaselineF64(self.Xb.shape[0], self.Xb.data_ptr(), self.KbX.data_ptr(), self.Sbb.data_ptr(),
                                   self.mub.data_ptr())


def make_params(*, alpha, temp, grid_res, q=1, num_mc=1, num_probes=8, max_grid_refinements=0, kind="ucb",
                optimistic=True, randomized=False, eps_per_candidate=False, beta_prime=0.0, best_f=0.0,
                nei=None, refine_reduce=None, grid_cov_scale=None, ext_adjoints=None, grid_grad_leak=False) -> "_lib.Params":
    """`refine_reduce`: a `_lib.REFINE_REDUCE_FN` instance (kept alive by the caller) or None."""
    fn = C.cast(refine_reduce, C.c_void_p) if refine_reduce is not None else None
    return _lib.Params(float(alpha), float(temp), int(grid_res), int(q), int(num_mc), int(num_probes),
                       int(max_grid_refinements), ACQ_KINDS[kind], int(bool(optimistic)), int(bool(randomized)),
                       int(bool(eps_per_candidate)), float(beta_prime), float(best_f),
                       C.pointer(nei) if nei is not None else None, fn, None, grid_cov_scale,
                       C.pointer(ext_adjoints) if ext_adjoints is not None else None, int(bool(grid_grad_leak)))


def workspace(state: GPModelState, b: int, params, chunk: int = 0, grad: bool = False, stream: Optional[int] = None) -> torch.Tensor:
    """Caller-owned scratch (torch caching allocator), cached per (device, stream)."""
    lib = _lib.load()
    fn = lib.cbo_workspace_bytes_grad if grad else lib.cbo_workspace_bytes
    nbytes = fn(state.N, state.d, b, C.byref(params), chunk)
    key = (state.device.index, stream if stream is not None else torch.cuda.current_stream().cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes + 256:           # 256 spare bytes: the pointer is aligned up by the callers
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=state.device)
        _WORKSPACES[key] = ws
    return ws


def count_launches(state: GPModelState, b: int, params, ratio: bool, chunk: int = 0, grad: bool = False,
                   extra_outputs: bool = False) -> int:
    """Kernels one cbo_conformal_acq[_grad]_f64 call enqueues (mirrors csrc/cbo_api.cu)."""
    C_ = b if params.max_grid_refinements > 0 else min(b, chunk if chunk > 0 else 65536)
    n_chunks = (b + C_ - 1) // C_
    small = b * params.q <= 32                     # latency path: split-K GEMMs (gemm + reduce [+ row norms])
    per_call = 9                                   # scale_transpose, pad_probes, 2 x (split-K gemm + reduce) for T, ratio/fill, eps_absmax, train_tables
    g1, g2 = (3, 2) if small else (1, 1)
    per_chunk = 1 + (1 if ratio else 0) + g1 + g2 + 1    # rows, ratio(cand), gemm pass 1, pass 2, score
    if grad:                                       # score value+reverse (replaces score unless extra outputs), 2 x gemm, rows reverse, ratio reverse
        per_chunk += (1 if extra_outputs else 0) + 2 * g2 + 1 + (1 if ratio else 0)
    return per_call + n_chunks * per_chunk


def conformal_acq(state: GPModelState, Xq: torch.Tensor, *, eps_grid: torch.Tensor, z: torch.Tensor,
                  eps_mc: torch.Tensor, u_rand: Optional[torch.Tensor] = None, ratio: Optional[RatioWeights] = None,
                  outputs: Sequence[str] = (), chunk: int = 0, want_grad: bool = False, **param_kw) -> Dict[str, torch.Tensor]:
    """cbo_conformal_acq_f64 (or cbo_conformal_acq_grad_f64 with `want_grad`: adds 'dacq_dX' (b, q, d)).
    Xq (b, q, d) or (b, d).  Returns {'acq': (b,), + requested outputs}."""
    global LAUNCH_COUNT
    lib = _lib.load()
    _require_cuda(Xq, eps_grid, z, eps_mc, u_rand, state.X)
    if Xq.ndim == 2:
        Xq = Xq.unsqueeze(1)
    b, q, _ = Xq.shape
    if q > 4:
        raise _lib.CboError("unsupported shape: q > 4")
    G = int(param_kw["grid_res"])
    Xq, eps_grid, z, eps_mc, u_rand = map(_f64, (Xq, eps_grid, z, eps_mc, u_rand))
    eps_per_candidate = eps_grid.numel() == b * G * q and eps_grid.numel() != G * q
    if eps_grid.numel() not in (G * q, b * G * q):
        raise _lib.CboError(f"eps_grid must hold G*q or b*G*q values, got {tuple(eps_grid.shape)}")
    reducer = param_kw.pop("refine_reduce", None)            # callable(list of 4 floats) -> list of 4 floats, or None
    reduce_c = None
    hook_error = []
    if reducer is not None:
        def _hook(ptr, _ctx, _f=reducer):
            # ctypes swallows exceptions raised inside a callback: keep the first one and re-raise it after the call
            try:
                vals = _f([ptr[0], ptr[1], ptr[2], ptr[3]])
                for k in range(4):
                    ptr[k] = float(vals[k])
            except BaseException as e:  # noqa: BLE001
                if not hook_error:
                    hook_error.append(e)
        reduce_c = _lib.REFINE_REDUCE_FN(_hook)              # kept alive until the call returns
    cov_scale = param_kw.pop("grid_cov_scale", None)        # (b, q) device tensor: initial row scales of the grid covariance
    if cov_scale is not None:
        cov_scale = _f64(cov_scale)
        _require_cuda(cov_scale)
        if cov_scale.numel() != b * q:
            raise _lib.CboError("grid_cov_scale must hold b*q scales")
    ext = param_kw.pop("ext_adjoints", None)                # (mask_bar (b,G,q), ood_bar (b,q), f_bar (b,S,G,q) or None): VJP mode
    ext_c = None
    if ext is not None:
        if not want_grad:
            raise _lib.CboError("ext_adjoints belong to the gradient call")
        ext = tuple(None if t is None else _f64(t) for t in ext)
        _require_cuda(*ext)
        ext_c = _lib.ExtAdjointsF64(*[_ptr(t) for t in ext])
    baseline = param_kw.pop("nei_baseline", None)
    nei_c = baseline.c_struct() if baseline is not None else None          # kept alive until the call returns
    per_sample = (baseline.Xb.shape[0] + q) if baseline is not None else q
    num_mc = eps_mc.numel() // (G * per_sample)
    params = make_params(q=q, num_mc=num_mc, num_probes=z.shape[-1], eps_per_candidate=eps_per_candidate, nei=nei_c,
                         refine_reduce=reduce_c, grid_cov_scale=_ptr(cov_scale), ext_adjoints=ext_c, **param_kw)
    if ext is not None:
        want = (b * G * q, b * q, b * num_mc * G * q)
        for t, n in zip(ext, want):
            if t is not None and t.numel() != n:
                raise _lib.CboError("ext_adjoints: mask_bar (b,G,q), ood_bar (b,q), f_bar (b,S,G,q)")
    if z.shape[0] != state.N + q:
        raise _lib.CboError(f"z must have N+q = {state.N + q} rows, got {z.shape[0]}")
    if u_rand is not None and u_rand.numel() != b * G * q:
        raise _lib.CboError("u_rand must hold b*G*q uniforms")
    dev = Xq.device
    out_t = {"acq": torch.empty(b, dtype=torch.float64, device=dev)}
    shapes = {"target_grid": (b, G, q), "conf_pred_mask": (b, G, q), "q_conf_scores": (b, G, q), "grid_logp": (b, G),
              "ood_mask": (b, q), "post_mean": (b, q), "num_accepted": (b, q), "post_var": (b,) if q == 1 else (b, q, q),
              "refit_mean": (b, G), "refit_sd": (b,)}
    if q > 1 and ("refit_mean" in outputs or "refit_sd" in outputs):
        raise _lib.CboError("refit_mean / refit_sd are q = 1 outputs")
    shapes["mc_values"] = (b, num_mc, G, q)             # kind nei / joint: the per-sample values the integration consumed
    if "mc_values" in outputs and baseline is None:
        raise _lib.CboError("mc_values is an output of the joint-sample kinds (nei, joint)")
    for name in outputs:
        out_t[name] = torch.empty(shapes[name], dtype=torch.float64, device=dev)
    cout = _lib.OutputsF64(**{k: _ptr(v) for k, v in out_t.items()})
    cs = _stream()
    ws = workspace(state, b, params, chunk, grad=want_grad, stream=cs)
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    gp_c = state.c_struct()
    ratio_c = ratio.c_struct() if ratio is not None else None
    ratio_p = C.byref(ratio_c) if ratio_c is not None else None
    if want_grad:
        dX = torch.empty(b, q, Xq.shape[-1], dtype=torch.float64, device=dev)
        st = lib.cbo_conformal_acq_grad_f64(C.byref(gp_c), ratio_p, C.byref(params), Xq.data_ptr(), b, eps_grid.data_ptr(),
                                            z.data_ptr(), eps_mc.data_ptr(), _ptr(u_rand), C.byref(cout), dX.data_ptr(),
                                            ws_ptr, ws.numel() - 256, chunk, cs)
        if hook_error:
            raise hook_error[0]
        _lib.check(st, "cbo_conformal_acq_grad_f64")
    else:
        st = lib.cbo_conformal_acq_f64(C.byref(gp_c), ratio_p, C.byref(params), Xq.data_ptr(), b, eps_grid.data_ptr(),
                                       z.data_ptr(), eps_mc.data_ptr(), _ptr(u_rand), C.byref(cout), ws_ptr,
                                       ws.numel() - 256, chunk, cs)
        if hook_error:
            raise hook_error[0]
        _lib.check(st, "cbo_conformal_acq_f64")
    LAUNCH_COUNT += count_launches(state, b, params, ratio is not None, chunk, grad=want_grad, extra_outputs=len(outputs) > 0)
    if q == 1:   # squeeze the q axis (the q = 1 layout of round-1 callers)
        out_t = {k: (v.squeeze(-1) if k in ("target_grid", "conf_pred_mask", "q_conf_scores", "ood_mask", "post_mean",
                                             "num_accepted") else v) for k, v in out_t.items()}
    if want_grad:
        out_t["dacq_dX"] = dX
    return out_t


def conformal_acq_host(state: GPModelState, Xq_host: torch.Tensor, acq_host: torch.Tensor, *, eps_grid, z, eps_mc,
                       u_rand=None, ratio: Optional[RatioWeights] = None, chunk: int = 0, staging=None, **param_kw):
    """cbo_conformal_acq_host_f64: HOST (pinned) candidates in, HOST acquisition values out; H2D and D2H inside."""
    global LAUNCH_COUNT
    lib = _lib.load()
    assert not Xq_host.is_cuda and not acq_host.is_cuda
    Xq_host = Xq_host.reshape(Xq_host.shape[0], -1)
    b, G = Xq_host.shape[0], int(param_kw["grid_res"])
    num_mc = eps_mc.reshape(-1, G).shape[0]
    params = make_params(num_mc=num_mc, num_probes=z.shape[-1], **param_kw)
    dev = state.device
    if staging is None:
        staging = (torch.empty(b, state.d, dtype=torch.float64, device=dev), torch.empty(b, dtype=torch.float64, device=dev))
    ws = workspace(state, b, params, chunk)
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    gp_c = state.c_struct()
    ratio_c = ratio.c_struct() if ratio is not None else None
    st = lib.cbo_conformal_acq_host_f64(C.byref(gp_c), C.byref(ratio_c) if ratio_c is not None else None,
                                        C.byref(params), Xq_host.data_ptr(), b, eps_grid.data_ptr(), z.data_ptr(),
                                        eps_mc.data_ptr(), _ptr(u_rand), acq_host.data_ptr(), staging[0].data_ptr(),
                                        staging[1].data_ptr(), ws_ptr, ws.numel() - 256, chunk, _stream())
    _lib.check(st, "cbo_conformal_acq_host_f64")
    LAUNCH_COUNT += count_launches(state, b, params, ratio is not None, chunk)
    return acq_host


def cehvi_integrate(per_output, eps_mc: torch.Tensor, cell_lower: torch.Tensor, cell_upper: torch.Tensor, *, alpha: float,
                    optimistic: bool) -> torch.Tensor:
    """cbo_cehvi_integrate_f64.  `per_output`: list (one per objective) of the dicts returned by `conformal_acq` with
    outputs conf_pred_mask, grid_logp, target_grid, post_mean, ood_mask, refit_mean, refit_sd.  eps_mc (S, G, m)."""
    global LAUNCH_COUNT
    lib = _lib.load()
    m = len(per_output)
    keys = ("conf_pred_mask", "grid_logp", "target_grid", "post_mean", "ood_mask", "refit_mean", "refit_sd")
    b, G = per_output[0]["conf_pred_mask"].shape
    arrs = []
    keep = []
    for kname in keys:
        ts = [_f64(o[kname]) for o in per_output]
        keep.append(ts)
        arrs.append((C.c_void_p * m)(*[t.data_ptr() for t in ts]))
    eps_mc, cell_lower, cell_upper = map(_f64, (eps_mc, cell_lower, cell_upper))
    _require_cuda(eps_mc, cell_lower, cell_upper)
    S = eps_mc.numel() // (G * m)
    n_cells = cell_lower.shape[0]
    acq = torch.empty(b, dtype=torch.float64, device=eps_mc.device)
    st = lib.cbo_cehvi_integrate_f64(m, b, G, S, *arrs, eps_mc.data_ptr(), cell_lower.data_ptr(), cell_upper.data_ptr(),
                                     n_cells, float(alpha), int(bool(optimistic)), acq.data_ptr(), _stream())
    _lib.check(st, "cbo_cehvi_integrate_f64")
    LAUNCH_COUNT += 1
    return acq


def cehvi_integrate_q(per_output, noise, eps_mc: Optional[torch.Tensor], cell_lower: torch.Tensor, cell_upper: torch.Tensor, *,
                      q: int, alpha: float, optimistic: bool, cells_per_sample: bool = False, want_values: bool = False):
    """cbo_cehvi_integrate_q_f64: joint batches q <= 4 and cNEHVI.  `per_output`: list (one per objective) of the dicts returned
    by `conformal_acq(..., outputs=(conf_pred_mask, grid_logp, target_grid, post_mean, ood_mask, post_var[, mc_values]))`.
    Entries that carry 'mc_values' (b, S, G, q) supply the samples themselves (cNEHVI), the others are sampled from the refit
    posterior with eps_mc (S, G, q, m).  cells: (n_cells, m) or, with `cells_per_sample`, (S, n_cells, m)."""
    lib = _lib.load()
    m = len(per_output)
    dev = per_output[0]["conf_pred_mask"].device
    b, G = per_output[0]["grid_logp"].shape
    keep = []
    def ptrs(key, shape):
        ts = [_f64(o[key]).reshape(shape) for o in per_output]
        keep.append(ts)
        return (C.c_void_p * m)(*[t.data_ptr() for t in ts])
    mask, grid = ptrs("conf_pred_mask", (b, G, q)), ptrs("target_grid", (b, G, q))
    logp, mu, ood = ptrs("grid_logp", (b, G)), ptrs("post_mean", (b, q)), ptrs("ood_mask", (b, q))
    ext = ["mc_values" in o for o in per_output]
    if any(ext) and not all(ext):
        raise _lib.CboError("either every output supplies mc_values or none does")
    fx = cov = None
    if all(ext):
        S = per_output[0]["mc_values"].shape[1]
        fx = ptrs("mc_values", (b, S, G, q))
    else:
        cov = ptrs("post_var", (b, q, q))
        eps_mc = _f64(eps_mc)
        _require_cuda(eps_mc)
        S = eps_mc.numel() // (G * q * m)
    cell_lower, cell_upper = _f64(cell_lower), _f64(cell_upper)
    _require_cuda(cell_lower, cell_upper)
    n_cells = cell_lower.shape[-2]
    if cells_per_sample and cell_lower.shape[0] != S:
        raise _lib.CboError("per-sample cells must be (S, n_cells, m)")
    noise_c = (C.c_double * m)(*[float(v) for v in noise])
    acq = torch.empty(b, dtype=torch.float64, device=dev)
    values = torch.empty(b, G, dtype=torch.float64, device=dev) if want_values else None
    st = lib.cbo_cehvi_integrate_q_f64(m, b, G, S, q, mask, logp, grid, mu, ood, cov, noise_c, fx, _ptr(eps_mc) if fx is None else None,
                                       cell_lower.data_ptr(), cell_upper.data_ptr(), n_cells, int(bool(cells_per_sample)),
                                       float(alpha), int(bool(optimistic)), acq.data_ptr(), _ptr(values), _stream())
    _lib.check(st, "cbo_cehvi_integrate_q_f64")
    return (acq, values) if want_values else acq


def cehvi_integrate_q_grad(per_output, noise, eps_mc: Optional[torch.Tensor], cell_lower: torch.Tensor, cell_upper: torch.Tensor, *,
                           q: int, alpha: float, optimistic: bool, cells_per_sample: bool = False):
    """cbo_cehvi_integrate_q_grad_f64: `cehvi_integrate_q` plus its reverse pass.  Returns (acq (b,), values (b,G),
    [(mask_bar (b,G,q), ood_bar (b,q), f_bar (b,S,G,q)) per output]) -- the `ext_adjoints` of the per-output gradient calls."""
    global LAUNCH_COUNT
    lib = _lib.load()
    m = len(per_output)
    dev = per_output[0]["conf_pred_mask"].device
    b, G = per_output[0]["grid_logp"].shape
    keep = []
    def ptrs(key, shape):
        ts = [_f64(o[key]).reshape(shape) for o in per_output]
        keep.append(ts)
        return (C.c_void_p * m)(*[t.data_ptr() for t in ts])
    mask, grid = ptrs("conf_pred_mask", (b, G, q)), ptrs("target_grid", (b, G, q))
    logp, mu, ood = ptrs("grid_logp", (b, G)), ptrs("post_mean", (b, q)), ptrs("ood_mask", (b, q))
    ext = ["mc_values" in o for o in per_output]
    if any(ext) and not all(ext):
        raise _lib.CboError("either every output supplies mc_values or none does")
    fx = cov = None
    if all(ext):
        S = per_output[0]["mc_values"].shape[1]
        fx = ptrs("mc_values", (b, S, G, q))
    else:
        cov = ptrs("post_var", (b, q, q))
        eps_mc = _f64(eps_mc)
        _require_cuda(eps_mc)
        S = eps_mc.numel() // (G * q * m)
    cell_lower, cell_upper = _f64(cell_lower), _f64(cell_upper)
    _require_cuda(cell_lower, cell_upper)
    n_cells = cell_lower.shape[-2]
    if cells_per_sample and cell_lower.shape[0] != S:
        raise _lib.CboError("per-sample cells must be (S, n_cells, m)")
    noise_c = (C.c_double * m)(*[float(v) for v in noise])
    acq = torch.empty(b, dtype=torch.float64, device=dev)
    values = torch.empty(b, G, dtype=torch.float64, device=dev)
    bars = [(torch.empty(b, G, q, dtype=torch.float64, device=dev), torch.empty(b, q, dtype=torch.float64, device=dev),
             torch.empty(b, S, G, q, dtype=torch.float64, device=dev)) for _ in range(m)]
    arr = lambda i: (C.c_void_p * m)(*[t[i].data_ptr() for t in bars])
    st = lib.cbo_cehvi_integrate_q_grad_f64(m, b, G, S, q, mask, logp, grid, mu, ood, cov, noise_c, fx,
                                            _ptr(eps_mc) if fx is None else None, cell_lower.data_ptr(), cell_upper.data_ptr(),
                                            n_cells, int(bool(cells_per_sample)), float(alpha), int(bool(optimistic)),
                                            acq.data_ptr(), values.data_ptr(), arr(0), arr(1), arr(2), _stream())
    _lib.check(st, "cbo_cehvi_integrate_q_grad_f64")
    LAUNCH_COUNT += 1
    return acq, values, bars


def conf_mask_to_bounds(target_grid: torch.Tensor, conf_pred_mask: torch.Tensor):
    """conf_mask_to_bounds (helpers.py:15-91): (b, G, 1, t) x2 -> lb, ub (b, 1, t)."""
    lib = _lib.load()
    _require_cuda(target_grid, conf_pred_mask)
    assert target_grid.shape == conf_pred_mask.shape and target_grid.shape[-2] == 1, "q = 1 only (helpers.py:35-37)"
    b, G, _, t = target_grid.shape
    grid = _f64(target_grid.reshape(b, G, t))
    mask = _f64(conf_pred_mask.reshape(b, G, t))
    lb = torch.empty(b, t, dtype=torch.float64, device=grid.device)
    ub = torch.empty_like(lb)
    st = lib.cbo_conf_mask_to_bounds_f64(grid.data_ptr(), mask.data_ptr(), b, G, t, lb.data_ptr(), ub.data_ptr(), _stream())
    _lib.check(st, "cbo_conf_mask_to_bounds_f64")
    return lb.unsqueeze(1), ub.unsqueeze(1)


def ratio_forward(ratio: RatioWeights, x: torch.Tensor) -> torch.Tensor:
    """RatioEstimator.forward (ratio_estimation.py:97-99) on (..., d) points."""
    lib = _lib.load()
    _require_cuda(x)
    shape = x.shape[:-1]
    x2 = _f64(x.reshape(-1, x.shape[-1]))
    out = torch.empty(x2.shape[0], dtype=torch.float64, device=x.device)
    rc = ratio.c_struct()
    st = lib.cbo_ratio_forward_f64(C.byref(rc), x2.data_ptr(), x2.shape[0], x2.shape[1], out.data_ptr(), _stream())
    _lib.check(st, "cbo_ratio_forward_f64")
    return out.reshape(shape)


def ratio_num_params(d: int, hidden: int = 64) -> int:
    return hidden * d + hidden + hidden * hidden + hidden + hidden + 1


def ratio_train_step(theta: torch.Tensor, adam_m: torch.Tensor, adam_v: torch.Tensor, target: torch.Tensor, *,
                     x_neg: torch.Tensor, x_pos: torch.Tensor, pos_idx: Optional[torch.Tensor], n_pos: int, step: int,
                     lr: float, weight_decay: float, ema_weight: float, betas=(0.9, 0.999), eps: float = 1e-8,
                     loss: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None):
    """One classifier training step + EMA, in place on the flat fp64 device vectors (ratio_estimation.py:133-144)."""
    lib = _lib.load()
    _require_cuda(theta, adam_m, adam_v, target, x_neg, x_pos, pos_idx, loss)
    d = x_neg.shape[-1]
    assert theta.numel() == ratio_num_params(d) and x_pos.shape[-1] == d
    for t in (theta, adam_m, adam_v, target, x_neg, x_pos):
        assert t.dtype == torch.float64 and t.is_contiguous()
    assert pos_idx is None or (pos_idx.dtype == torch.int64 and pos_idx.numel() >= n_pos)
    n_rows = x_neg.shape[0] + n_pos
    need = lib.cbo_ratio_train_workspace_bytes(n_rows, d)
    if workspace is None or workspace.numel() * 8 < need:
        workspace = torch.empty((need + 7) // 8, dtype=torch.float64, device=theta.device)
    st = lib.cbo_ratio_train_step_f64(d, 64, theta.data_ptr(), adam_m.data_ptr(), adam_v.data_ptr(), target.data_ptr(),
                                      x_neg.data_ptr(), x_neg.shape[0], x_pos.data_ptr(), _ptr(pos_idx), n_pos, step,
                                      lr, betas[0], betas[1], eps, weight_decay, ema_weight, _ptr(loss),
                                      workspace.data_ptr(), workspace.numel() * 8, _stream())
    _lib.check(st, "cbo_ratio_train_step_f64")
    return workspace


def sample_subset(n: int, k: int, seed: int, device, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """cbo_sample_subset_i64: k distinct pseudo-random indices out of [0, n) (the first k values of a keyed permutation), drawn
    on the device in one launch -- the minibatch draw of RatioEstimator.optimize_callback (ratio_estimation.py:129-131)."""
    global LAUNCH_COUNT
    lib = _lib.load()
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.CboError("sample_subset draws on the device (libcbo has no CPU path)")
    if out is None or out.numel() < k:
        out = torch.empty(max(k, 1), dtype=torch.int64, device=device)
    assert out.dtype == torch.int64 and out.is_contiguous() and out.device.type == "cuda"
    with torch.cuda.device(out.device):
        st = lib.cbo_sample_subset_i64(int(n), int(k), int(seed) & 0xFFFFFFFFFFFFFFFF, out.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "cbo_sample_subset_i64")
    LAUNCH_COUNT += 1
    return out[:k]


def sgld_update(X: torch.Tensor, momentum_buf: Optional[torch.Tensor], dacq_dX: torch.Tensor, noise: torch.Tensor,
                lb: torch.Tensor, ub: torch.Tensor, *, lr: float, momentum: float, temperature: float,
                penalty_scale: float = 1e2, penalty_width: float = 1e-2):
    """In-place SGLD / SGHMC update of the candidates with the box penalty (optim/optimize.py:224-235, optim/sgld.py:9-47)."""
    lib = _lib.load()
    _require_cuda(X, momentum_buf, dacq_dX, noise, lb, ub)
    d = X.shape[-1]
    for t in (X, momentum_buf, dacq_dX, noise, lb, ub):
        assert t is None or (t.dtype == torch.float64 and t.is_contiguous())
    assert dacq_dX.shape == X.shape and noise.shape == X.shape and lb.numel() == d and ub.numel() == d
    st = lib.cbo_sgld_update_f64(X.data_ptr(), _ptr(momentum_buf), dacq_dX.data_ptr(), noise.data_ptr(), lb.data_ptr(),
                                 ub.data_ptr(), X.numel() // d, d, lr, momentum, temperature, penalty_scale,
                                 penalty_width, _stream())
    _lib.check(st, "cbo_sgld_update_f64")


def gemm_tn(A: torch.Tensor, B: torch.Tensor, tri: int = 0, rownorm: bool = False):
    """C = A @ B^T restricted to the triangular K-range (test / profiling hook for the DMMA kernel)."""
    lib = _lib.load()
    _require_cuda(A, B)
    M, K = A.shape
    Nn = B.shape[0]
    Cm = torch.empty(M, Nn, dtype=torch.float64, device=A.device)
    rn = torch.empty(M, Nn // 128, dtype=torch.float64, device=A.device) if rownorm else None
    st = lib.cbo_gemm_tn_f64(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), Cm.data_ptr(), Cm.stride(0), M, Nn, K,
                             tri, _ptr(rn), _stream())
    _lib.check(st, "cbo_gemm_tn_f64")
    return (Cm, rn) if rownorm else Cm


def kernel_rows(state: GPModelState, Xq: torch.Tensor):
    lib = _lib.load()
    _require_cuda(Xq)
    Xq = _f64(Xq.reshape(-1, state.d))
    n = Xq.shape[0]
    U = torch.empty(n, state.Np, dtype=torch.float64, device=Xq.device)
    mu = torch.empty(n, dtype=torch.float64, device=Xq.device)
    scratch = torch.empty(state.d * state.Np, dtype=torch.float64, device=Xq.device)
    gp_c = state.c_struct()
    st = lib.cbo_kernel_rows_f64(C.byref(gp_c), Xq.data_ptr(), n, U.data_ptr(), mu.data_ptr(), scratch.data_ptr(), _stream())
    _lib.check(st, "cbo_kernel_rows_f64")
    return U, mu
