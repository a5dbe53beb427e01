"""ctypes binding of libcbo.so (include/cbo.h).  There is no CPU fallback: if the library is
missing or a call fails, this raises."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None

EXPORTS = [
    "cbo_version", "cbo_status_str", "cbo_last_cuda_error", "cbo_padded_n", "cbo_workspace_bytes",
    "cbo_conformal_acq_f64", "cbo_conformal_acq_host_f64", "cbo_conf_mask_to_bounds_f64",
    "cbo_ratio_forward_f64", "cbo_gemm_tn_f64", "cbo_kernel_rows_f64", "cbo_timing_enable", "cbo_timing_collect",
    "cbo_cehvi_integrate_f64", "cbo_cehvi_integrate_q_f64", "cbo_cehvi_integrate_q_grad_f64", "cbo_conformal_acq_grad_f64", "cbo_workspace_bytes_grad",
    "cbo_model_workspace_bytes", "cbo_model_build_f64",
    "cbo_model_append_workspace_bytes", "cbo_model_append_f64",
    "cbo_ratio_train_workspace_bytes", "cbo_ratio_train_step_f64", "cbo_sample_subset_i64", "cbo_sgld_update_f64", "cbo_launch_count",
]

c_dp = C.c_void_p  # device pointers travel as integers


class GpF64(C.Structure):
    _fields_ = [("N", C.c_int64), ("d", C.c_int64), ("Np", C.c_int64), ("X", c_dp), ("alpha", c_dp),
                ("linv", c_dp), ("linv_t", c_dp), ("inv_lengthscale", c_dp), ("mean", C.c_double),
                ("outputscale", C.c_double), ("noise", C.c_double)]


class RatioMlpF64(C.Structure):
    _fields_ = [("hidden", C.c_int64), ("W1", c_dp), ("b1", c_dp), ("W2", c_dp), ("b2", c_dp), ("W3", c_dp),
                ("b3", c_dp)]


class NeiBaselineF64(C.Structure):
    _fields_ = [("nb", C.c_int64), ("Xb", c_dp), ("KbX", c_dp), ("Sbb", c_dp), ("mub", c_dp)]


class ExtAdjointsF64(C.Structure):
    _fields_ = [("mask_bar", c_dp), ("ood_bar", c_dp), ("f_bar", c_dp)]


class Params(C.Structure):
    _fields_ = [("alpha", C.c_double), ("temp", C.c_double), ("grid_res", C.c_int64), ("q", C.c_int64),
                ("num_mc", C.c_int64), ("num_probes", C.c_int64), ("max_grid_refinements", C.c_int64),
                ("kind", C.c_int32), ("optimistic", C.c_int32), ("randomized", C.c_int32),
                ("eps_per_candidate", C.c_int32), ("beta_prime", C.c_double), ("best_f", C.c_double),
                ("nei", C.POINTER(NeiBaselineF64)), ("refine_reduce", C.c_void_p), ("refine_ctx", C.c_void_p),
                ("grid_cov_scale", C.c_void_p), ("ext_adjoints", C.POINTER(ExtAdjointsF64)), ("grid_grad_leak", C.c_int32)]


REFINE_REDUCE_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_void_p)


class OutputsF64(C.Structure):
    _fields_ = [(n, c_dp) for n in ("acq", "target_grid", "grid_logp", "conf_pred_mask", "ood_mask",
                                     "q_conf_scores", "post_mean", "post_var", "num_accepted", "refit_mean",
                                     "refit_sd", "mc_values")]


class CboError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB


def load():
    """Load libcbo.so, building it first if nvcc is available and the sources are newer."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("CBO_LIB_PATH") or lib_path()     # CBO_LIB_PATH: experiment builds (tools/); never a fallback
    if path == lib_path() and _build.is_stale():
        try:
            _build.build()
        except _build.NvccNotFound as e:    # no compiler on this host: the shipped .so is used -- if there is one
            if not os.path.exists(path):
                raise CboError(f"libcbo.so is missing and could not be built: {e}") from e
        # any other failure (a compile error in an edited source) propagates: never run stale kernels silently
    lib = C.CDLL(path)
    lib.cbo_version.restype = C.c_int
    lib.cbo_launch_count.restype = C.c_int64
    lib.cbo_status_str.restype = C.c_char_p
    lib.cbo_status_str.argtypes = [C.c_int]
    lib.cbo_last_cuda_error.restype = C.c_char_p
    lib.cbo_padded_n.restype = C.c_int64
    lib.cbo_padded_n.argtypes = [C.c_int64]
    lib.cbo_workspace_bytes.restype = C.c_size_t
    lib.cbo_workspace_bytes.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.POINTER(Params), C.c_int64]
    lib.cbo_conformal_acq_f64.restype = C.c_int
    lib.cbo_conformal_acq_f64.argtypes = [C.POINTER(GpF64), C.POINTER(RatioMlpF64), C.POINTER(Params), c_dp, C.c_int64,
                                          c_dp, c_dp, c_dp, c_dp, C.POINTER(OutputsF64), c_dp, C.c_size_t, C.c_int64,
                                          c_dp]
    lib.cbo_workspace_bytes_grad.restype = C.c_size_t
    lib.cbo_workspace_bytes_grad.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.POINTER(Params), C.c_int64]
    lib.cbo_conformal_acq_grad_f64.restype = C.c_int
    lib.cbo_conformal_acq_grad_f64.argtypes = [C.POINTER(GpF64), C.POINTER(RatioMlpF64), C.POINTER(Params), c_dp, C.c_int64,
                                               c_dp, c_dp, c_dp, c_dp, C.POINTER(OutputsF64), c_dp, c_dp, C.c_size_t,
                                               C.c_int64, c_dp]
    lib.cbo_conformal_acq_host_f64.restype = C.c_int
    lib.cbo_conformal_acq_host_f64.argtypes = [C.POINTER(GpF64), C.POINTER(RatioMlpF64), C.POINTER(Params), c_dp,
                                               C.c_int64, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, C.c_size_t,
                                               C.c_int64, c_dp]
    lib.cbo_conf_mask_to_bounds_f64.restype = C.c_int
    lib.cbo_conf_mask_to_bounds_f64.argtypes = [c_dp, c_dp, C.c_int64, C.c_int64, C.c_int64, c_dp, c_dp, c_dp]
    lib.cbo_ratio_forward_f64.restype = C.c_int
    lib.cbo_ratio_forward_f64.argtypes = [C.POINTER(RatioMlpF64), c_dp, C.c_int64, C.c_int64, c_dp, c_dp]
    lib.cbo_gemm_tn_f64.restype = C.c_int
    lib.cbo_gemm_tn_f64.argtypes = [c_dp, C.c_int64, c_dp, C.c_int64, c_dp, C.c_int64, C.c_int64, C.c_int64,
                                    C.c_int64, C.c_int32, c_dp, c_dp]
    lib.cbo_kernel_rows_f64.restype = C.c_int
    lib.cbo_kernel_rows_f64.argtypes = [C.POINTER(GpF64), c_dp, C.c_int64, c_dp, c_dp, c_dp, c_dp]
    PP = C.POINTER(C.c_void_p)
    lib.cbo_cehvi_integrate_f64.restype = C.c_int
    lib.cbo_cehvi_integrate_f64.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int64, PP, PP, PP, PP, PP, PP, PP, c_dp,
                                            c_dp, c_dp, C.c_int64, C.c_double, C.c_int32, c_dp, c_dp]
    lib.cbo_cehvi_integrate_q_f64.restype = C.c_int
    lib.cbo_cehvi_integrate_q_f64.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, PP, PP, PP, PP, PP, PP,
                                              C.POINTER(C.c_double), PP, c_dp, c_dp, c_dp, C.c_int64, C.c_int32, C.c_double,
                                              C.c_int32, c_dp, c_dp, c_dp]
    lib.cbo_cehvi_integrate_q_grad_f64.restype = C.c_int
    lib.cbo_cehvi_integrate_q_grad_f64.argtypes = lib.cbo_cehvi_integrate_q_f64.argtypes[:-1] + [PP, PP, PP, c_dp]
    lib.cbo_model_workspace_bytes.restype = C.c_size_t
    lib.cbo_model_workspace_bytes.argtypes = [C.c_int64, C.c_int64]
    lib.cbo_model_build_f64.restype = C.c_int
    lib.cbo_model_build_f64.argtypes = [c_dp, c_dp, C.c_int64, C.c_int64, c_dp, C.c_double, C.c_double, C.c_double, c_dp,
                                        c_dp, c_dp, c_dp, c_dp, C.c_size_t, c_dp]
    lib.cbo_model_append_workspace_bytes.restype = C.c_size_t
    lib.cbo_model_append_workspace_bytes.argtypes = [C.c_int64]
    lib.cbo_model_append_f64.restype = C.c_int
    lib.cbo_model_append_f64.argtypes = [c_dp, C.c_int64, C.c_int64, C.c_int64, c_dp, C.c_double, C.c_double, C.c_double,
                                         c_dp, c_dp, c_dp, c_dp, c_dp, C.c_int64, c_dp, c_dp, C.c_size_t, c_dp]
    lib.cbo_ratio_train_workspace_bytes.restype = C.c_size_t
    lib.cbo_ratio_train_workspace_bytes.argtypes = [C.c_int64, C.c_int64]
    lib.cbo_ratio_train_step_f64.restype = C.c_int
    lib.cbo_ratio_train_step_f64.argtypes = [C.c_int64, C.c_int64, c_dp, c_dp, c_dp, c_dp, c_dp, C.c_int64, c_dp, c_dp,
                                             C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double,
                                             C.c_double, C.c_double, c_dp, c_dp, C.c_size_t, c_dp]
    lib.cbo_sample_subset_i64.restype = C.c_int
    lib.cbo_sample_subset_i64.argtypes = [C.c_int64, C.c_int64, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.cbo_sgld_update_f64.restype = C.c_int
    lib.cbo_sgld_update_f64.argtypes = [c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, C.c_int64, C.c_int64, C.c_double, C.c_double,
                                        C.c_double, C.c_double, C.c_double, c_dp]
    lib.cbo_timing_enable.restype = None
    lib.cbo_timing_enable.argtypes = [C.c_int]
    lib.cbo_timing_collect.restype = C.c_int
    lib.cbo_timing_collect.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    if lib.cbo_version() != 104:
        raise CboError(f"libcbo.so version mismatch: {lib.cbo_version()}")
    _LIB = lib
    return lib


def check(status: int, what: str):
    if status != 0:
        lib = load()
        msg = lib.cbo_status_str(status).decode()
        if status == 4:
            msg += " -- " + lib.cbo_last_cuda_error().decode()
        raise CboError(f"{what}: {msg} (status {status})")
