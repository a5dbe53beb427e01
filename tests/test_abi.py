"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads and exports every symbol of include/cbo.h."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "cbo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cbo_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from conformal_bayesopt_b200 import _lib, build
    path = build.build()
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert set(declared) == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    lib.cbo_status_str.restype = ctypes.c_char_p
    assert lib.cbo_version() == 104
    assert lib.cbo_status_str(0) == b"ok" and lib.cbo_status_str(2) == b"unsupported shape"
    lib.cbo_padded_n.restype = ctypes.c_int64
    lib.cbo_padded_n.argtypes = [ctypes.c_int64]
    assert [lib.cbo_padded_n(n) for n in (1, 128, 129, 1024)] == [128, 128, 256, 1024]


def test_workspace_bytes_and_argument_validation():
    from conformal_bayesopt_b200 import _lib, ops
    lib = _lib.load()
    p = ops.make_params(alpha=0.1, temp=0.03, grid_res=64)
    small = lib.cbo_workspace_bytes(1024, 8, 4096, ctypes.byref(p), 0)
    big = lib.cbo_workspace_bytes(1024, 8, 1 << 20, ctypes.byref(p), 0)
    assert small % 256 == 0 and 2 * 4096 * 1024 * 8 < small < 2.2 * 4096 * 1024 * 8
    assert big < 2.2 * 65536 * 1024 * 8          # chunked: independent of b beyond the chunk size
    pr = ops.make_params(alpha=0.1, temp=0.03, grid_res=64, max_grid_refinements=2)
    assert lib.cbo_workspace_bytes(1024, 8, 1 << 16, ctypes.byref(pr), 0) > 2 * (1 << 16) * 1024 * 8   # refinement keeps all candidates
    # null pointers / bad sizes are rejected with status codes before any CUDA call
    assert lib.cbo_conformal_acq_f64(None, None, ctypes.byref(p), None, 4, None, None, None, None, None, None, 0, 0, None) == 1
    assert lib.cbo_gemm_tn_f64(None, 0, None, 0, None, 0, 1, 128, 128, 0, None, None) == 1
    assert lib.cbo_conf_mask_to_bounds_f64(None, None, 1, 8, 1, None, None, None) == 1
    # gradient call: extra solve scratch + adjoints; model builder: O(Np^2) scratch
    gsmall = lib.cbo_workspace_bytes_grad(1024, 8, 4096, ctypes.byref(p), 0)
    assert gsmall % 256 == 0 and small + 4096 * 1024 * 8 <= gsmall < small + 1.1 * 4096 * 1024 * 8
    assert lib.cbo_conformal_acq_grad_f64(None, None, ctypes.byref(p), None, 4, None, None, None, None, None, None, None, 0, 0, None) == 1
    mb = lib.cbo_model_workspace_bytes(1000, 8)
    assert mb % 256 == 0 and 1024 * 1024 * 8 < mb < 2 * 1024 * 1024 * 8
    assert lib.cbo_model_build_f64(None, None, 10, 2, None, 0.0, 1.0, 0.1, None, None, None, None, None, 0, None) == 1


def test_cpu_model_state_is_refused():
    import torch
    from conformal_bayesopt_b200 import GPModelState, _lib
    with pytest.raises(_lib.CboError):
        GPModelState.fit_fixed(torch.rand(8, 2, dtype=torch.float64), torch.zeros(8, dtype=torch.float64), 0.0, 0.5, 1.0, 1e-2)


def test_ops_refuse_cpu_tensors():
    import torch
    from conformal_bayesopt_b200 import _lib, ops
    with pytest.raises(_lib.CboError):
        ops.conf_mask_to_bounds(torch.zeros(2, 4, 1, 1, dtype=torch.float64), torch.zeros(2, 4, 1, 1, dtype=torch.float64))


def test_torch_extension_builds_loads_and_registers_its_operators():
    """lib/libcbo_torch.so (csrc/torch_ext.cpp): TORCH_LIBRARY(cbo) schemas are visible without a GPU; a CPU tensor finds no
    kernel (there is no CPU path to fall back to)."""
    import pytest
    import torch
    from conformal_bayesopt_b200 import torch_ops
    t = torch_ops.load()
    for name in ("conformal_acq", "conformal_acq_raw", "model_build", "conf_mask_to_bounds", "ratio_forward", "ratio_train_step",
                 "sgld_update"):
        assert hasattr(t, name), name
    schema = str(torch._C._jit_get_schemas_for_operator("cbo::conformal_acq")[0])
    assert "Tensor Xq" in schema and "Tensor? u_rand" in schema and schema.endswith("-> Tensor")
    with pytest.raises((NotImplementedError, RuntimeError)):
        t.conf_mask_to_bounds(torch.zeros(2, 4, 1, 1, dtype=torch.float64), torch.zeros(2, 4, 1, 1, dtype=torch.float64))
