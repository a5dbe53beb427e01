"""CUDA-graph replay of a fixed-shape value (+ gradient) call.

A call with a handful of candidate points (the SGLD loop of conformalbo/optim/optimize.py:224-235: `num_restarts x q x d`
with grad, hundreds of steps per round) enqueues ~20 small kernels and is bound by their launches on the host.  The library
only enqueues work on the caller's stream (no allocation, no synchronisation when `max_grid_refinements == 0`), so the whole
call can be captured once with torch's stream capture and replayed: per step one `cudaGraphLaunch` plus the copies of the
new candidates / probes / uniforms into the graph's static input buffers.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .helpers import draw_probes
from .model_state import GPModelState


class GraphedConformalAcq:
    """`ops.conformal_acq(state, X (b, q, d), ...)` for ONE fixed (b, q) captured into a CUDA graph.

    __call__(X, z=None, u_rand=None) -> dict(acq (b,), dacq_dX (b, q, d) if want_grad); the returned tensors are the
    graph's static outputs (overwritten by the next replay -- clone them to keep).  Fresh Rademacher probes are drawn per
    call like the reference does (helpers.py:399-400) unless `z` is given; `u_rand` likewise for `randomized=True`.
    """

    def __init__(self, state: GPModelState, b: int, q: int, *, eps_grid: torch.Tensor, eps_mc: torch.Tensor,
                 ratio: Optional["ops.RatioWeights"] = None, want_grad: bool = True, num_probes: int = 8,
                 device_probes: bool = False, **param_kw):
        if int(param_kw.get("max_grid_refinements", 0)) > 0:
            raise ValueError("grid refinement synchronises with the host once per pass and cannot be captured")
        dev = state.device
        self.state, self.b, self.q, self.want_grad = state, int(b), int(q), bool(want_grad)
        self.randomized = bool(param_kw.get("randomized", False))
        self.device_probes = bool(device_probes)          # probes from the device generator: no host work per replay
        G = int(param_kw["grid_res"])
        f64 = dict(dtype=torch.float64, device=dev)
        self.X = torch.zeros(b, q, state.d, **f64)
        self.z = torch.ones(state.N + q, num_probes, **f64)
        self.u = torch.full((b, G, q), 0.5, **f64) if self.randomized else None
        self._kw = dict(eps_grid=eps_grid.to(**f64).contiguous(), eps_mc=eps_mc.to(**f64).contiguous(), ratio=ratio,
                        want_grad=self.want_grad, **param_kw)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm-up outside the capture (lazy function attributes, allocations)
            for _ in range(2):
                ops.conformal_acq(state, self.X, z=self.z, u_rand=self.u, **self._kw)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # capture_begin / capture_end by hand on the side stream: the `torch.cuda.graph` context manager first runs
        # gc.collect() and torch.cuda.empty_cache() -- 20-570 ms per capture in a BO process (it also drops the cached
        # workspace of the trajectory-scoring call), against ~3 ms for the capture itself and 22 ms for 128 replayed steps
        self.graph = torch.cuda.CUDAGraph()
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self.graph.capture_begin()
            try:
                self.out = ops.conformal_acq(state, self.X, z=self.z, u_rand=self.u, **self._kw)
            finally:
                self.graph.capture_end()
        torch.cuda.current_stream(dev).wait_stream(side)

    def __call__(self, X: torch.Tensor, z: Optional[torch.Tensor] = None, u_rand: Optional[torch.Tensor] = None):
        self.X.copy_(X.reshape(self.X.shape), non_blocking=True)
        self.z.copy_(z if z is not None else draw_probes(self.z.shape[0], self.z.shape[1], device=self.z.device,
                                                         on_device=self.device_probes), non_blocking=True)
        if self.randomized:
            if u_rand is not None:
                self.u.copy_(u_rand.reshape(self.u.shape), non_blocking=True)
            else:
                self.u.uniform_()
        self.graph.replay()
        return self.out
