#!/usr/bin/env python
"""bench.py -- conformal acquisition evaluations/sec (candidates x grid points) on B200.

Contract (see DESIGN.md "Measurement"):
  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  N > 1:  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
One step = one pass of the full-conformal acquisition hot path (cUCB forward) over one batch of
synthetic candidates at N_train=1024, d=8, grid=64 (BASELINE.json configs[4]: the 2^20-candidate
sweep; every rank scores its own 2^20 candidates -> weak scaling, then the ranks exchange
(max, argmax) pairs with one tiny NCCL all-gather).  `value` is device-resident throughput,
`e2e` goes through the C ABI's host-buffer entry point (H2D of the candidates and D2H of the
acquisition values inside the timed region).  Rank 0 prints ONE JSON line.

--impl reference times the CPU restatement of the reference's own dense algorithm (oracle/dense.py,
"port": botorch/gpytorch are absent from this image) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRAIN, DIM, GRID = 1024, 8, 64
FLOP_PER_EVAL = 51811.0            # algorithmic fp64 flop per (candidate x grid point)
GEMM_FLOP_PER_CAND_PASS = float(N_TRAIN) * N_TRAIN      # one triangular solve pass: N^2 flop (N^2/2 FMA)
SCORE_FLOP_PER_EVAL = 17.0 * (N_TRAIN + 1)              # SURVEY.md §8d per-(cand,grid) term
FP64_PEAK_TFLOPS = 37.0            # measured on this pool's B200 (profiles/r01_fp64_peak.log): DMMA.8x8x4 = whole fp64 pipe
NCU_DRAM_BYTES_PER_LAUNCH = {"score_rank_integrate": 551.5e6, "gemm_pass1_dmma": 1395.0e6, "gemm_pass2_dmma": 1391.1e6,
                             "kernel_rows": 487.0e6}         # profiles/r02_ncu_summary.txt, first block (dram__bytes_read + write per launch of 65536 candidates)
BYTES_PER_EVAL = (8 * DIM + 8) / GRID                   # SURVEY.md §8d algorithmic HBM bytes per eval (no mask output)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cands", type=int, default=1 << 20, help="candidates per rank per step")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--temp", type=float, default=3e-2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-grad-leg", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the ratio-estimator and strong-scaling legs")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock / throttle reasons during the timed region (NVML; nvidia-smi fallback)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def workload_cpu(seed_rank: int, b: int):
    import torch
    from conformal_bayesopt_b200.synthetic import make_inputs, make_problem
    prob = make_problem(N_TRAIN, DIM, seed=0)                      # same model on every rank
    inp = make_inputs(N_TRAIN, DIM, 1, 1, GRID, seed=0)            # shared random inputs (eps_grid, z, eps_mc)
    g = torch.Generator().manual_seed(2 + seed_rank)
    Xq = torch.rand(b, DIM, generator=g, dtype=torch.float64)      # torch.manual_seed(2) candidates
    return prob, inp, Xq


PARAMS = dict(kind="ucb", alpha=1.0 / math.sqrt(N_TRAIN), grid_res=GRID, max_grid_refinements=0, randomized=False,
              optimistic=True, beta_prime=math.sqrt(0.2 * math.pi / 2))


# ------------------------------------------------------------------------------------------------
def cpu_baseline(temp: float):
    """Rank 0, N=1 only: the oracle timed on the host cores, on a bounded sample of the same workload."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    from oracle import closed_form, dense
    from oracle import gp as ogp
    prob, inp, Xq = workload_cpu(0, 2048)
    gp = ogp.build_state(prob["X"], prob["y"], prob["mean"], prob["lengthscale"], prob["outputscale"], prob["noise"])
    kw = dict(eps_grid=inp["eps_grid"], z=inp["z"], eps_mc=inp["eps_mc"], temp=temp, **PARAMS)
    cores = torch.get_num_threads()
    # (i) dense literal restatement of the reference algorithm: b = 2 candidates (128 evals), once after a tiny warm-up
    dense.conformal_acquisition(gp, Xq[:1, None, :], **{**kw, "grid_res": 2, "eps_grid": inp["eps_grid"][:2], "eps_mc": inp["eps_mc"][:, :2]})
    t0 = time.perf_counter()
    dense.conformal_acquisition(gp, Xq[:2, None, :], **kw)
    t_dense = time.perf_counter() - t0
    # (ii) closed-form restatement (algorithm-matched): 2048 candidates in blocks of 128
    closed_form.conformal_acquisition(gp, Xq[:128, None, :], **kw)
    t0 = time.perf_counter()
    for i in range(0, 2048, 128):
        closed_form.conformal_acquisition(gp, Xq[i:i + 128, None, :], **kw)
    t_cf = time.perf_counter() - t0
    return {
        "value": 2 * GRID / t_dense, "unit": "evals/s", "cores": cores, "kind": "port",
        "sample": f"oracle/dense.py (reference algorithm, dense (N+1)^2 refit per (cand,grid)), 2 candidates x {GRID} grid, {t_dense:.1f} s",
        "closed_form": {"value": 2048 * GRID / t_cf, "unit": "evals/s", "cores": cores,
                        "sample": f"oracle/closed_form.py (same algebra as the kernels), 2048 candidates x {GRID} grid, {t_cf:.1f} s"},
    }


def run_reference(args):
    """Reference arm: CPU restatement of the reference's dense algorithm, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)    # torchrun exports OMP_NUM_THREADS=1: the CPU arm must not be strangled by it
    from oracle import dense
    from oracle import gp as ogp
    b = 2 if (args.steps + args.warmup) <= 16 else 1
    prob, inp, Xq = workload_cpu(0, b)
    gp = ogp.build_state(prob["X"], prob["y"], prob["mean"], prob["lengthscale"], prob["outputscale"], prob["noise"])
    kw = dict(eps_grid=inp["eps_grid"], z=inp["z"], eps_mc=inp["eps_mc"], temp=args.temp, **PARAMS)
    for _ in range(args.warmup):
        dense.conformal_acquisition(gp, Xq[:, None, :], **kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dense.conformal_acquisition(gp, Xq[:, None, :], **kw)
    dt = time.perf_counter() - t0
    val = args.steps * b * GRID / dt
    cores = torch.get_num_threads()
    sample = f"oracle/dense.py: {b} candidates x {GRID} grid points per step (dense (N+1)^2 refit per pair), N_train={N_TRAIN}"
    print(json.dumps({
        "impl": "reference", "metric": "conformal acq evals/sec (cands x grid)", "value": val, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic 2^20-candidate acquisition sweep at N=1024, d=8, grid=64 (BASELINE configs[4])",
                   "n_train": N_TRAIN, "d": DIM, "grid": GRID, "acq": "cucb", "temp": args.temp, "alpha": PARAMS["alpha"],
                   "randomized": False, "ratio_estimator": None,
                   "sample_per_step": f"{b} candidates x {GRID} grid points of that workload (bounded CPU sample)"},
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from conformal_bayesopt_b200 import GPModelState, _lib, ops
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    b = args.cands
    prob, inp, Xq_cpu = workload_cpu(rank, b)
    state = GPModelState.fit_fixed(prob["X"], prob["y"], prob["mean"], prob["lengthscale"], prob["outputscale"],
                                   prob["noise"], device=dev)
    eps_grid = inp["eps_grid"][:, 0].to(dev)
    z = inp["z"].to(dev)
    eps_mc = inp["eps_mc"][:, :, 0].to(dev)
    Xq = Xq_cpu.to(dev)
    kw = dict(eps_grid=eps_grid, z=z, eps_mc=eps_mc, temp=args.temp, chunk=args.chunk, **PARAMS)
    gathered = [torch.empty(2, dtype=torch.float64, device=dev) for _ in range(world)] if world > 1 else None

    def step():
        acq = ops.conformal_acq(state, Xq, **kw)["acq"]
        best = torch.max(acq, dim=0)
        if world > 1:   # the path's only exchange: (value, global index) pairs, 16 bytes per rank
            pair = torch.stack([best.values, (best.indices + rank * b).double()])
            dist.all_gather(gathered, pair)
        return acq

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    lib.cbo_timing_enable(1)
    ms = (C.c_double * 8)()
    nl = (C.c_int64 * 8)()
    lib.cbo_timing_collect(ms, nl)
    sampler = ClockSampler(local)
    launches0 = lib.cbo_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    elapsed_ms = e0.elapsed_time(e1)
    launches = lib.cbo_launch_count() - launches0      # counted by the library at its launch sites
    lib.cbo_timing_collect(ms, nl)
    lib.cbo_timing_enable(0)
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = world * b * GRID * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end: host (pinned) candidates -> C ABI host entry point -> host acquisition values -------------
    e2e = None
    if not args.no_e2e:
        Xh = Xq_cpu.pin_memory()
        acq_h = torch.empty(b, dtype=torch.float64).pin_memory()
        staging = (torch.empty(b, DIM, dtype=torch.float64, device=dev), torch.empty(b, dtype=torch.float64, device=dev))
        kwh = {k: v for k, v in kw.items()}
        for _ in range(2):
            ops.conformal_acq_host(state, Xh, acq_h, staging=staging, **kwh)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ops.conformal_acq_host(state, Xh, acq_h, staging=staging, **kwh)
            if world > 1:
                m = acq_h.max(dim=0)
                pair = torch.tensor([float(m.values), float(m.indices + rank * b)], dtype=torch.float64, device=dev)
                dist.all_gather(gathered, pair)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3       # host wall clock: the call synchronises its stream on return
        if world > 1:
            t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        e2e = {"value": world * b * GRID * args.steps / (e2e_ms * 1e-3), "unit": "evals/s",
               "h2d_bytes_per_step": b * DIM * 8, "d2h_bytes_per_step": b * 8,
               "api": "cbo_conformal_acq_host_f64 (pinned host candidates in, host acquisition values out; H2D / D2H pipelined "
                      "per 65536-candidate chunk on a copy stream)"}

    # ---- BASELINE configs[1] (4096 raw candidates per call), device-resident, rank 0 only, outside the timed region ----
    cfg2 = None
    if rank == 0:
        X2 = Xq[:4096].contiguous()
        for _ in range(3):
            ops.conformal_acq(state, X2, **kw)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(20):
            ops.conformal_acq(state, X2, **kw)
        g1.record()
        torch.cuda.synchronize()
        ms2 = g0.elapsed_time(g1) / 20
        cfg2 = {"workload": "ackley 8-D cucb, N=1024, 64-point grid, 4096 raw candidates per call (BASELINE configs[1])",
                "ms_per_call": ms2, "value": 4096 * GRID / (ms2 * 1e-3), "unit": "evals/s"}

    # ---- value + gradient sweep (cbo_conformal_acq_grad_f64: what optimize_acqf's L-BFGS / SGLD steps call), rank 0 only,
    # outside the timed region; per-kernel-class CUDA-event times from the library
    cfg_grad = None
    if rank == 0 and not args.no_grad_leg:
        bg_ = min(b, 1 << 18)
        Xg = Xq[:bg_].contiguous()
        for _ in range(2):
            ops.conformal_acq(state, Xg, want_grad=True, **kw)
        torch.cuda.synchronize()
        lib.cbo_timing_enable(1)
        msg_, nlg_ = (C.c_double * 8)(), (C.c_int64 * 8)()
        lib.cbo_timing_collect(msg_, nlg_)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(3):
            ops.conformal_acq(state, Xg, want_grad=True, **kw)
        g1.record()
        torch.cuda.synchronize()
        lib.cbo_timing_collect(msg_, nlg_)
        lib.cbo_timing_enable(0)
        msg3 = g0.elapsed_time(g1) / 3
        gnames = ["prep_tables", "kernel_rows", "gemm_pass1_dmma", "gemm_pass2_dmma", "score_rank_integrate",
                  "score_value_and_reverse", "gemm_reverse_x2_dmma", "kernel_rows_reverse"]
        cfg_grad = {"workload": f"value + d acq/dX of {bg_} candidates per call (N=1024, d=8, grid=64, cucb), device-resident",
                    "ms_per_call": msg3, "value": bg_ * GRID / (msg3 * 1e-3), "unit": "evals/s (each with its gradient row)",
                    "kernels_ms_per_call": {n: msg_[i] / 3 for i, n in enumerate(gnames) if msg_[i] > 0}}

    # ---- one SGLD step of the reference's default inner loop (optim/optimize.py:224-235: 8 restarts x q = 2, value+gradient,
    # randomised, ratio estimator), replayed as a CUDA graph with fresh probes / uniforms per step; rank 0, outside the timed region
    cfg_sgld = None
    if rank == 0 and not args.no_grad_leg:
        from conformal_bayesopt_b200.graphs import GraphedConformalAcq
        from conformal_bayesopt_b200.synthetic import make_inputs, make_ratio_weights
        import time as _time
        inp_s = make_inputs(N_TRAIN, DIM, 8, 2, GRID, seed=3)
        pk = {k: v for k, v in kw.items() if k not in ("eps_grid", "z", "eps_mc", "u_rand", "outputs", "chunk")}
        pk.update(randomized=True, ratio=ops.RatioWeights(*[t.to(dev) for t in make_ratio_weights(DIM)]))
        gacq = GraphedConformalAcq(state, 8, 2, eps_grid=inp_s["eps_grid"].to(dev), eps_mc=inp_s["eps_mc"].to(dev), **pk)
        Xs = inp_s["Xq"].to(dev)
        for _ in range(5):
            gacq(Xs)
        torch.cuda.synchronize()
        t0 = _time.perf_counter()
        for _ in range(50):
            gacq(Xs)
        torch.cuda.synchronize()
        ms_s = (_time.perf_counter() - t0) / 50 * 1e3
        cfg_sgld = {"workload": "one SGLD step of the reference's default loop: value + d acq/dX of 8 restarts x q=2 (N=1024, d=8, "
                                "grid=64, cucb, randomised, ratio estimator), CUDA-graph replay, fresh probes and uniforms per step",
                    "ms_per_call": ms_s, "value": 8 * GRID / (ms_s * 1e-3), "unit": "evals/s (joint batches x grid points)"}

        # the whole inner loop of optimize_acqf_sgld as the reference runs it (optim/optimize.py:196-247: callback = one training
        # step of the ratio estimator per SGLD step after warm-up, value + gradient, SGHMC update, trajectory scored at the end)
        from conformal_bayesopt_b200.acquisitions import qConformalUpperConfidenceBound
        from conformal_bayesopt_b200.optim import optimize_acqf_sgld
        from conformal_bayesopt_b200.ratio_estimation import RatioEstimator
        torch.manual_seed(0)
        est = RatioEstimator(in_size=DIM, device=dev, dtype=torch.float64, lr=5e-3, ema_weight=2e-2, weight_decay=5e-6)
        est.set_neg_samples(state.X[:N_TRAIN].contiguous())
        cacq = qConformalUpperConfidenceBound(alpha=N_TRAIN ** -0.35, temp=args.temp, grid_res=GRID, max_grid_refinements=0,
                                              ratio_estimator=est, randomized=True, model=state, beta=0.2)
        cacq.device_probes = True
        bnds = torch.stack([torch.zeros(DIM), torch.ones(DIM)]).to(dev, torch.float64)
        ics = torch.rand(8, 2, DIM, dtype=torch.float64, device=dev)

        def sgld_loop(n):
            torch.cuda.synchronize()
            t0_ = _time.perf_counter()
            optimize_acqf_sgld(cacq, bnds, q=2, num_restarts=8, batch_initial_conditions=ics, sgld_steps=n, warmup_steps=32, lr=2e-2,
                               temperature=4e-3, return_full_traj=True, options=dict(callback=est.optimize_callback))
            torch.cuda.synchronize()
            return _time.perf_counter() - t0_
        sgld_loop(40)
        t128 = min(sgld_loop(128) for _ in range(2))
        t512 = min(sgld_loop(512) for _ in range(2))
        cfg_sgld["loop"] = {"workload": "optimize_acqf_sgld, 8 restarts x q=2, ratio estimator trained by the callback, graph-replayed "
                                        "steps, probes and minibatch drawn on the device, trajectory scored at the end",
                            "ms_per_step_128_steps_all_in": t128 / 128 * 1e3, "ms_per_step_marginal": (t512 - t128) / 384 * 1e3}

    # ---- density-ratio-weighted, randomised leg (seeded MLP 8->64->64->1, randomized=True with injected uniforms) at
    # the headline size, device-resident; rank 0, outside the timed region; with its own per-kernel times and roofline fraction
    cfg_ratio = None
    if rank == 0 and not args.no_extra_legs:
        from conformal_bayesopt_b200.synthetic import make_ratio_weights
        ratio_w = ops.RatioWeights(*[t.to(dev) for t in make_ratio_weights(DIM)])
        u_rand = torch.rand(b, GRID, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
        kwr = dict(kw, randomized=True, u_rand=u_rand, ratio=ratio_w)
        for _ in range(2):
            ops.conformal_acq(state, Xq, **kwr)
        torch.cuda.synchronize()
        lib.cbo_timing_enable(1)
        msr_, nlr_ = (C.c_double * 8)(), (C.c_int64 * 8)()
        lib.cbo_timing_collect(msr_, nlr_)
        n_r = 3
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(n_r):
            ops.conformal_acq(state, Xq, **kwr)
        g1.record()
        torch.cuda.synchronize()
        lib.cbo_timing_collect(msr_, nlr_)
        lib.cbo_timing_enable(0)
        ms_r = g0.elapsed_time(g1) / n_r
        score_tf = b * GRID * SCORE_FLOP_PER_EVAL / (msr_[4] / n_r * 1e-3) / 1e12
        cfg_ratio = {"workload": f"{b}-candidate sweep at N=1024, d=8, grid=64, cucb WITH the density-ratio MLP (8->64->64->1, seeded) on "
                                 "train points and candidates and randomized=True (injected uniforms, 8 bytes per eval), device-resident",
                     "ms_per_step": ms_r, "value": b * GRID / (ms_r * 1e-3), "unit": "evals/s",
                     "kernels_ms_per_step": {n: msr_[i] / n_r for i, n in enumerate(
                         ["prep_tables", "kernel_rows+ratio_mlp", "gemm_pass1_dmma", "gemm_pass2_dmma", "score_rank_integrate"])},
                     "roofline": {"bound": "tensor", "kernel": "score_rank_integrate", "achieved": score_tf, "peak": FP64_PEAK_TFLOPS,
                                  "unit": "TFLOP/s", "frac": score_tf / FP64_PEAK_TFLOPS,
                                  "whole_path_frac": b * GRID / (ms_r * 1e-3) * FLOP_PER_EVAL / 1e12 / FP64_PEAK_TFLOPS}}
        del u_rand, kwr

    # ---- strong scaling (north star: ONE optimize_acqf candidate batch sharded over the GPUs of the box, then a tiny exchange):
    # a fixed TOTAL of 2^20 and of 4096 candidates; per step = local scoring + exchange + the host read of the winner
    strong, argmax_check = [], None
    if world > 1 and not args.no_extra_legs:
        from conformal_bayesopt_b200 import distributed as cdist
        _, _, Xall_cpu = workload_cpu(0, 1 << 20)                  # the same global candidate set on every rank
        acq_fn = lambda X: ops.conformal_acq(state, X, **kw)["acq"]
        for total in (1 << 20, 4096):
            lo, hi = cdist.shard_bounds(total, rank, world)
            Xs = Xall_cpu[lo:hi].to(dev)
            res = {}
            for mode in ("compute_only", "argmax_pairs", "score_all_gather"):
                if mode == "compute_only":
                    fn = lambda: acq_fn(Xs)
                elif mode == "argmax_pairs":
                    fn = lambda: cdist.sharded_argmax(acq_fn, Xs, lo)
                else:
                    Xfull = Xall_cpu[:total].to(dev)
                    fn = lambda: int(cdist.sharded_acquisition(acq_fn, Xfull).argmax())
                n_it = 3 if total > 100000 else 30
                for _ in range(3):
                    fn()
                barrier()
                t0 = time.perf_counter()
                for _ in range(n_it):
                    fn()
                barrier()
                t = torch.tensor([(time.perf_counter() - t0) * 1e3 / n_it], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                res[mode] = float(t.item())
            strong.append({"total_candidates": total, "per_rank": hi - lo, "ms_per_step": res,
                           "evals_per_s": {k: total * GRID / (v * 1e-3) for k, v in res.items()},
                           "exchange_and_host_sync_share": {k: 1.0 - res["compute_only"] / res[k] for k in ("argmax_pairs", "score_all_gather")},
                           "timing": "host wall clock around synchronised steps (every step ends in a host read of the winner), max over ranks"})
        # the sharded (max, argmax) must be what ONE GPU finds on the same candidates: 2^16 subset, outside any timed region
        sub = 1 << 16
        lo, hi = cdist.shard_bounds(sub, rank, world)
        got = cdist.sharded_argmax(acq_fn, Xall_cpu[lo:hi].to(dev), lo)
        if rank == 0:
            single = acq_fn(Xall_cpu[:sub].to(dev))
            argmax_check = bool(int(single.argmax()) == got[1] and float(single.max()) == got[0])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (CUDA-event times from the same timed region) ---------------------------
    names = ["prep_tables", "kernel_rows", "gemm_pass1_dmma", "gemm_pass2_dmma", "score_rank_integrate"]
    tot_ms = sum(ms[:5])
    cand_total = b * args.steps
    alg_flop = [0.0, b * args.steps * N_TRAIN * (3 * DIM + 12 + 6 * 8 + 8.0), cand_total * GEMM_FLOP_PER_CAND_PASS,
                cand_total * GEMM_FLOP_PER_CAND_PASS, cand_total * GRID * SCORE_FLOP_PER_EVAL]
    kernels = []
    for i, nm in enumerate(names):
        k = {"name": nm, "ms_per_step": ms[i] / args.steps, "share": ms[i] / tot_ms if tot_ms else None,
             "launches": int(nl[i])}
        if alg_flop[i] and ms[i] > 0:
            k["achieved_tflops"] = alg_flop[i] / (ms[i] * 1e-3) / 1e12
            k["frac_of_fp64_peak"] = k["achieved_tflops"] / FP64_PEAK_TFLOPS
        kernels.append(k)
    dom = max(range(5), key=lambda i: ms[i])
    per_launch_ms = ms[dom] / max(int(nl[dom]), 1)
    roofline = {
        "bound": "tensor", "pipe": "fp64 (DMMA.8x8x4 and DFMA share one pipe; no fp64 tcgen05 kind exists)",
        "kernel": names[dom], "achieved": kernels[dom].get("achieved_tflops"), "peak": FP64_PEAK_TFLOPS,
        "unit": "TFLOP/s", "frac": kernels[dom].get("frac_of_fp64_peak"),
        "peak_source": "measured on this pool (profiles/r01_fp64_peak.log); MEASURED_PEAKS.json has no fp64 figure",
        "avg_launch_ms": per_launch_ms,
        # dram__bytes_read.sum + dram__bytes_write.sum per launch of that kernel from one `ncu --set full` capture at the
        # bench's chunk size (65536 candidates / launch): profiles/r02_ncu_summary.txt (first block)
        "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(names[dom]),
        "whole_path": {"achieved": value / world * FLOP_PER_EVAL / 1e12, "frac": value / world * FLOP_PER_EVAL / 1e12 / FP64_PEAK_TFLOPS,
                       "flop_per_eval": FLOP_PER_EVAL},
        "hbm": {"achieved_gbs": value / world * BYTES_PER_EVAL / 1e9, "bytes_per_eval": BYTES_PER_EVAL, "peak_gbs": 6539.2,
                "frac": value / world * BYTES_PER_EVAL / 1e9 / 6539.2,
                # what the kernels really move (ncu dram bytes per 65536-candidate launch x launches per step / step time)
                "actual_dram_gbs": sum(NCU_DRAM_BYTES_PER_LAUNCH.values()) * (b / 65536.0) / (elapsed_ms / args.steps * 1e-3) / 1e9,
                "note": "path is fp64-compute bound, not HBM bound (SURVEY.md 8d): the U/V/W row round trips overlap the GEMMs"},
        "kernels": kernels,
    }
    line = {
        "metric": "conformal acq evals/sec (cands x grid)", "value": value, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic 2^20-candidate acquisition sweep at N=1024, d=8, grid=64 (BASELINE configs[4])"
                   if b == 1 << 20 else f"synthetic {b}-candidate acquisition sweep at N=1024, d=8, grid=64",
                   "n_train": N_TRAIN, "d": DIM, "grid": GRID, "candidates_per_rank_per_step": b, "acq": "cucb",
                   "temp": args.temp, "alpha": PARAMS["alpha"], "randomized": False, "ratio_estimator": None,
                   "parallelism": f"candidate-sharded x{world}, (max,argmax) all-gather",
                   "l2": "per-step working set (candidates + chunked U/V workspace) ~0.4 GB > 126 MB L2; no explicit flush"},
        "clocks": clocks, "gpu_launches": launches, "roofline": roofline,
    }
    if cfg2 is not None:
        line["other_configs"] = [c for c in (cfg2, cfg_ratio, cfg_grad, cfg_sgld) if c is not None]
    if strong:
        line["strong_scaling"] = strong
        line["argmax_check"] = argmax_check
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.temp)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
