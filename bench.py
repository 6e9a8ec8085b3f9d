#!/usr/bin/env python
"""bench.py -- BiCGSTAB-DILU time-to-solve / iterations per second on the SPE10-sized
(60x220x85 = 1.122M cells, 3x3 blocks) synthetic black-oil Jacobian of BASELINE.json, through the
C ABI of libopmb200.so.

One "step" = one per-Newton-step linear solve as Flow does it (NonlinearSystemBlackOilReservoir_
impl.hpp:459-469): refresh the Jacobian values + DILU refactorisation (prepare), then BiCGSTAB to
the reference's default reduction 1e-2 (solve).
  value  : Jacobian values, rhs and x already resident in HBM (device pointers)
  e2e    : the same call sequence with HOST buffers (pinned): H2D of values, x, b and D2H of x and
           the residual inside the timed region; `e2e.pageable` is the same with plain (pageable)
           numpy buffers, which the library page-locks once (cudaHostRegister, as
           gpuistl/ISTLSolverGPUISTL.hpp:429-432 does) -- what a Flow caller sees
Unit: Mcell-iterations/s = (global block rows x BiCGSTAB iterations) / second / 1e6, so that the
weak-scaling runs (one 60x220x85 slab per GPU, block-Jacobi DILU, halo + all-reduce) aggregate;
iterations/s and time-to-solve are given beside it.

After the timed regions, at EVERY N, the line carries a `parity` object: the solve converged, the
global true residual |b - A x| / |b| computed with the library's SpMV and owner-masked dot meets the
reduction, and on a reduced slab (60x220x8 per rank) iterations and solution are compared with the
oracle emulating the same ranks (orc.ParSystem).  With --gpus 8 (or --extra-configs) the other
multi-GPU configurations of BASELINE.json -- C4 400^3 as 400x400x50 per GPU, C5 200^3 4x4 blocks as
200x200x25 per GPU, DILU and ILU0 -- are run after the headline and reported under `configs`.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "Mcell-iterations/s"
METRIC = "bicgstab_dilu_cell_iterations_per_second"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples SM clocks / throttle reasons while the timed region runs: NVML in a thread every 5 ms (the timed
    region of the default run is ~150 ms: an `nvidia-smi -lms` child process would not have started by then);
    falls back to one `nvidia-smi` query if NVML is not importable"""
    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, device_index: int):
        self.dev = device_index
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._h = None
        try:
            import pynvml

            pynvml.nvmlInit()
            # torch's device ordinal follows CUDA_VISIBLE_DEVICES; NVML's does not
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = device_index
            if vis:
                try:
                    idx = int(vis.split(",")[device_index])
                except ValueError:
                    idx = device_index
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None

    def _sample(self):
        nv = self._nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for name, bit in self.REASONS:
            if mask & bit:
                self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                break
            self._stop.wait(0.005)

    def start(self):
        if self._h is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread:
            if not self.sm:  # a region shorter than the first sample: take one now, still under load
                try:
                    self._sample()
                except Exception:
                    pass
            self._stop.set()
            self._thread.join(timeout=2)
        elif self._h is None:
            try:
                q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                    "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
                r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.dev)],
                                   capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                self.sm, self.max_mhz = [float(r[0])], float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def workload(config, rank, world, nz_per_rank=None, seed_shift=0):
    """-> this rank's system: world == 1 the whole configuration, else one slab per rank stacked along z
    (Flow's contiguous partition of the natural ordering, ghost-last local numbering, one overlap layer)"""
    from opm_simulators_b200 import generators, partition

    cfg = dict(generators.CONFIGS[config])
    cfg.pop("n_active", None)
    cfg.pop("nnc", None)
    if nz_per_rank:
        cfg["nz"] = nz_per_rank
    nx, ny, nz = cfg["nx"], cfg["ny"], cfg["nz"]
    b = cfg["b"]
    if world == 1:
        s = generators.config(config) if not nz_per_rank else generators.blackoil_system(**cfg)
        A = s["A"]
        return dict(A=A, n_interior=A.n, halo=None, rhs=s["rhs2"], n_global=A.n, dims=(nx, ny, nz), b=b, cfg=cfg,
                    l2g=None)
    p = dict(cfg)
    p["nz"] = nz * world
    slab = generators.blackoil_system(z_range=(rank * nz, (rank + 1) * nz), with_rhs=False, **p)
    n_global = nx * ny * nz * world
    bounds = np.arange(world + 1) * (nx * ny * nz)
    owner_of = lambda g: np.searchsorted(bounds, g, side="right") - 1  # noqa: E731
    ls = partition.localize_rows(slab["row0"], slab["rowptr"], slab["gcol"], slab["val"], owner_of, rank)
    rhs = np.zeros(ls.n * ls.A.b)
    rr = np.random.Generator(np.random.Philox(key=[cfg["seed"] + seed_shift, 7000 + rank]))
    rhs[: ls.n_interior * ls.A.b] = rr.standard_normal(ls.n_interior * ls.A.b)
    return dict(A=ls.A, n_interior=ls.n_interior, halo=ls.halo, rhs=rhs, n_global=n_global, dims=(nx, ny, nz * world),
                b=b, cfg=p, l2g=ls.l2g)


class Dist:
    """torch.distributed plumbing (rendezvous, barriers, max over ranks): plumbing, not the product"""

    def __init__(self, gpus):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != gpus and self.world == 1 and gpus > 1:
            raise SystemExit("launch with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", self.local_rank))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def max(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def bcast(self, obj):
        if self.world == 1:
            return obj
        box = [obj]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]


def workload_string(config, dims, n_cells, prec, tol):
    """the same text in both arms (the driver compares the two `config.workload` strings)"""
    return (f"{config} {dims[0]}x{dims[1]}x{dims[2]} ({n_cells} cells, 3x3 BCSR) BiCGSTAB+{prec.upper()} tol={tol}, "
            f"step = value refresh + refactorisation + solve")


def make_solver(D, w, prec, tol, collectives, comm, extra=None):
    from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter

    opts = {"solver": "bicgstab", "tol": tol, "maxiter": 200, "verbosity": 0,
            "preconditioner": {"type": prec, "relaxation": 1.0}}
    env_extra = os.environ.get("OPMB200_BENCH_OPTS")  # experiments: extra "b200" keys as JSON, e.g. {"halo_overlap": 0}
    if env_extra:
        extra = dict(extra or {}, **json.loads(env_extra))
    if extra:
        opts["b200"] = extra
    fs = FlexibleSolver(MatrixAdapter(w["A"], w["n_interior"], comm, w["halo"]), opts)
    if D.world > 1 and collectives == "p2p":
        fs.enable_p2p(D.gather)
    return fs


def true_residual(fs, w, x_d, rhs_d):
    """|b - A x| / |b| over the owner rows of all ranks, with the library's own SpMV and scalar product
    (the solver's x is consistent on the ghosts, as Dune's is)"""
    import torch

    r = rhs_d.clone()
    fs.op.applyscaleadd(-1.0, x_d, r)  # r -= A x (ghost rows of r are zeroed by the operator)
    torch.cuda.synchronize()
    return float(np.sqrt(fs.dot(r, r) / max(fs.dot(rhs_d, rhs_d), 1e-300)))


def timed_steps(D, fs, step, steps, warmup, sampler=None):
    for k in range(warmup):
        step(k)
    D.barrier()
    if sampler:
        sampler.start()
    l0 = fs.info()["kernel_launches"]
    t0 = time.perf_counter()
    fs.timer_start()
    iters, t_upd, t_slv, last = 0, [], [], None
    for k in range(warmup, warmup + steps):
        last = step(k)
        iters += last.iterations
        i_ = fs.info()
        t_upd.append(i_["t_update_ms"])
        t_slv.append(i_["t_solve_ms"])
    ms_dev = fs.timer_stop()
    D.barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    return dict(ms_dev=D.max(ms_dev), wall=D.max(wall), iters=iters, t_upd=float(np.mean(t_upd)), t_slv=float(np.mean(t_slv)),
                launches=fs.info()["kernel_launches"] - l0, clocks=clocks, last=last)


def reduced_parity(D, args, comm):
    """a 60x220x8 slab per rank through the same code path, against the oracle emulating the same ranks in one
    process on rank 0 (orc.ParSystem: block-Jacobi, ghost-last local systems, owner-masked scalar products)"""
    import torch

    from opm_simulators_b200 import generators, partition

    w = workload("C3", D.rank, D.world, nz_per_rank=8)
    fs = make_solver(D, w, args.prec, args.tol, args.collectives, comm)
    b = w["b"]
    x_d = torch.zeros(len(w["rhs"]), dtype=torch.float64, device="cuda")
    r_d = torch.from_numpy(w["rhs"]).cuda()
    res = fs.apply(x_d, r_d)
    torch.cuda.synchronize()
    x_own = x_d.cpu().numpy().reshape(-1, b)[: w["n_interior"]]
    xs = D.gather(x_own)
    rhs_own = D.gather(w["rhs"].reshape(-1, b)[: w["n_interior"]])
    out = None
    if D.rank == 0:
        from oracle import oracle as orc

        p = dict(w["cfg"])
        full = generators.blackoil_system(with_rhs=False, **p)
        A = full["A"]
        rhs_g = np.concatenate(rhs_own).reshape(-1)
        if D.world > 1:
            part = partition.partition_simple(A.n, D.world)
            locs = [partition.localize(A, part, r) for r in range(D.world)]
            ps = orc.ParSystem([dict(rowptr=l.A.rowptr, col=l.A.col, val=l.A.val, interior=l.n_interior, l2g=l.l2g)
                                for l in locs], nglobal=A.n)
            bs = [l.scatter_global(rhs_g) for l in locs]
            for l, bb in zip(locs, bs):
                bb.reshape(-1, b)[l.n_interior:] = 0.0
        else:
            locs = None
            ps = orc.ParSystem.serial(A.rowptr, A.col, A.val)
            bs = [rhs_g]
        ps.prec_update(args.prec)
        xo, _, ro, _ = ps.bicgstab(bs, tol=args.tol, maxiter=200)
        xo_g = np.concatenate([(x.reshape(-1, b)[: l.n_interior] if locs else x.reshape(-1, b))
                               for x, l in zip(xo, locs or [None])]).reshape(-1)
        xg = np.concatenate(xs).reshape(-1)
        err = float(np.linalg.norm(xg - xo_g) / max(np.linalg.norm(xo_g), 1e-300))
        true = float(np.linalg.norm(rhs_g - A.to_scipy() @ xg) / np.linalg.norm(rhs_g))
        out = {"cells_per_rank": int(w["n_interior"]), "ranks": D.world, "iterations": res.iterations,
               "oracle_iterations": ro["iterations"], "x_rel_err": err, "true_reduction": true,
               "ok": bool(res.converged and abs(res.iterations - ro["iterations"]) <= 1 and err < 1e-8 and true < 1.01 * args.tol)}
    fs.close()
    return out


_WORKLOADS = {}


def measure_config(D, args, comm, name, config, nz_per_rank, prec, steps, warmup):
    """one extra multi-GPU configuration: device-resident steps + per-rank kernel times + correctness"""
    import torch

    t0 = time.perf_counter()
    key = (config, nz_per_rank)
    if key not in _WORKLOADS:  # DILU and ILU0 of a configuration share the generated slab
        _WORKLOADS.clear()
        _WORKLOADS[key] = workload(config, D.rank, D.world, nz_per_rank=nz_per_rank)
    w = _WORKLOADS[key]
    t_gen = time.perf_counter() - t0
    fs = make_solver(D, w, prec, args.tol, args.collectives, comm)
    info = fs.info()
    vals_d = torch.from_numpy(w["A"].val).cuda()
    rhs_d = torch.from_numpy(w["rhs"]).cuda()
    x_d = torch.zeros_like(rhs_d)
    r_d = torch.empty_like(rhs_d)

    def step(_k):
        fs.update(vals_d)
        x_d.zero_()
        r_d.copy_(rhs_d)
        torch.cuda.current_stream().synchronize()
        return fs.apply(x_d, r_d)

    T = timed_steps(D, fs, step, steps, warmup)
    true = true_residual(fs, w, x_d, rhs_d)
    conv = all(D.gather(bool(T["last"].converged)))
    peak, _ = peaks()
    kern = {}
    for what, nm in ((0, "spmv"), (4, "lower"), (5, "upper"), (2, "update")):
        ms, nbytes = fs.time_kernel(what, 2, 5)
        kern[nm] = {"ms": round(ms, 4), "frac_of_peak": round(nbytes / ms / 1e6 / peak, 3)}
    mine = {"rank": D.rank, "levels": info["n_levels"], "rows": info["n_rows"], "schedule": info["schedule"],
            **{k + "_ms": v["ms"] for k, v in kern.items()}}
    per_rank = D.gather(mine)
    b = w["b"]
    N, nnzb = info["n_rows"], info["nnzb"]
    prec_bytes = (nnzb * (8 * b * b + 4) + 32 * b * N + 16 * (N + 1) + 8 * N) if prec == "ilu0" else \
        ((nnzb - N) * (8 * b * b + 4) + 16 * b * b * N + 40 * b * N + 16 * (N + 1) + 8 * N)
    b_iter = 2 * (nnzb * (8 * b * b + 4) + 4 * (N + 1) + 16 * b * N) + 2 * prec_bytes + 19 * 8 * b * N
    it_ms = T["t_slv"] / max(T["iters"] / steps, 1)
    out = {"workload": f"{config} {w['dims'][0]}x{w['dims'][1]}x{w['dims'][2]} ({w['n_global']} cells, {b}x{b}) "
                       f"BiCGSTAB+{prec.upper()} tol={args.tol}", "cells_per_gpu": int(w["n_interior"]),
           "value": round(w["n_global"] * T["iters"] / (T["ms_dev"] * 1e-3) / 1e6, 2), "unit": UNIT,
           "ms_per_step": round(T["ms_dev"] / steps, 3), "iterations_per_solve": T["iters"] / steps,
           "update_ms": round(T["t_upd"], 3), "iteration_ms": round(it_ms, 4),
           "iteration_roofline_frac": round(b_iter / 1e6 / peak / it_ms, 4), "levels": info["n_levels"],
           "schedule": "tiles" if info["schedule"] == 1 else "levels",
           "converged": conv, "true_reduction": true, "ok": bool(conv and true < 1.01 * args.tol),
           "generate_s": round(t_gen, 1), "per_rank": per_rank}
    fs.close()
    del vals_d, rhs_d, x_d, r_d
    torch.cuda.empty_cache()
    return out


def run_b200(args):
    import torch

    from opm_simulators_b200 import _lib
    from opm_simulators_b200.flexible_solver import Communication

    D = Dist(args.gpus)
    rank, world, torch_ = D.rank, D.world, torch
    _lib.check(_lib.lib().opmb200_set_device(D.local_rank))
    comm = None
    if world > 1:
        comm = Communication(rank, world, D.bcast(Communication.unique_id() if rank == 0 else None))

    w = workload(args.config, rank, world)
    A, b = w["A"], w["b"]
    fs = make_solver(D, w, args.prec, args.tol, args.collectives, comm)
    info0 = fs.info()

    # ---- device-resident arm ---------------------------------------------------------------------
    vals_d = torch.from_numpy(A.val).cuda()
    rhs_d = torch.from_numpy(w["rhs"]).cuda()
    x_d = torch.zeros_like(rhs_d)
    r_d = torch.empty_like(rhs_d)

    def step_device(_k):
        fs.update(vals_d)
        x_d.zero_()
        r_d.copy_(rhs_d)
        torch_.cuda.current_stream().synchronize()
        return fs.apply(x_d, r_d)

    T = timed_steps(D, fs, step_device, args.steps, args.warmup, ClockSampler(D.local_rank) if rank == 0 else None)
    n_global = w["n_global"]
    value = n_global * T["iters"] / (T["ms_dev"] * 1e-3) / 1e6
    # ---- correctness of what was just timed, at every N ---------------------------------------------
    true = true_residual(fs, w, x_d, rhs_d)
    conv_all = all(D.gather(bool(T["last"].converged)))
    iters_all = D.gather(int(T["last"].iterations))
    # Dune's iteration count is (int) of a half-step counter: a solve reported as 10 iterations ran 10.5.  The cost of
    # one iteration is taken (a) from the executed half steps of the timed solves and (b) as the MARGINAL cost between
    # the timed solve and one long solve (tol 1e-6) of the same system, so that what is fixed per solve (staging, initial
    # defect, read-back) cancels.
    half_steps = len(fs.history()) - 1
    fs.update(vals_d)
    x_d.zero_()
    r_d.copy_(rhs_d)
    torch_.cuda.current_stream().synchronize()
    fs.apply(x_d, r_d, 1e-6)
    long_ms, long_half = fs.info()["t_solve_ms"], len(fs.history()) - 1

    # ---- end-to-end arms: host buffers through the same calls ----------------------------------------
    nb = args.steps + args.warmup
    vals_h = torch.from_numpy(A.val).pin_memory()
    rhs_h = [torch.from_numpy(w["rhs"]).clone().pin_memory() for _ in range(nb)]
    x_h = [torch.zeros(len(w["rhs"]), dtype=torch.float64).pin_memory() for _ in range(nb)]

    def step_host(k):
        fs.update(vals_h)
        return fs.apply(x_h[k], rhs_h[k])

    E = timed_steps(D, fs, step_host, args.steps, args.warmup)
    e2e_value = n_global * E["iters"] / E["wall"] / 1e6
    # pageable caller buffers (plain numpy, what Dune's BCRSMatrix / BlockVector storage is): the library
    # page-locks each buffer the first time it sees it
    vals_p = A.val
    rhs_p = [w["rhs"].copy() for _ in range(nb)]
    x_p = [np.zeros(len(w["rhs"])) for _ in range(nb)]
    rhs_keep = w["rhs"]

    def step_pageable(k):
        np.copyto(rhs_p[0], rhs_keep)  # Flow assembles into the same storage every Newton step
        x_p[0].fill(0.0)
        fs.update(vals_p)
        return fs.apply(x_p[0], rhs_p[0])

    P = timed_steps(D, fs, step_pageable, args.steps, args.warmup)
    page_value = n_global * P["iters"] / P["wall"] / 1e6
    vec_bytes = len(w["rhs"]) * 8
    h2d = A.val.nbytes + 2 * vec_bytes
    d2h = 2 * vec_bytes
    del vals_h, rhs_h, x_h

    # ---- roofline of the dominant kernel, measured live on the library's stream -------------------------
    peak, peak_src = peaks()
    tiles = info0["schedule"] == 1
    sweep_name = "tw_sweep_kernel" if tiles else "sweep_kernel"
    kern = {}
    for what, name in ((0, "spmv_kernel"), (4, f"{sweep_name}<lower>"), (5, f"{sweep_name}<upper>"),
                       (3, "vec_p_update+vec_half1+vec_half2"), (2, "relayout+factor" + ("+stream fill" if tiles else ""))):
        ms, nbytes = fs.time_kernel(what, 3, 20)
        kern[name] = {"ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1),
                      "GBps": round(nbytes / ms / 1e6, 1), "frac_of_peak": round(nbytes / ms / 1e6 / peak, 3)}
    # DRAM traffic per launch of the dominant kernel from the committed ncu --set full capture
    # (scripts/ncu_traffic.py); null when the capture is of another workload
    traffic = None
    caps = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_traffic.json")))
    if caps and args.config == "C3" and args.prec == "dilu" and world == 1:
        ks = json.load(open(caps[-1]))["kernels"]
        sw = [v["dram_bytes_per_launch"] for k, v in ks.items() if sweep_name in k and "<3" in k.replace("(int)", "")]
        if sw:
            traffic = round(sum(sw) / len(sw) / 1e6, 1)  # MB per launch, like `achieved`'s numerator
    mine = {"rank": rank, "levels": info0["n_levels"], "slices": info0["n_slices"], "rows": info0["n_rows"],
            "lower_ms": kern[f"{sweep_name}<lower>"]["ms"], "upper_ms": kern[f"{sweep_name}<upper>"]["ms"],
            "spmv_ms": kern["spmv_kernel"]["ms"]}
    per_rank = D.gather(mine) if world > 1 else None
    t_sweep = 0.5 * (kern[f"{sweep_name}<lower>"]["ms"] + kern[f"{sweep_name}<upper>"]["ms"])
    b_sweep = 0.5 * (kern[f"{sweep_name}<lower>"]["algorithmic_MB"] + kern[f"{sweep_name}<upper>"]["algorithmic_MB"])
    ach = b_sweep / t_sweep  # MB/ms == GB/s
    roofline = {"kernel": f"{sweep_name} (DILU lower/upper triangular sweep, "
                          f"{'tile walkers' if tiles else 'level schedule'}, 4 launches per iteration)",
                "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "traffic": traffic, "traffic_unit": "MB per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                "algorithmic_MB_per_launch": round(b_sweep, 1), "peak_source": peak_src, "per_kernel": kern}
    N, nnzb = info0["n_rows"], info0["nnzb"]
    b_prec = (nnzb * (8 * b * b + 4) + 32 * b * N + 16 * (N + 1) + 8 * N) if args.prec == "ilu0" \
        else ((nnzb - N) * (8 * b * b + 4) + 16 * b * b * N + 40 * b * N + 16 * (N + 1) + 8 * N)
    b_iter = 2 * (nnzb * (8 * b * b + 4) + 4 * (N + 1) + 16 * b * N) + 2 * b_prec + 19 * 8 * b * N
    fs.close()
    del vals_d, rhs_d, x_d, r_d
    torch.cuda.empty_cache()

    # ---- parity against the oracle at this N (reduced slab), extra configurations ----------------------
    reduced = reduced_parity(D, args, comm)
    configs = None
    if args.extra_configs or world == 8:
        configs = {}
        for name, config, nzr in (("C4", "C4", 50), ("C5", "C5", 25)):
            for prec in ("dilu", "ilu0"):
                try:
                    configs[f"{name}_{prec}"] = measure_config(D, args, comm, name, config, nzr, prec, 3, 2)
                except Exception as e:  # an extra configuration must not cost the headline line
                    configs[f"{name}_{prec}"] = {"error": repr(e)[:200]}

    out = None
    steps = args.steps
    if rank == 0:
        it_per = T["iters"] / steps
        out = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": round(T["ms_dev"] / steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(args.config, w["dims"], n_global, args.prec, args.tol),
                       "rhs": "N(0,1)", "cells_per_gpu": int(w["n_interior"]), "levels": info0["n_levels"],
                       "sweep_schedule": "tiles" if tiles else "levels",
                       "l2": "inputs larger than L2 (matrix 560 MB per GPU), no explicit flush",
                       "partition": (f"z-slabs, block-Jacobi DILU, halo + all-reduce over {args.collectives}, halo copy beside "
                                     f"the interior SpMV on a second stream") if world > 1 else "serial"},
            "iterations_per_solve": it_per, "iters_per_s": round(T["iters"] / (T["ms_dev"] * 1e-3), 2),
            "time_to_solve_ms": round(T["ms_dev"] / steps, 4), "wall_ms_per_step": round(T["wall"] * 1e3 / steps, 4),
            "update_ms": round(T["t_upd"], 4), "solve_ms": round(T["t_slv"], 4),
            "iteration_ms": round(T["t_slv"] / max(half_steps / 2, 0.5), 4),
            "iteration_ms_definition": "device time of a solve / executed iterations (half steps / 2), fixed per-solve work included",
            "iterations_executed_per_solve": half_steps / 2,
            "iteration_algorithmic_MB": round(b_iter / 1e6, 1),
            "iteration_roofline_frac": round((b_iter / 1e6 / peak) / (T["t_slv"] / max(half_steps / 2, 0.5)), 4),
            "iteration_marginal": (None if long_half <= half_steps else {
                "ms": round((long_ms - T["t_slv"]) / ((long_half - half_steps) / 2), 4),
                "roofline_frac": round((b_iter / 1e6 / peak) / ((long_ms - T["t_slv"]) / ((long_half - half_steps) / 2)), 4),
                "from": f"solve to 1e-6: {long_half / 2} iterations in {long_ms:.3f} ms against {half_steps / 2} in {T['t_slv']:.3f} ms"}),
            "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": round(E["wall"] * 1e3 / steps, 4),
                    "device_ms_per_step": round(E["ms_dev"] / steps, 4), "buffers": "pinned host memory",
                    "pageable": {"value": round(page_value, 3), "ms_per_step": round(P["wall"] * 1e3 / steps, 4),
                                 "buffers": "plain numpy (pageable) storage reused every step, page-locked once by the "
                                            "library (cudaHostRegister); the first registration is in the warm-up"}},
            "gpu_launches": int(T["launches"]), "clocks": T["clocks"], "roofline": roofline,
            "parity": {"converged_all_ranks": conv_all, "iterations_per_rank": iters_all,
                       "true_reduction": true, "requested_reduction": args.tol,
                       "iteration_convention": "Dune's iterations=(int)it and breakdown thresholds are restated "
                                               "(dune-istl is not in the reference tree): +-1 is against the restatement",
                       "reduced_slab_vs_oracle": reduced,
                       "ok": bool(conv_all and true < 1.01 * args.tol and len(set(iters_all)) == 1
                                  and reduced is not None and reduced["ok"])},
        }
        if per_rank:
            out["per_rank"] = per_rank
        if configs:
            out["configs"] = configs
        _WORKLOADS.clear()
    # ---- CPU baseline (rank 0, single-GPU run only) -------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_port_baseline(w, args)
    if comm:
        comm.close()
    if world > 1:
        D.dist.barrier()
        D.dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def cpu_port_baseline(w, args, max_iters=6):
    """the oracle (CPU restatement of the reference's FlexibleSolver BiCGSTAB+DILU path) on all host
    cores: R block-Jacobi subdomains, one OpenMP thread each == `mpirun -np R flow` with 1 thread per
    rank.  Bounded sample: factorisation + at most `max_iters` BiCGSTAB iterations of the same system."""
    from opm_simulators_b200 import partition
    from oracle import oracle as orc

    A = w["A"]
    R = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(R)
    t0 = time.perf_counter()
    if R > 1:
        part = partition.partition_simple(A.n, R)
        locs = [partition.localize(A, part, r) for r in range(R)]
        subs = [dict(rowptr=l.A.rowptr, col=l.A.col, val=l.A.val, interior=l.n_interior, l2g=l.l2g) for l in locs]
        ps = orc.ParSystem(subs, nglobal=A.n)
        bs = [l.scatter_global(w["rhs"]) for l in locs]
        for l, bb in zip(locs, bs):
            bb.reshape(-1, A.b)[l.n_interior:] = 0.0
    else:
        ps = orc.ParSystem.serial(A.rowptr, A.col, A.val)
        bs = [w["rhs"]]
    t_setup = time.perf_counter() - t0
    t0 = time.perf_counter()
    ps.prec_update(args.prec)
    _, _, res, _ = ps.bicgstab(bs, tol=args.tol, maxiter=max_iters)
    dt = time.perf_counter() - t0
    it = max(res["it"], 0.5)
    return {"value": round(A.n * it / dt / 1e6, 4), "unit": UNIT, "cores": R, "kind": "port",
            "sample": f"same {A.n}-cell system: DILU factorisation + {it} BiCGSTAB iterations (cap {max_iters}) "
                      f"on {R} block-Jacobi subdomains (1 OpenMP thread each), {dt:.1f} s; partition setup {t_setup:.1f} s not timed",
            "seconds": round(dt, 2)}


def run_reference(args):
    """the reference's OWN CPU implementation of the path: opm/simulators/linalg/mixed/{bsr,prec,bslv}.c
    (bslv_pbicgstab3d, double precision, DILU or ILU0), compiled unmodified into oracle/_ref.  It is a
    serial code (cores = 1).  Each step is the SAME work as a step of the GPU arm: factorisation + BiCGSTAB to
    the same reduction (no iteration cap), so the two arms are comparable; the R-core block-Jacobi port of the
    FlexibleSolver path is reported beside it by the default arm (`cpu_baseline`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from opm_simulators_b200 import generators
    from oracle import oracle as orc

    s = generators.config(args.config)
    A = s["A"]
    c_ = generators.CONFIGS[args.config]
    use_ref = orc.ref_available() and A.b == 3
    if use_ref:
        ref = orc.RefMixedSolver(A.rowptr, A.col, A.val, tol=args.tol, maxiter=200, use_dilu=(args.prec == "dilu"))
        run = lambda: ref.solve(s["rhs2"])[1]  # noqa: E731
        kind, cores = "reference", 1
        what = "opm/simulators/linalg/mixed bslv_pbicgstab3d (serial, unmodified)"
    else:
        ps = orc.ParSystem.serial(A.rowptr, A.col, A.val)

        def run():
            ps.prec_update(args.prec)
            return max(ps.bicgstab([s["rhs2"]], tol=args.tol, maxiter=200)[2]["it"], 0.5)
        kind, cores = "port", 1
        what = "oracle port (oracle/_ref unavailable)"
    steps = min(args.steps, 10)
    for _ in range(min(args.warmup, 1)):
        run()
    t0 = time.perf_counter()
    iters = 0.0
    for _ in range(steps):
        iters += run()
    dt = time.perf_counter() - t0
    value = A.n * iters / dt / 1e6
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus,
           "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": round(dt * 1e3 / steps, 3),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_string(args.config, (c_["nx"], c_["ny"], c_["nz"]), A.n, args.prec, args.tol),
                      "step_detail": "the reference solver takes the values, factorises and solves to the reduction in one call"},
           "iterations_per_solve": iters / steps,
           "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": kind,
                            "sample": f"{what}: each step = factorisation + {iters / steps:g} iterations to tol {args.tol} "
                                      f"of the full {A.n}-cell system (no cap); ONE core -- the reference's solver is serial"},
           "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C3")
    ap.add_argument("--prec", default="dilu", choices=["dilu", "ilu0"])
    ap.add_argument("--tol", type=float, default=1e-2)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--extra-configs", action="store_true",
                    help="also run C4 (400x400x50 per GPU) and C5 (200x200x25 per GPU, 4x4) after the headline (default at 8 GPUs)")
    ap.add_argument("--collectives", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: collectives inside the library's kernels over NVLink peer memory, or NCCL calls")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
