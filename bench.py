#!/usr/bin/env python
"""bench.py -- BiCGSTAB-DILU time-to-solve / iterations per second on the SPE10-sized
(60x220x85 = 1.122M cells, 3x3 blocks) synthetic black-oil Jacobian of BASELINE.json, through the
C ABI of libopmb200.so.

One "step" = one per-Newton-step linear solve as Flow does it (NonlinearSystemBlackOilReservoir_
impl.hpp:459-469): refresh the Jacobian values + DILU refactorisation (prepare), then BiCGSTAB to
the reference's default reduction 1e-2 (solve).
  value  : Jacobian values, rhs and x already resident in HBM (device pointers)
  e2e    : the same call sequence with HOST buffers (pinned): H2D of values, x, b and D2H of x and
           the residual inside the timed region
Unit: Mcell-iterations/s = (global block rows x BiCGSTAB iterations) / second / 1e6, so that the
weak-scaling runs (one 60x220x85 slab per GPU, block-Jacobi DILU, NCCL halo + all-reduce) aggregate;
iterations/s and time-to-solve are given beside it.

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
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.dev = device_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.dev)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload(args, rank, world):
    """-> (LocalSystem-like dict) for this rank"""
    from opm_simulators_b200 import generators, partition

    cfg = dict(generators.CONFIGS[args.config])
    cfg.pop("n_active", None)
    cfg.pop("nnc", None)
    nx, ny, nz = cfg["nx"], cfg["ny"], cfg["nz"]
    if world == 1:
        s = generators.config(args.config)
        A = s["A"]
        return dict(A=A, n_interior=A.n, halo=None, rhs=s["rhs2"], n_global=A.n, dims=(nx, ny, nz))
    # weak scaling: one (nx, ny, nz) slab per rank, stacked along z
    p = dict(cfg)
    p["nz"] = nz * world
    slab = generators.blackoil_system(z_range=(rank * nz, (rank + 1) * nz), with_rhs=False, **p)
    n_global = nx * ny * nz * world
    bounds = np.arange(world + 1) * (nx * ny * nz)
    owner_of = lambda g: np.searchsorted(bounds, g, side="right") - 1  # noqa: E731
    ls = partition.localize_rows(slab["row0"], slab["rowptr"], slab["gcol"], slab["val"], owner_of, rank)
    rhs = np.zeros(ls.n * ls.A.b)
    rr = np.random.Generator(np.random.Philox(key=[cfg["seed"], 7000 + rank]))
    rhs[: ls.n_interior * ls.A.b] = rr.standard_normal(ls.n_interior * ls.A.b)
    return dict(A=ls.A, n_interior=ls.n_interior, halo=ls.halo, rhs=rhs, n_global=n_global, dims=(nx, ny, nz * world))


def run_b200(args):
    import torch
    import torch.distributed as dist

    from opm_simulators_b200 import _lib
    from opm_simulators_b200.flexible_solver import Communication, FlexibleSolver, MatrixAdapter

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N")
    torch.cuda.set_device(local_rank)
    _lib.check(_lib.lib().opmb200_set_device(local_rank))
    comm = None
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
        ids = [Communication.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = Communication(rank, world, ids[0])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    w = workload(args, rank, world)
    A, b = w["A"], w["A"].b
    opts = {"solver": "bicgstab", "tol": args.tol, "maxiter": 200, "verbosity": 0,
            "preconditioner": {"type": args.prec, "relaxation": 1.0}}
    fs = FlexibleSolver(MatrixAdapter(A, w["n_interior"], comm, w["halo"]), opts)
    if world > 1 and args.collectives == "p2p":
        def allgather(blob):
            out = [None] * world
            dist.all_gather_object(out, blob)
            return out
        fs.enable_p2p(allgather)
    info0 = fs.info()

    # ---- device-resident arm ---------------------------------------------------------------------
    t_upd, t_slv = [], []
    vals_d = torch.from_numpy(A.val).cuda()
    rhs_d = torch.from_numpy(w["rhs"]).cuda()
    x_d = torch.zeros_like(rhs_d)
    r_d = torch.empty_like(rhs_d)

    def step_device():
        fs.update(vals_d)
        x_d.zero_()
        r_d.copy_(rhs_d)
        torch.cuda.current_stream().synchronize()
        return fs.apply(x_d, r_d)

    for _ in range(args.warmup):
        res = step_device()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = fs.info()["kernel_launches"]
    t0 = time.perf_counter()
    fs.timer_start()
    iters = 0
    for _ in range(args.steps):
        res = step_device()
        iters += res.iterations
        i_ = fs.info()
        t_upd.append(i_["t_update_ms"])
        t_slv.append(i_["t_solve_ms"])
    ms_dev = fs.timer_stop()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    launches = fs.info()["kernel_launches"] - l0
    ms_dev = max_over_ranks(ms_dev)
    wall = max_over_ranks(wall)
    n_global = w["n_global"]
    value = n_global * iters / (ms_dev * 1e-3) / 1e6

    # ---- end-to-end arm: host (pinned) buffers through the same calls ---------------------------------
    nb = args.steps + args.warmup
    vals_h = torch.from_numpy(A.val).pin_memory()
    rhs_h = [torch.from_numpy(w["rhs"]).clone().pin_memory() for _ in range(nb)]
    x_h = [torch.zeros(len(w["rhs"]), dtype=torch.float64).pin_memory() for _ in range(nb)]

    def step_host(k):
        fs.update(vals_h)
        return fs.apply(x_h[k], rhs_h[k])

    for k in range(args.warmup):
        step_host(k)
    barrier()
    fs.timer_start()
    t0 = time.perf_counter()
    iters_e = 0
    for k in range(args.warmup, nb):
        iters_e += step_host(k).iterations
    ms_e2e = fs.timer_stop()
    barrier()
    wall_e2e = max_over_ranks(time.perf_counter() - t0)
    ms_e2e = max(max_over_ranks(ms_e2e), 0.0)
    e2e_value = n_global * iters_e / (wall_e2e) / 1e6
    vec_bytes = len(w["rhs"]) * 8
    h2d = A.val.nbytes + 2 * vec_bytes
    d2h = 2 * vec_bytes

    # ---- roofline of the dominant kernel, measured live on the library's stream -------------------------
    peak, peak_src = peaks()
    kern = {}
    for what, name in ((0, "spmv_kernel"), (4, "sweep_kernel<lower>"), (5, "sweep_kernel<upper>"),
                       (3, "vec_p_update+vec_half1+vec_half2"), (2, "relayout+factor")):
        ms, nbytes = fs.time_kernel(what, 3, 20)
        kern[name] = {"ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1),
                      "GBps": round(nbytes / ms / 1e6, 1), "frac_of_peak": round(nbytes / ms / 1e6 / peak, 3)}
    # DRAM traffic per launch of the dominant kernel from the committed ncu --set full capture
    # (scripts/ncu_traffic.py); null when the capture is of another workload
    traffic = None
    caps = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_traffic.json")))
    if caps and args.config == "C3" and args.prec == "dilu" and world == 1:
        ks = json.load(open(caps[-1]))["kernels"]
        sw = [v["dram_bytes_per_launch"] for k, v in ks.items() if k.startswith("opmb200::sweep_kernel<3, false") or k.startswith("sweep_kernel<3, 0")]
        if sw:
            traffic = round(sum(sw) / len(sw) / 1e6, 1)  # MB per launch, like `achieved`'s numerator
    per_rank = None
    if world > 1:  # load balance of the dominant kernel across the ranks
        mine = {"rank": rank, "levels": info0["n_levels"], "slices": info0["n_slices"], "rows": info0["n_rows"],
                "lower_ms": kern["sweep_kernel<lower>"]["ms"], "upper_ms": kern["sweep_kernel<upper>"]["ms"],
                "spmv_ms": kern["spmv_kernel"]["ms"]}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    t_sweep = 0.5 * (kern["sweep_kernel<lower>"]["ms"] + kern["sweep_kernel<upper>"]["ms"])
    b_sweep = 0.5 * (kern["sweep_kernel<lower>"]["algorithmic_MB"] + kern["sweep_kernel<upper>"]["algorithmic_MB"])
    ach = b_sweep / t_sweep  # MB/ms == GB/s
    roofline = {"kernel": "sweep_kernel (DILU lower/upper triangular sweep, 4 launches per iteration)",
                "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "traffic": traffic, "traffic_unit": "MB per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                "algorithmic_MB_per_launch": round(b_sweep, 1), "peak_source": peak_src, "per_kernel": kern}
    N, nnzb = info0["n_rows"], info0["nnzb"]
    b_iter = 2 * (nnzb * (8 * b * b + 4) + 4 * (N + 1) + 16 * b * N) \
        + 2 * ((nnzb - N) * (8 * b * b + 4) + 16 * b * b * N + 40 * b * N + 16 * (N + 1) + 8 * N) + 19 * 8 * b * N
    last_solve_ms = fs.info()["t_solve_ms"]  # of the last e2e solve: includes the x/b staging copies
    out = None
    if rank == 0:
        out = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_dev / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.config} {w['dims'][0]}x{w['dims'][1]}x{w['dims'][2]} "
                                   f"({n_global} cells, 3x3 BCSR) BiCGSTAB+{args.prec.upper()} tol={args.tol}, "
                                   f"step = value refresh + refactorisation + solve",
                       "rhs": "N(0,1)", "cells_per_gpu": int(w["n_interior"]), "levels": info0["n_levels"],
                       "l2": "inputs larger than L2 (matrix 560 MB per GPU), no explicit flush",
                       "partition": (f"z-slabs, block-Jacobi DILU, halo + all-reduce over {args.collectives}") if world > 1 else "serial"},
            "iterations_per_solve": iters / args.steps, "iters_per_s": round(iters / (ms_dev * 1e-3), 2),
            "time_to_solve_ms": round(ms_dev / args.steps, 4), "wall_ms_per_step": round(wall * 1e3 / args.steps, 4),
            "update_ms": round(float(np.mean(t_upd)), 4), "solve_ms": round(float(np.mean(t_slv)), 4),
            "iteration_ms": round(float(np.mean(t_slv)) / max(iters / args.steps, 1), 4),
            "iteration_algorithmic_MB": round(b_iter / 1e6, 1),
            "iteration_roofline_frac": round((b_iter / 1e6 / peak) / (float(np.mean(t_slv)) / max(iters / args.steps, 1)), 4),
            "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": round(wall_e2e * 1e3 / args.steps, 4),
                    "device_ms_per_step": round(ms_e2e / args.steps, 4)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        }
        if per_rank:
            out["per_rank"] = per_rank
    # ---- CPU baseline (rank 0, single-GPU run only) -------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_port_baseline(w, args)
    fs.close()
    if comm:
        comm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
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
    serial code (cores = 1).  Each step is a bounded sample: factorisation + at most 3 iterations."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from opm_simulators_b200 import generators
    from oracle import oracle as orc

    s = generators.config(args.config)
    A = s["A"]
    cap = 3
    use_ref = orc.ref_available() and A.b == 3
    if use_ref:
        ref = orc.RefMixedSolver(A.rowptr, A.col, A.val, tol=args.tol, maxiter=cap, use_dilu=(args.prec == "dilu"))
        run = lambda: ref.solve(s["rhs2"])[1]  # noqa: E731
        kind, cores = "reference", 1
        what = "opm/simulators/linalg/mixed bslv_pbicgstab3d (serial, unmodified)"
    else:
        ps = orc.ParSystem.serial(A.rowptr, A.col, A.val)

        def run():
            ps.prec_update(args.prec)
            return max(ps.bicgstab([s["rhs2"]], tol=args.tol, maxiter=cap)[2]["it"], 0.5)
        kind, cores = "port", 1
        what = "oracle port (oracle/_ref unavailable)"
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    iters = 0.0
    for _ in range(args.steps):
        iters += run()
    dt = time.perf_counter() - t0
    value = A.n * iters / dt / 1e6
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3 / args.steps, 3),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"{args.config} {A.n} cells 3x3 BCSR BiCGSTAB+{args.prec.upper()} tol={args.tol}"},
           "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": kind,
                            "sample": f"{what}: each step = factorisation + {iters / args.steps:g} iterations "
                                      f"(cap {cap}) of the full system"},
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
