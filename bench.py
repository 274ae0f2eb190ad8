#!/usr/bin/env python
"""
bench.py -- the reference's hot path (extract -> assemble -> M^T A M / M^T b ->
BCs -> solve) on synthetic tensor-product B-spline patches.

  python bench.py --gpus N --steps K --warmup W            (our CUDA path)
  python bench.py --impl reference --gpus N --steps K ...  (CPU reference arm)

Metric (BASELINE.json): IGA DoF/s end to end.  One "step" = one complete pass of the hot path
over the patch, starting from the knot vectors: generators and Greville control net (made on
the device inside the step), 1-D tables, Gauss-point kernel, global sum-factorised assembly of
C = M^T A M and M^T b, homogeneous BCs, solve, solution vector.  Workload at N=1: BASELINE
configs[1], 3-D cubic B-spline Poisson on 256^3 cells (17.4 M IGA DoFs), element-fused path
(global A_FE would be 682 GB).  N>1: the same patch, slabs of IGA planes per
rank (strong scaling), launched by torchrun.

  value      device-timed throughput (max over ranks)
  e2e        same pass through the tIGAr API with the host->device copies of its inputs
             (knot vectors, zero-DoF lists) and the device->host read of the solution vector
             inside the timed region
  roofline   the kernel class with the largest share of the step, timed live with CUDA events
             on the launching stream inside the timed region; `rooflines` lists every class
             (HBM fraction on algorithmic bytes, FP64 fraction on algorithmic flops against
             the DFMA peak measured by tg_fp64_peak)
  parity     checks of the LAST timed step printed in the line: true residual |b - C U|/|b|
             recomputed with an independent SpMV, checksums of U, L2 error against the
             manufactured solution -- identical across N up to round-off
  cpu_baseline  the numpy/scipy oracle ("port") on a bounded sample, host cores
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "DoF/s end-to-end (extract+assemble+PtAP+solve)"
UNIT = "DoF/s"
P = 3
CG_RTOL = 1e-10


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs")), "measured"
    except Exception:
        return 6650.0, "fallback"


# ------------------------------------------------------------------ clocks
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------ problem
def build_inputs(nel, net=True):
    """Host-side inputs of one step: knot vectors and the control-mesh object (and, for
    callers that want the host array, the Greville control net -- the timed step generates it
    on the device)."""
    import torch
    from tIGAr.BSplines import ExplicitBSplineControlMesh, uniformKnots
    kv = [uniformKnots(P, 0.0, 1.0, nel) for _ in range(3)]
    cm = ExplicitBSplineControlMesh([P] * 3, kv)
    if not net:
        return kv, cm, None
    pinned = torch.from_numpy(cm.controlNet())
    if torch.cuda.is_available():
        pinned = pinned.pin_memory()
    return kv, cm, pinned


def bench_pc():
    return os.environ.get("TIGAR_B200_BENCH_PC", "fd")


WORKLOAD = ["explicit"]        # "annulus": BASELINE configs[3], cubic NURBS quarter annulus


def build_annulus(nel):
    """configs[3]: the igakit-like NURBS object of the quarter annulus (the workload's INPUT;
    built once on the host, its homogeneous control net is uploaded inside the timed e2e step)
    wrapped in the reference's NURBSControlMesh."""
    from tIGAr.NURBS import NURBSControlMesh, quarter_annulus
    return NURBSControlMesh(quarter_annulus(P, [nel] * 3, 3))


def one_step(kv, cm, control_net, mode, rtol, to_host, keep=False):
    """One pass of the hot path through the tIGAr API.  ``control_net`` None: the generator
    makes the Greville net on the device (the default of the product); ``cm`` None: the
    control-mesh object is built inside the step as well.  Returns
    (n_dofs, cg_iterations, stage event list, result, MTAM[, extras])."""
    import torch
    from tIGAr import (EqualOrderSpline, ExtractedSpline, TrialFunction, TestFunction,
                       Function, KrylovSolver, inner, sin, pi)
    from tIGAr.BSplines import ExplicitBSplineControlMesh
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    annulus = WORKLOAD[0] == "annulus"
    if cm is None:
        cm = ExplicitBSplineControlMesh([P] * 3, kv)
    gen = EqualOrderSpline(1, cm)
    sp = gen.getScalarSpline(0)
    for d in range(3):
        for side in (0, 1):
            gen.addZeroDofs(0, sp.getSideDofs(d, side))
    spline = ExtractedSpline(gen, 2 * P, mode=mode, controlNet=control_net)
    ks = KrylovSolver("cg", bench_pc())
    ks.parameters["relative_tolerance"] = rtol
    spline.setSolverOptions(linearSolver=ks)
    ev[1].record()                                              # extract done
    u, v = TrialFunction(spline.V), TestFunction(spline.V)
    x = spline.spatialCoordinates()
    if annulus:
        # u = (r-1)(2-r) sin(2 theta) sin(pi z): zero on the whole boundary of the patch;
        # f = -div grad u built symbolically as poisson-nurbs.py:127-133 does
        from tIGAr import sqrt
        u, v = spline.rationalize(u), spline.rationalize(v)
        r = sqrt(x[0] * x[0] + x[1] * x[1])
        soln = (r - 1.0) * (2.0 - r) * (2.0 * x[0] * x[1] / (r * r)) * sin(pi * x[2])
        a = inner(spline.grad(u), spline.grad(v)) * spline.dx
        L = inner(-spline.div(spline.grad(soln)), v) * spline.dx
    else:
        soln = sin(pi * x[0]) * sin(pi * x[1]) * sin(pi * x[2])
        a = inner(spline.grad(u), spline.grad(v)) * spline.dx
        L = inner(3 * pi ** 2 * soln, v) * spline.dx
    MTAM, MTb = spline.assembleLinearSystem(a, L)
    ev[2].record()                                              # assemble + PtAP + BCs done
    uh = Function(spline.V)
    U = spline.solveLinearSystem(MTAM, MTb, uh)
    ev[3].record()                                              # solve done
    res = U.get_local() if to_host else U.t
    ev[4].record()
    out = (spline._patch.n_iga, spline.lastSolve["iterations"], ev, res, MTAM)
    if keep:
        # (the Function object itself must stay alive: forms refer to it by id)
        out = out + (dict(spline=spline, MTb=MTb, U=U, soln=soln, uh_function=uh,
                          uh=spline.rationalize(uh) if annulus else uh),)
    return out


def spmv_bytes(W):
    """Algorithmic bytes of one windowed SpMV launch: the exact non-zeros once
    (8 B/nnz, no column indices exist; SELL padding is NOT counted), x read + y
    write per row (+ row pointer for the row-major layout)."""
    return 8 * W.nnz + (16 if W.layout == 1 else 24) * W.nrows


def parity_checks(ex, MTAM, world):
    """Driver-visible parity of the last timed step (outside the timed region): the true
    residual with an independent SpMV on this rank's rows, checksums of the (replicated) IGA
    DoF vector and the L2 error against the manufactured solution."""
    import torch
    from tIGAr import assemble
    spline, U, b = ex["spline"], ex["U"].t, ex["MTb"].t
    W = getattr(MTAM, "window", None)
    out = {}
    if W is not None:
        p = spline._patch
        if p.part is not None:
            pp, pl = p.pp, p.plane
            xe = U[pp["c0"] * pl:pp["c1"] * pl].contiguous()
        else:
            xe = U
        r = MTAM.matvec(xe)
        r.sub_(b)
        s = torch.stack([(r * r).sum(), (b * b).sum()])
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(s)
        out["true_relative_residual"] = math.sqrt(float(s[0]) / max(float(s[1]), 1e-300))
    out["sum_U"] = float(U.sum())
    out["norm_U"] = float(U.norm())
    out["l2_error_vs_manufactured"] = math.sqrt(abs(assemble(((ex["uh"] - ex["soln"]) ** 2)
                                                             * spline.dx)))
    out["solver"] = spline.lastSolve.get("method")
    out["solver_relative_residual"] = spline.lastSolve.get("relative_residual")
    return out


# ------------------------------------------------------------------ CPU arm
_CPU_THREADS_USED = 1


def cpu_port_step(nel, stages=None):
    import numpy as np
    from oracle import pipeline as OP
    from oracle import bsplines as OB
    kv = [OB.uniform_knots(P, 0.0, 1.0, nel)] * 3
    pr = OP.Problem([P] * 3, kv)
    f = lambda X: 3 * math.pi ** 2 * np.prod(np.sin(math.pi * X), axis=-1)
    t, c = time.perf_counter(), time.process_time()
    pr.extract()
    pr.assemble(f)
    pr.ptap()
    pr.solve("cg", CG_RTOL)
    dt = time.perf_counter() - t
    global _CPU_THREADS_USED          # process CPU time / wall time of the last CPU step
    _CPU_THREADS_USED = max(1, int(round((time.process_time() - c) / max(dt, 1e-9))))
    if stages is not None:
        stages.update({k: round(v, 3) for k, v in pr.times.items()})
    return pr.ts.ncp, dt, pr.iters


def _ref_worker(nel):
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    nd, dt, its = cpu_port_step(nel)
    return nd, dt, its


def run_reference(args):
    """CPU arm.  The numpy/scipy port is single-threaded (like one PETSc rank), so to
    use all host cores each step runs one independent copy of the bounded sample per
    core concurrently -- the ideal-scaling bound of an MPI run of the reference
    (SURVEY 8d "divide by 8 ideal") -- and counts the DoFs of all copies."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import concurrent.futures as cf
    import multiprocessing as mp
    nel = args.ref_nel
    procs = args.ref_procs
    if procs <= 0:                       # all cores, bounded by memory (~1.5 GB per 16^3 copy)
        procs = min(os.cpu_count() or 1, 32)
        try:
            import psutil
            procs = max(1, min(procs, int(psutil.virtual_memory().available // (2.5 * 2 ** 30))))
        except Exception:
            pass
    ctx = mp.get_context("fork")
    with cf.ProcessPoolExecutor(max_workers=procs, mp_context=ctx) as ex:
        def step():
            return [f.result() for f in [ex.submit(_ref_worker, nel) for _ in range(procs)]]
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        n = 0
        for _ in range(args.steps):
            res = step()
            n += sum(r[0] for r in res)
        dt = time.perf_counter() - t0
    nd = res[0][0]
    val = n / dt
    sample = ("%d concurrent independent copies (one per host core) of: 3-D cubic B-spline "
              "Poisson, %d^3 cells (%d DoFs), Jacobi-CG rtol %g, single-copy time %.2f s"
              % (procs, nel, nd, CG_RTOL, res[0][1]))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "3D cubic B-spline Poisson %d^3 cells" % args.nel,
                   "note": "CPU arm runs a bounded sample of the workload: " + sample
                           + ".  The per-DoF cost of the port grows with size (732 DoF/s per "
                             "core at 32^3 vs 790 at 16^3, DESIGN.md 5), so the small sample "
                             "flatters the CPU arm"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": sample + "; numpy/scipy oracle (FEniCS/PETSc absent)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------ GPU arm
def rooflines_from(prof, steps, ms_total, hbm_peak, fp64_peak):
    rows = []
    for name, d in prof.items():
        if d["launches"] == 0 or d["ms"] <= 0:
            continue
        avg_ms = d["ms"] / d["launches"]
        bpl = d["bytes"] / d["launches"]
        fpl = d["flops"] / d["launches"]
        gbs = bpl / (avg_ms * 1e-3) / 1e9
        tfs = fpl / (avg_ms * 1e-3) / 1e12
        rows.append({"kernel": name, "launches_per_step": d["launches"] / max(steps, 1),
                     "ms_per_step": d["ms"] / max(steps, 1), "avg_launch_ms": avg_ms,
                     "bytes_per_launch": bpl, "achieved_gbs": gbs, "hbm_frac": gbs / hbm_peak,
                     "flops_per_launch": fpl, "achieved_tflops": tfs,
                     "fp64_frac": (tfs / fp64_peak) if fp64_peak else None,
                     "share_of_step": d["ms"] / ms_total if ms_total > 0 else None})
    rows.sort(key=lambda r: -r["ms_per_step"])
    return rows


def run_ours(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from tigar_b200._lib import lib, check
    from tigar_b200 import dev

    nel, mode = args.nel, args.mode
    if world > 1 and mode != "fused":
        raise SystemExit("multi-GPU runs use the element-fused path")
    kv, cm, _ = build_inputs(nel, net=False)
    WORKLOAD[0] = args.workload
    net_bytes = 0
    if args.workload == "annulus":
        cm = build_annulus(nel)
        net_bytes = int(cm.bnet.size * 8)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    fp64 = C.c_double(0.0)
    check(lib.tg_fp64_peak(dev.ptr(dev.zeros(1)), C.byref(fp64), dev.stream()))
    fp64_peak = fp64.value
    hbm_peak, which = measured_peaks()

    for _ in range(args.warmup):
        one_step(kv, cm, None, mode, CG_RTOL, False)
    barrier()

    # ---- device timing (value) with live per-kernel event timing ------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.tg_prof_enable(1)
    dev.PROF.start()
    l0 = lib.tg_launch_count()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    stages = []
    r = MTAM = None
    for k in range(args.steps):
        del r, MTAM
        r = one_step(kv, cm, None, mode, CG_RTOL, False, keep=(k == args.steps - 1))
        n_dofs, iters, ev, _, MTAM = r[:5]
        stages.append(ev)
    t1.record()
    barrier()
    ms = maxr(t0.elapsed_time(t1))
    launches = lib.tg_launch_count() - l0
    prof = dev.PROF.stop()
    spmv_ms, spmv_n = C.c_double(0), C.c_int64(0)
    lib.tg_prof_get(C.byref(spmv_ms), C.byref(spmv_n))
    lib.tg_prof_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    extras = r[5]
    W = getattr(MTAM, "window", None)          # None: matrix-free operator (--mode matfree)
    if spmv_n.value > 0 and W is not None:     # Jacobi-CG inside the C driver
        prof["k_win_spmv (CG matvec + p.Ap)"] = dict(launches=int(spmv_n.value),
                                                      ms=spmv_ms.value,
                                                      bytes=spmv_bytes(W) * spmv_n.value,
                                                      flops=2.0 * W.nnz * spmv_n.value)
    stage_ms = [0.0, 0.0, 0.0]
    for ev in stages:
        for i in range(3):
            stage_ms[i] += ev[i].elapsed_time(ev[i + 1]) / max(args.steps, 1)
    stage_ms = [maxr(v) for v in stage_ms]
    parity = parity_checks(extras, MTAM, world)
    nz = len(extras["spline"]._zeroDofsRaw)
    if net_bytes:      # bytes of the control net each rank actually uploads (its slab + halo)
        mine = float(getattr(extras["spline"], "_h2d_bytes", net_bytes))
        if world > 1:
            t = torch.tensor([mine], dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            mine = float(t.item())
        net_bytes_all = int(mine)
    else:
        net_bytes_all = 0
    local_nnz = W.nnz if W is not None else 0
    del MTAM, stages, extras, r
    value = n_dofs * args.steps / (ms * 1e-3)

    # ---- end to end: host inputs in, host solution out ------------------------------------
    cm_e2e = cm if args.workload == "annulus" else None     # explicit: rebuilt from the knots
    one_step(kv, cm_e2e, None, mode, CG_RTOL, rank == 0)
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        rr = one_step(kv, cm_e2e, None, mode, CG_RTOL, rank == 0)
        del rr
    e1.record()
    barrier()
    e2e_ms = maxr(max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - wall0)))
    e2e = {"value": n_dofs * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": int((sum(len(k) for k in kv) * 8 + nz * 8) * world
                                     + net_bytes_all),
           "d2h_bytes_per_step": int(n_dofs * 8),
           "note": ("inputs = knot vectors + zero-DoF lists (every rank uploads its copy); the "
                    "Greville control net is generated on the device inside the step; output = "
                    "the IGA DoF vector read back by rank 0") if not net_bytes else
                   ("inputs = the NURBS object's homogeneous control net (every rank uploads "
                    "the planes its cell layers read: its slab + halo), knot vectors, zero-DoF "
                    "lists; output = the IGA DoF vector read back by rank 0")}
    if rank != 0:
        dist.barrier()
        dist.destroy_process_group()
        return

    rows = rooflines_from(prof, args.steps, ms, hbm_peak, fp64_peak)
    traffic = {}
    tfile = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    if os.path.exists(tfile):
        try:
            with open(tfile) as f:
                traffic = json.load(f).get("%d_n%d" % (nel, world), {})
        except Exception:
            traffic = {}
    for r_ in rows:
        r_["traffic"] = traffic.get(r_["kernel"])
    if rows:
        # the dominant kernel: largest share of the step among the classes whose limiter is HBM
        # (the FP64-bound Gauss-point kernel is listed in `rooflines` with its FP64 fraction)
        hb = [r_ for r_ in rows if r_["hbm_frac"] >= (r_["fp64_frac"] or 0.0)]
        top = hb[0] if hb else rows[0]
        hbm_bound = top["hbm_frac"] >= (top["fp64_frac"] or 0.0)
        roofline = {"bound": "hbm", "kernel": top["kernel"], "achieved": top["achieved_gbs"],
                    "peak": hbm_peak, "peak_source": which, "unit": "GB/s",
                    "frac": top["hbm_frac"], "traffic": top["traffic"],
                    "traffic_source": ("ncu --set full dram__bytes per launch, profiles/"
                                       "r2_ncu_traffic.json") if top["traffic"] else None,
                    "bytes_per_launch": top["bytes_per_launch"],
                    "avg_launch_ms": top["avg_launch_ms"],
                    "launches_timed": int(round(top["launches_per_step"] * args.steps)),
                    "share_of_step": top["share_of_step"],
                    "fp64_frac": top["fp64_frac"], "fp64_peak_tflops": fp64_peak,
                    "limiter": "hbm" if hbm_bound else "fp64 pipe (no FP64 tensor path on B200)"}
    else:
        roofline = {"bound": "hbm", "kernel": "none timed", "achieved": 0.0, "peak": hbm_peak,
                    "peak_source": which, "unit": "GB/s", "frac": 0.0, "traffic": None}

    if args.workload == "annulus":
        wl = ("3D cubic NURBS Poisson on the quarter annulus, %d^3 cells, %d GPU%s (BASELINE "
              "configs[3]%s)" % (nel, world, "" if world == 1 else "s",
                                 "" if nel == 512 else " geometry at reduced size"))
    else:
        wl = "3D cubic B-spline Poisson %d^3 cells" % nel
        if nel == 256:
            wl += ", %d GPU%s (BASELINE configs[1])" % (world, "" if world == 1 else "s")
    pc = bench_pc()
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl, "degree": P, "cells": nel ** 3, "iga_dofs": n_dofs,
                   "path": mode, "quad_degree": 2 * P, "cg_rtol": CG_RTOL,
                   "cg_iterations": iters,
                   "preconditioner": {"fd": "fast diagonalisation (tensor-product eigenbasis)",
                                      "jacobi": "jacobi"}.get(pc, pc),
                   "partition": None if world == 1 else
                   "slabs of IGA planes (last direction), p halo cell layers recomputed, "
                   "row-distributed CG over NCCL (halo send/recv, all-reduced scalars, "
                   "all-to-all transposes inside the FD preconditioner)",
                   "l2": ("inputs larger than L2 (local matrix %.1f GB, intermediates of the "
                          "assembly streamed once)" % (8e-9 * local_nnz)) if W is not None else
                         "inputs larger than L2 (matrix-free: control net, operand and "
                         "coefficient chunks streamed every CG iteration)"},
        "stage_ms": {"extract": stage_ms[0], "assemble_ptap_bcs": stage_ms[1],
                     "solve": stage_ms[2]},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "parity": parity,
        "roofline": roofline, "rooflines": rows,
        "fp64_peak_tflops_measured": fp64_peak}

    gs = [r_ for r_ in rows if r_["kernel"].startswith("k_gsf matrix")]
    if gs:
        t = sum(r_["ms_per_step"] for r_ in gs)
        byts = sum(r_["bytes_per_launch"] * r_["launches_per_step"] for r_ in gs)
        out["ptap_fused"] = {
            "what": "M^T A M of the production (element-fused) path = the march stages: "
                    "algorithmic bytes of all stages / their summed device time",
            "ms": t, "bytes": byts, "achieved": byts / (t * 1e-3) / 1e9, "unit": "GB/s",
            "frac": byts / (t * 1e-3) / 1e9 / hbm_peak, "peak": hbm_peak}
    if world == 1 and not args.no_ptap:
        torch.cuda.empty_cache()
        try:
            out["ptap"] = ptap_roofline(args.ptap_nel, hbm_peak)
        except Exception as e:                  # keep the headline line if the extra fails
            out["ptap"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if world == 1 and not args.no_cpu:
        st = {}
        nd, dt, its = cpu_port_step(args.cpu_nel, st)
        out["cpu_baseline"] = {
            "value": nd / dt, "unit": UNIT, "cores": _CPU_THREADS_USED, "kind": "port",
            "stage_s": st,
            "sample": "same problem at %d^3 cells (%d DoFs), one pass, %.1f s, Jacobi-CG its %d; "
                      "numpy/scipy oracle; cores = process CPU time / wall time (%d host "
                      "cores available)" % (args.cpu_nel, nd, dt, its, os.cpu_count() or 1)}
    print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ptap_roofline(nel, peak):
    """The global-CSR M^T A M (MatPtAP, common.py:1194-1195) at a 3-D cubic size
    whose operands fit one GPU -- BASELINE.json's second metric.
    ``achieved`` uses BASELINE.md's algorithmic bytes (CSR(A) + CSR(M) + CSR(C)
    with 8 B value + 4 B column per nnz + 4 B row pointer per row, each operand
    once); ``streamed`` is what the three march passes move by design (8 B/value,
    no column indices, M never read: A read, two shrinking intermediates written
    and read, C written)."""
    import torch
    from tigar_b200.engine import TensorPatch, WinMatrix
    from tIGAr.BSplines import uniformKnots
    kv = [uniformKnots(P, 0.0, 1.0, nel)] * 3
    patch = TensorPatch([P] * 3, kv)
    wA = patch.window("A")
    A = WinMatrix(wA, torch.empty(wA.nnz, dtype=torch.float64, device="cuda"))
    chunk = 1 << 28
    for o in range(0, wA.nnz, chunk):             # random fill without a second 57 GB tensor
        A.vals[o:o + chunk].uniform_(0.0, 1.0)
    ts = []
    for rep in range(6):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record()
        Cm = patch.ptap(A)
        e[1].record()
        torch.cuda.synchronize()
        ts.append(e[0].elapsed_time(e[1]))
        del Cm
    ms = min(ts[3:])
    wA, wM, wC = (patch.window(k) for k in "AMC")
    csr = lambda w: 12 * w.nnz + 4 * (w.nrows + 1)
    alg = csr(wA) + csr(wM) + csr(wC)
    out = {"ms": ms, "bytes": alg, "achieved": alg / (ms * 1e-3) / 1e9, "unit": "GB/s",
           "frac": alg / (ms * 1e-3) / 1e9 / peak, "peak": peak}
    march = patch._march_setup()
    if march is not None and os.environ.get("TIGAR_B200_PTAP", "march") == "march":
        inter = [P_["wY"].nnz for P_ in march[1][:-1]]
        streamed = 8 * (wA.nnz + 2 * sum(inter) + wC.nnz)
        out["workload"] = ("3D cubic %d^3 global M^T A M: one two-sided march pass per direction "
                           "(k_ptap_march_w x3), A read once, M never formed" % nel)
        out["nnz"] = {"A": wA.nnz, "M": wM.nnz, "Z0": inter[0], "Z1": inter[1], "C": wC.nnz}
    else:
        wP = patch.window("P")
        inter = [patch._win[k].nnz for k in ("K0", "K1") if k in patch._win]
        streamed = 8 * (wA.nnz + 2 * wP.nnz + 2 * sum(inter) + wC.nnz)
        out["workload"] = "3D cubic %d^3 global M^T A M (row-wise Kronecker kernels)" % nel
        out["nnz"] = {"A": wA.nnz, "M": wM.nnz, "AP": wP.nnz, "C": wC.nnz}
    out["streamed_bytes"] = streamed
    out["streamed_gbs"] = streamed / (ms * 1e-3) / 1e9
    out["streamed_frac"] = out["streamed_gbs"] / peak
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nel", type=int, default=256)
    ap.add_argument("--mode", default="fused", choices=["fused", "csr", "matfree"])
    ap.add_argument("--workload", default="explicit", choices=["explicit", "annulus"],
                    help="explicit: BASELINE configs[1] (default); annulus: configs[3]")
    ap.add_argument("--cpu-nel", type=int, default=24)
    ap.add_argument("--ref-nel", type=int, default=16)
    ap.add_argument("--ref-procs", type=int, default=0, help="CPU arm: concurrent copies (0 = all cores)")
    ap.add_argument("--ptap-nel", type=int, default=128)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ptap", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    main()
