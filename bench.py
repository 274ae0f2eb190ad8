#!/usr/bin/env python
"""
bench.py -- the reference's hot path (extract -> assemble -> M^T A M / M^T b ->
BCs -> solve) on synthetic tensor-product B-spline patches.

  python bench.py --gpus N --steps K --warmup W            (our CUDA path)
  python bench.py --impl reference --gpus N --steps K ...  (CPU reference arm)

Metric (BASELINE.json): IGA DoF/s end to end.  One "step" = one complete pass
of the hot path over the patch.  Workload at N=1: BASELINE configs[1], 3-D
cubic B-spline Poisson on 256^3 cells (17.4 M IGA DoFs), element-fused path
(global A_FE would be 682 GB).

  value     device-timed throughput, control net / knots already resident in HBM
  e2e       same pass through the tIGAr API from HOST buffers (pinned control
            net H2D, solution vector D2H inside the timed region)
  roofline  the dominant kernel (windowed SpMV inside CG, HBM-bound), timed
            live with CUDA events on the solver stream
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
                 "--format=csv,noheader,nounits", "-lms", "200"],
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
    """Host-side inputs of one step: knot vectors (and, for callers that want the host
    array, the Greville control net -- the timed step generates it on the device)."""
    import numpy as np
    import torch
    from tIGAr.BSplines import ExplicitBSplineControlMesh, uniformKnots
    kv = [uniformKnots(P, 0.0, 1.0, nel) for _ in range(3)]
    cm = ExplicitBSplineControlMesh([P] * 3, kv)
    if not net:
        return kv, cm, None
    net = cm.controlNet()
    pinned = torch.from_numpy(net)
    if torch.cuda.is_available():
        pinned = pinned.pin_memory()
    return kv, cm, pinned


def one_step(kv, cm, control_net, mode, rtol, to_host):
    """One pass of the hot path through the tIGAr API.  Returns
    (n_dofs, cg_iterations, stage event list, result)."""
    import torch
    from tIGAr import (EqualOrderSpline, ExtractedSpline, TrialFunction, TestFunction,
                       Function, KrylovSolver, inner, sin, pi)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    gen = EqualOrderSpline(1, cm)
    sp = gen.getScalarSpline(0)
    for d in range(3):
        for side in (0, 1):
            gen.addZeroDofs(0, sp.getSideDofs(d, side))
    spline = ExtractedSpline(gen, 2 * P, mode=mode, controlNet=control_net)
    ks = KrylovSolver("cg", os.environ.get("TIGAR_B200_BENCH_PC", "fd"))
    ks.parameters["relative_tolerance"] = rtol
    spline.setSolverOptions(linearSolver=ks)
    ev[1].record()                                              # extract done
    u, v = TrialFunction(spline.V), TestFunction(spline.V)
    x = spline.spatialCoordinates()
    f = 3 * pi ** 2 * sin(pi * x[0]) * sin(pi * x[1]) * sin(pi * x[2])
    a = inner(spline.grad(u), spline.grad(v)) * spline.dx
    L = inner(f, v) * spline.dx
    MTAM, MTb = spline.assembleLinearSystem(a, L)
    ev[2].record()                                              # assemble + PtAP + BCs done
    uh = Function(spline.V)
    U = spline.solveLinearSystem(MTAM, MTb, uh)
    ev[3].record()                                              # solve done
    res = U.get_local() if to_host else U.t
    ev[4].record()
    return spline._patch.n_iga, spline.lastSolve["iterations"], ev, res, MTAM


def spmv_bytes(W):
    """Algorithmic bytes of one windowed SpMV launch: the exact non-zeros once
    (8 B/nnz, no column indices exist; SELL padding is NOT counted), x read + y
    write per row (+ row pointer for the row-major layout)."""
    return 8 * W.nnz + (16 if W.layout == 1 else 24) * W.nrows


# ------------------------------------------------------------------ CPU arm
_CPU_THREADS_USED = 1


def cpu_port_step(nel):
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
              "Poisson, %d^3 cells (%d DoFs), CG rtol %g, single-copy time %.2f s"
              % (procs, nel, nd, CG_RTOL, res[0][1]))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "3D cubic B-spline Poisson %d^3 cells" % args.nel,
                   "note": "CPU arm runs a bounded sample of the workload: " + sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": sample + "; numpy/scipy oracle (FEniCS/PETSc absent)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from tigar_b200._lib import lib
    import ctypes as C

    if world > 1:
        from tigar_b200 import multigpu
        return multigpu.bench(args, METRIC, UNIT, CG_RTOL)

    nel = args.nel
    mode = args.mode
    kv, cm, pinned = build_inputs(nel, net=False)

    def barrier():
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step(kv, cm, None, mode, CG_RTOL, False)
    barrier()

    # ---- device-resident timing (value) + live SpMV timing --------------------
    sampler = ClockSampler(local)
    sampler.start()
    lib.tg_prof_enable(1)
    l0 = lib.tg_launch_count()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    stages = []
    for _ in range(args.steps):
        n_dofs, iters, ev, _, MTAM = one_step(kv, cm, None, mode, CG_RTOL, False)
        stages.append(ev)
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)
    launches = lib.tg_launch_count() - l0
    spmv_ms, spmv_n = C.c_double(0), C.c_int64(0)
    lib.tg_prof_get(C.byref(spmv_ms), C.byref(spmv_n))
    lib.tg_prof_enable(0)
    clocks = sampler.stop()
    W = getattr(MTAM, "window", None)          # None: matrix-free operator (--mode matfree)
    stage_ms = {"extract": 0.0, "assemble_ptap_bcs": 0.0, "solve": 0.0}
    for ev in stages:
        stage_ms["extract"] += ev[0].elapsed_time(ev[1])
        stage_ms["assemble_ptap_bcs"] += ev[1].elapsed_time(ev[2])
        stage_ms["solve"] += ev[2].elapsed_time(ev[3])
    for k in stage_ms:
        stage_ms[k] /= max(args.steps, 1)
    del MTAM, stages
    value = n_dofs * args.steps / (ms * 1e-3)

    # ---- end to end from host buffers -----------------------------------------
    for _ in range(1):
        one_step(kv, cm, None, mode, CG_RTOL, True)
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        _, _, _, res, MT = one_step(kv, cm, None, mode, CG_RTOL, True)
        del MT
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - wall0))
    e2e = {"value": n_dofs * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": int(sum(len(k) for k in kv) * 8),
           "d2h_bytes_per_step": int(n_dofs * 8)}

    peak, which = measured_peaks()
    bytes_per_launch = spmv_bytes(W) if W is not None else 0
    avg_ms = spmv_ms.value / max(spmv_n.value, 1)
    achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    kname = {0: "k_win_spmv<true>", 1: "k_win_spmv_tma<true>", 2: "k_sell_spmv<7,true>", 3: "k_win_spmv_pf<true>"}[
        int(lib.tg_last_spmv_kind())]
    if W is None:
        kname = "none timed (matrix-free: tigar_qp + k_assemble_vector per CG iteration)"
    roofline = {"bound": "hbm", "kernel": kname + " (CG matvec + p.Ap partials)",
                "achieved": achieved, "peak": peak, "peak_source": which, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None,
                "bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_ms,
                "launches_timed": int(spmv_n.value),
                "share_of_step": spmv_ms.value / ms if ms > 0 else None}
    tfile = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    if os.path.exists(tfile):
        try:
            with open(tfile) as f:
                roofline["traffic"] = json.load(f).get(str(nel))
        except Exception:
            pass

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "3D cubic B-spline Poisson %d^3 cells, 1 GPU (BASELINE configs[1])"
                               % nel if nel == 256 else
                               "3D cubic B-spline Poisson %d^3 cells" % nel,
                   "degree": P, "cells": nel ** 3, "iga_dofs": n_dofs, "path": mode,
                   "quad_degree": 2 * P, "cg_rtol": CG_RTOL, "cg_iterations": iters,
                   "preconditioner": os.environ.get("TIGAR_B200_BENCH_PC", "fd"),
                   "l2": ("inputs larger than L2 (matrix %.1f GB streamed every CG iteration)"
                          % (8e-9 * W.nnz)) if W is not None else
                         "inputs larger than L2 (matrix-free: control net, operand and "
                         "coefficient chunks streamed every CG iteration)"},
        "stage_ms": stage_ms, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline}

    if not args.no_ptap:
        try:
            out["ptap"] = ptap_roofline(args.ptap_nel, peak)
        except Exception as e:                  # keep the headline line if the extra fails
            out["ptap"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if not args.no_cpu:
        nd, dt, its = cpu_port_step(args.cpu_nel)
        out["cpu_baseline"] = {
            "value": nd / dt, "unit": UNIT, "cores": _CPU_THREADS_USED, "kind": "port",
            "sample": "same problem at %d^3 cells (%d DoFs), one pass, %.1f s, CG its %d; "
                      "numpy/scipy oracle; cores = process CPU time / wall time (%d host "
                      "cores available)" % (args.cpu_nel, nd, dt, its, os.cpu_count() or 1)}
    print(json.dumps(out))


def ptap_roofline(nel, peak):
    """The global-CSR M^T A M (MatPtAP, common.py:1194-1195) at a 3-D cubic size
    whose operands are cheap to hold -- BASELINE.json's second metric.
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
    A = WinMatrix(patch.window("A"))
    A.vals.copy_(torch.rand(A.window.nnz, dtype=torch.float64, device="cuda"))
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
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nel", type=int, default=256)
    ap.add_argument("--mode", default="fused", choices=["fused", "csr", "matfree"])
    ap.add_argument("--cpu-nel", type=int, default=20)
    ap.add_argument("--ref-nel", type=int, default=16)
    ap.add_argument("--ref-procs", type=int, default=0, help="CPU arm: concurrent copies (0 = all cores)")
    ap.add_argument("--ptap-nel", type=int, default=64)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ptap", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    main()
