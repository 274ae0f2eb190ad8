#!/usr/bin/env python
"""CPU baselines of BASELINE.md section 3, per stage, on THIS host's cores (no GPU):

  B0 "faithful"   the reference's own way to build M (common.py:1497-1509): a Python loop
                  over FE nodes -> getNodesAndEvals -> scalar sparse insertion
                  (oracle.extraction.build_M_loop)
  B1 "strong CPU" the vectorised numpy/scipy oracle, stage by stage
                  (extract / assemble / PtAP+BCs / solve), best of ``--reps`` after a warm-up

Usage: python tools/cpu_baseline.py [--reps 3] > profiles/rN_cpu_baseline_stages.json
The oracle is test infrastructure; this script only times it (like bench.py's CPU legs).
"""
import argparse
import json
import math
import os
import platform
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                      # noqa: E402
from oracle import bsplines as OB       # noqa: E402
from oracle import extraction as OX     # noqa: E402
from oracle import pipeline as OP       # noqa: E402


def host():
    model = ""
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    model = ln.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return {"cores": os.cpu_count(), "cpu": model or platform.processor(),
            "python": platform.python_version(), "numpy": np.__version__}


def rhs(dim, form):
    if form == "biharmonic":
        return lambda X: 4 * math.pi ** 4 * np.prod(np.cos(math.pi * X[..., :dim]), axis=-1)
    return lambda X: dim * math.pi ** 2 * np.prod(np.sin(math.pi * X[..., :dim]), axis=-1)


def b1(deg, nel, form, method, reps, lo=0.0, hi=1.0, nLayers=1):
    kv = [OB.uniform_knots(p, lo, hi, n) for p, n in zip(deg, nel)]
    best = None
    for rep in range(reps + 1):
        pr = OP.Problem(deg, kv, form=form, nLayers=nLayers)
        t0 = time.perf_counter()
        pr.extract()
        pr.assemble(rhs(len(deg), form))
        pr.ptap()
        pr.solve(method, 1e-10)
        tot = time.perf_counter() - t0
        if rep == 0:
            continue                              # warm-up
        if best is None or tot < best["total_s"]:
            best = dict(pr.times)
            best = {k + "_s": v for k, v in best.items()}
            best["total_s"] = tot
    best.update(dofs=int(pr.ts.ncp), nnz_A=int(pr.Afe.nnz), nnz_M=int(pr.M.nnz),
                nnz_C=int(pr.C0.nnz), solver=method, cg_iterations=int(pr.iters),
                dofs_per_s=pr.ts.ncp / best["total_s"],
                dofs_per_s_div8_ideal=8 * pr.ts.ncp / best["total_s"])
    return best


def b0(deg, nel):
    kv = [OB.uniform_knots(p, 0.0, 1.0, n) for p, n in zip(deg, nel)]
    ts = OB.TensorSpline(deg, kv)
    t0 = time.perf_counter()
    M = OX.build_M_loop(ts)
    t_loop = time.perf_counter() - t0
    t0 = time.perf_counter()
    Mk = OX.build_M_kron(ts)
    t_kron = time.perf_counter() - t0
    assert abs(M - Mk).max() == 0.0
    return {"fe_nodes": int(M.shape[0]), "dofs": int(M.shape[1]), "nnz_M": int(M.nnz),
            "loop_s": t_loop, "us_per_fe_node": 1e6 * t_loop / M.shape[0],
            "kron_s": t_kron, "note": "M built twice by the reference (control + fields)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--big", action="store_true", help="also 3-D cubic 48^3 (needs ~25 GB)")
    a = ap.parse_args()
    out = {"host": host(), "B0_faithful_M": {}, "B1_strong_cpu": {}}
    out["B0_faithful_M"]["cfg0 2-D p=2 64x64"] = b0([2, 2], [64, 64])
    out["B0_faithful_M"]["3-D p=3 12^3"] = b0([3, 3, 3], [12, 12, 12])
    out["B1_strong_cpu"]["cfg0 2-D p=2 64x64 (LU)"] = b1([2, 2], [64, 64], "poisson", "lu", a.reps)
    out["B1_strong_cpu"]["cfg2-like 2-D p=4 biharmonic 128x128 (LU)"] = b1(
        [4, 4], [128, 128], "biharmonic", "lu", a.reps, -1.0, 1.0, 2)
    for n in (16, 24, 32) + ((48,) if a.big else ()):
        out["B1_strong_cpu"]["3-D p=3 %d^3 (Jacobi-CG 1e-10)" % n] = b1(
            [3, 3, 3], [n] * 3, "poisson", "cg", a.reps if n <= 24 else 1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
