"""Worker launched by torchrun from tests/test_gpu_multi.py (one rank per GPU):
distributed fused assembly + row-distributed CG against the oracle."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gpu_util import make_pair, uk, rel
    from tIGAr import TrialFunction, TestFunction, Function, inner, sin, pi, assemble, mpirank
    PI = math.pi
    cases = [(s, d, n) for s in ("fd", "jacobi")
             for d, n in [([3, 3, 3], [6, 5, 9]), ([2, 2, 2], [5, 5, 4]), ([2, 2], [9, 11])]]
    for solver, deg, nels in cases:
        os.environ["TIGAR_B200_SOLVER"] = solver      # row-distributed FD-CG / Jacobi-CG
        kv = [uk(p, n) for p, n in zip(deg, nels)]
        gen, spline, pr = make_pair(deg, kv, mode=None)
        assert spline.mode == "fused" and spline.patch().part is not None
        u, v = TrialFunction(spline.V), TestFunction(spline.V)
        x = spline.spatialCoordinates()
        f = len(deg) * PI ** 2
        soln = 1.0
        for d in range(len(deg)):
            soln = soln * sin(PI * x[d])
        uh = Function(spline.V)
        U = spline.solveLinearVariationalProblem(
            inner(spline.grad(u), spline.grad(v)) * spline.dx == inner(f * soln, v) * spline.dx, uh)
        Uo = pr.run(lambda X: len(deg) * PI ** 2 * np.prod(np.sin(PI * X[..., :len(deg)]), axis=-1))
        err = rel(U.get_local(), Uo)
        l2 = math.sqrt(assemble(((uh - soln) ** 2) * spline.dx))
        l2o = pr.error(Uo, "l2", lambda X: np.prod(np.sin(PI * X[..., :len(deg)]), axis=-1))
        ok = err < 1e-10 and abs(l2 - l2o) / l2o < 1e-6
        ok = ok and spline.lastSolve["method"] == solver
        if solver == "fd":                            # exact preconditioner on this geometry
            ok = ok and spline.lastSolve["iterations"] <= 2
        print("rank %d %s deg %s: rel diff vs oracle %.2e, its %d, L2 %.6e vs %.6e -> %s"
              % (dist.get_rank(), solver, deg, err, spline.lastSolve["iterations"], l2, l2o,
                 "OK" if ok else "FAIL"), flush=True)
        if not ok:
            sys.exit(1)
    # curved rational geometry (configs[3] in small): FD-CG across ranks against the oracle
    os.environ["TIGAR_B200_SOLVER"] = "fd"
    from test_gpu_nurbs import build, forms
    spline, pr = build(3, [5, 6, 7], 3, None)
    a, L = forms(spline, 3)
    uh = Function(spline.V)
    U = spline.solveLinearVariationalProblem(a == L, uh)
    Uo = pr.run(lambda X: np.sin(PI * X[..., 0]) * X[..., 1] + 1.0)
    err = rel(U.get_local(), Uo)
    print("rank %d annulus: rel diff vs oracle %.2e, FD-CG its %d" % (
        dist.get_rank(), err, spline.lastSolve["iterations"]), flush=True)
    if not (err < 1e-10 and spline.lastSolve["iterations"] < 60):
        sys.exit(1)
    dist.barrier()
    dist.destroy_process_group()
    print("MGPU_PARITY_OK", flush=True)


if __name__ == "__main__":
    main()
