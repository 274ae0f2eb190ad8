"""
Poisson on an explicit B-spline patch (the flow of the reference's
demos/poisson/poisson.py, written against the same tIGAr API): manufactured
solution sin(pi x) sin(pi y), three refinement levels, L2 error and rate.
Usage: python examples/poisson_explicit.py [p] [base_nel] [levels]
"""
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tIGAr import *                 # noqa: E402,F401,F403
from tIGAr.BSplines import *        # noqa: E402,F401,F403

p = int(sys.argv[1]) if len(sys.argv) > 1 else 3
base = int(sys.argv[2]) if len(sys.argv) > 2 else 10
levels = int(sys.argv[3]) if len(sys.argv) > 3 else 3
errs = zeros(levels)
for level in range(levels):
    nel = base * 2 ** level
    t0 = time.perf_counter()
    mesh = ExplicitBSplineControlMesh([p, p], [uniformKnots(p, 0.0, 1.0, nel)] * 2)
    gen = EqualOrderSpline(1, mesh)
    scalar = gen.getScalarSpline(0)
    for direction in (0, 1):
        for side in (0, 1):
            gen.addZeroDofs(0, scalar.getSideDofs(direction, side))
    gen.writeExtraction("./extraction")
    spline = ExtractedSpline(gen, 2 * p)
    u, v = TrialFunction(spline.V), TestFunction(spline.V)
    x = spline.spatialCoordinates()
    soln = sin(pi * x[0]) * sin(pi * x[1])
    f = -spline.div(spline.grad(soln))
    uh = Function(spline.V)
    spline.solveLinearVariationalProblem(
        inner(spline.grad(u), spline.grad(v)) * spline.dx == inner(f, v) * spline.dx, uh)
    errs[level] = math.sqrt(assemble(((uh - soln) ** 2) * spline.dx))
    rate = "--" if level == 0 else "%.3f" % (math.log(errs[level - 1] / errs[level]) / math.log(2.0))
    if mpirank == 0:
        print("level %d: %dx%d elements, mode %s, CG its %d, L2 error %.6e (rate %s), %.2f s"
              % (level, nel, nel, spline.mode, spline.lastSolve["iterations"], errs[level], rate,
                 time.perf_counter() - t0))
File("results/u.pvd") << uh
