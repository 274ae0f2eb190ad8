"""
Biharmonic problem on a C^3 quartic explicit B-spline patch (the flow of the
reference's demos/biharmonic/biharmonic.py): residual form split with lhs/rhs,
two layers of clamped control points, energy-norm error and rate.
Usage: python examples/biharmonic.py [p] [base_nel] [levels]
"""
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tIGAr import *                 # noqa: E402,F401,F403
from tIGAr.BSplines import *        # noqa: E402,F401,F403

p = int(sys.argv[1]) if len(sys.argv) > 1 else 4
base = int(sys.argv[2]) if len(sys.argv) > 2 else 10
levels = int(sys.argv[3]) if len(sys.argv) > 3 else 3
errs = zeros(levels)
for level in range(levels):
    nel = base * 2 ** level
    t0 = time.perf_counter()
    mesh = ExplicitBSplineControlMesh([p, p], [uniformKnots(p, -1.0, 1.0, nel)] * 2)
    gen = EqualOrderSpline(1, mesh)
    scalar = gen.getScalarSpline(0)
    for direction in (0, 1):
        for side in (0, 1):
            gen.addZeroDofs(0, scalar.getSideDofs(direction, side, nLayers=2))
    spline = ExtractedSpline(gen, 2 * p)

    def lap(w):
        return spline.div(spline.grad(w))
    u, v = TrialFunction(spline.V), TestFunction(spline.V)
    x = spline.spatialCoordinates()
    soln = (cos(pi * x[0]) + 1.0) * (cos(pi * x[1]) + 1.0)
    f = lap(lap(soln))
    res = inner(lap(u), lap(v)) * spline.dx - inner(f, v) * spline.dx
    uh = Function(spline.V)
    spline.solveLinearVariationalProblem(res, uh)
    errs[level] = math.sqrt(assemble((lap(uh - soln) ** 2) * spline.dx))
    rate = "--" if level == 0 else "%.3f" % (math.log(errs[level - 1] / errs[level]) / math.log(2.0))
    if mpirank == 0:
        print("level %d: %dx%d elements, mode %s, CG its %d, energy error %.6e (rate %s), %.2f s"
              % (level, nel, nel, spline.mode, spline.lastSolve["iterations"], errs[level], rate,
                 time.perf_counter() - t0))
