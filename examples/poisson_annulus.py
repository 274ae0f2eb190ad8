"""
Poisson on a cubic NURBS quarter annulus (BASELINE configs[3] geometry; the
flow of the reference's demos/poisson/poisson-nurbs.py with rationalised trial
and test functions).  Usage: python examples/poisson_annulus.py [dim] [nel] [levels]
"""
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tIGAr import *                 # noqa: E402,F401,F403
from tIGAr.NURBS import *           # noqa: E402,F401,F403

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 2
base = int(sys.argv[2]) if len(sys.argv) > 2 else 8
levels = int(sys.argv[3]) if len(sys.argv) > 3 else 3
errs = zeros(levels)
for level in range(levels):
    nel = base * 2 ** level
    t0 = time.perf_counter()
    gen = EqualOrderSpline(1, NURBSControlMesh(quarter_annulus(3, nel, dim)))
    scalar = gen.getScalarSpline(0)
    for direction in range(dim):
        for side in (0, 1):
            gen.addZeroDofs(0, scalar.getSideDofs(direction, side))
    spline = ExtractedSpline(gen, 6)
    u = spline.rationalize(TrialFunction(spline.V))
    v = spline.rationalize(TestFunction(spline.V))
    x = spline.spatialCoordinates()
    r = sqrt(x[0] * x[0] + x[1] * x[1])
    soln = (r - 1.0) * (2.0 - r) * (2.0 * x[0] * x[1] / (r * r))
    if dim == 3:
        soln = soln * sin(pi * x[2])
    f = -spline.div(spline.grad(soln))
    uh = Function(spline.V)
    spline.solveLinearVariationalProblem(
        inner(spline.grad(u), spline.grad(v)) * spline.dx == inner(f, v) * spline.dx, uh)
    errs[level] = math.sqrt(assemble(((spline.rationalize(uh) - soln) ** 2) * spline.dx))
    rate = "--" if level == 0 else "%.3f" % (math.log(errs[level - 1] / errs[level]) / math.log(2.0))
    if mpirank == 0:
        print("level %d: %d^%d elements, %d DoFs, mode %s, CG its %d, L2 error %.6e (rate %s), %.2f s"
              % (level, nel, dim, spline.patch().n_iga, spline.mode, spline.lastSolve["iterations"],
                 errs[level], rate, time.perf_counter() - t0))
