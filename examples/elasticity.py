"""
Plane-strain linear elasticity with two displacement fields on an explicit B-spline
patch -- the equal-order multi-field path (EqualOrderSpline(2, ...): MixedElement of the
reference, common.py:337-351): manufactured solution, L2 error and rate.
Usage: python examples/elasticity.py [p] [base_nel] [levels]
"""
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tIGAr import *                 # noqa: E402,F401,F403
from tIGAr.BSplines import *        # noqa: E402,F401,F403

p = int(sys.argv[1]) if len(sys.argv) > 1 else 3
base = int(sys.argv[2]) if len(sys.argv) > 2 else 8
levels = int(sys.argv[3]) if len(sys.argv) > 3 else 3
MU, LAM = 1.0, 2.0
errs = zeros(levels)
for level in range(levels):
    nel = base * 2 ** level
    t0 = time.perf_counter()
    mesh = ExplicitBSplineControlMesh([p, p], [uniformKnots(p, 0.0, 1.0, nel)] * 2)
    gen = EqualOrderSpline(2, mesh)
    scalar = gen.getScalarSpline(0)
    for field in (0, 1):
        for direction in (0, 1):
            for side in (0, 1):
                gen.addZeroDofs(field, scalar.getSideDofs(direction, side))
    spline = ExtractedSpline(gen, 2 * p)
    u, v = TrialFunction(spline.V), TestFunction(spline.V)
    x = spline.spatialCoordinates()
    bubble = sin(pi * x[0]) * sin(pi * x[1])
    soln = as_vector([bubble, 0.5 * bubble * x[0]])

    def eps(w):
        g = spline.grad(w)
        return 0.5 * (g + g.T)

    def sigma(w):
        e = eps(w)
        return 2.0 * MU * e + LAM * tr(e) * as_matrix([[1.0, 0.0], [0.0, 1.0]])
    f = -spline.div(sigma(soln))
    uh = Function(spline.V)
    ks = PETScKrylovSolver("cg", "jacobi")
    ks.parameters["relative_tolerance"] = 1e-12
    spline.setSolverOptions(linearSolver=ks)
    spline.solveLinearVariationalProblem(
        inner(sigma(u), eps(v)) * spline.dx == inner(f, v) * spline.dx, uh)
    errs[level] = math.sqrt(assemble(inner(uh - soln, uh - soln) * spline.dx))
    rate = "--" if level == 0 else "%.3f" % (math.log(errs[level - 1] / errs[level]) / math.log(2.0))
    if mpirank == 0:
        print("level %d: %dx%d elements, 2 fields, mode %s, CG its %d, L2 error %.6e (rate %s), "
              "%.2f s" % (level, nel, nel, spline.mode, spline.lastSolve["iterations"],
                          errs[level], rate, time.perf_counter() - t0))
