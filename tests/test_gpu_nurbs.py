"""GPU parity on rational (NURBS) geometry -- BASELINE configs[3] family: cubic
quarter annulus, rationalised trial/test functions as in poisson-nurbs.py:121-133."""
import math

import numpy as np
import pytest

from gpu_util import rel, relm
from oracle import bsplines as OB
from oracle import pipeline as OP

pytestmark = pytest.mark.gpu
PI = math.pi


def build(p, nel, dim, mode):
    from tIGAr import EqualOrderSpline, ExtractedSpline
    from tIGAr.NURBS import NURBSControlMesh, quarter_annulus
    nrb = quarter_annulus(p, nel, dim)
    cm = NURBSControlMesh(nrb)
    gen = EqualOrderSpline(1, cm)
    sp = gen.getScalarSpline(0)
    for d in range(dim):
        for side in (0, 1):
            gen.addZeroDofs(0, sp.getSideDofs(d, side))
    spline = ExtractedSpline(gen, 2 * p, mode=mode)
    kv = [list(k) for k in nrb.knots]
    pr = OP.Problem(list(nrb.degree), kv, P=cm.bnet, rationalize=True, quadDeg=2 * p)
    return spline, pr


def forms(spline, dim):
    from tIGAr import TrialFunction, TestFunction, inner, sin
    u = spline.rationalize(TrialFunction(spline.V))
    v = spline.rationalize(TestFunction(spline.V))
    x = spline.spatialCoordinates()
    f = sin(PI * x[0]) * x[1] + 1.0
    return inner(spline.grad(u), spline.grad(v)) * spline.dx, inner(f, v) * spline.dx


@pytest.mark.parametrize("mode", ["csr", "fused"])
@pytest.mark.parametrize("dim,nel", [(2, [5, 6]), (3, [3, 4, 2])])
def test_annulus_system_and_solution_match_oracle(dim, nel, mode):
    from tIGAr import Function
    spline, pr = build(3, nel, dim, mode)
    a, L = forms(spline, dim)
    f = lambda X: np.sin(PI * X[..., 0]) * X[..., 1] + 1.0
    Uo = pr.run(f)
    C = spline.assembleMatrix(a)
    assert relm(C.to_scipy(), pr.C) < 1e-11
    b = spline.assembleVector(L)
    assert rel(b.get_local(), pr.b) < 1e-11
    uh = Function(spline.V)
    U = spline.solveLinearVariationalProblem(a == L, uh)
    assert rel(U.get_local(), Uo) < 1e-10


def test_annulus_manufactured_solution_converges():
    """u = (r-1)(2-r) sin(2 theta) vanishes on the whole boundary; f = -div grad u
    built symbolically as the demo does (poisson-nurbs.py:127-133).  Cubic NURBS:
    L2 rate ~ 4."""
    from tIGAr import TrialFunction, TestFunction, Function, inner, sqrt, assemble
    errs = []
    for nel in (4, 8):
        spline, _ = build(3, [nel, nel], 2, "fused")
        u = spline.rationalize(TrialFunction(spline.V))
        v = spline.rationalize(TestFunction(spline.V))
        x = spline.spatialCoordinates()
        r = sqrt(x[0] * x[0] + x[1] * x[1])
        soln = (r - 1.0) * (2.0 - r) * (2.0 * x[0] * x[1] / (r * r))       # sin(2 theta)
        f = -spline.div(spline.grad(soln))
        uh = Function(spline.V)
        spline.solveLinearVariationalProblem(
            inner(spline.grad(u), spline.grad(v)) * spline.dx == inner(f, v) * spline.dx, uh)
        errs.append(math.sqrt(assemble(((spline.rationalize(uh) - soln) ** 2) * spline.dx)))
    rate = math.log(errs[0] / errs[1]) / math.log(2.0)
    assert errs[1] < 1e-4 and 3.5 < rate < 4.8, (errs, rate)
