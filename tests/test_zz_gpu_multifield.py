"""GPU parity of the equal-order multi-field path: vector-valued linear
elasticity on a curved rational patch through the tIGAr API (EqualOrderSpline with
nFields = nsd -> per-field-block assembly on the scalar kernels -> block BCs -> block
Jacobi-CG), against the oracle's independently assembled block system.

Tolerances as for the scalar path: matrix / vector 1e-12 relative, IGA DoF vector 1e-10
relative (BASELINE.json north_star).  (File name sorts last on purpose: this widening was
written after the round's GPU budget was spent)"""
import math
import os

import numpy as np
import pytest

from gpu_util import rel, relm
from oracle import bsplines as OB
from oracle import pipeline as OP

pytestmark = pytest.mark.gpu

MU, LAM = 0.7, 1.9


def patch(deg, nel, amp=0.08):
    """Explicit B-spline patch with a perturbed control net and non-constant weights."""
    dim = len(deg)
    kv = [OB.uniform_knots(p, 0.0, 1.0, n) for p, n in zip(deg, nel)]
    ts = OB.TensorSpline(deg, kv)
    P = OB.explicit_control_net(ts, 0).copy()
    X = P[:, :dim].copy()
    P[:, 0] = X[:, 0] + amp * np.prod(np.sin(math.pi * X), axis=1)
    P[:, 1] = X[:, 1] + amp * X[:, 0] * (1 - X[:, 0]) * np.cos(1.3 * X[:, 1])
    w = 1.0 + 0.2 * X[:, 0] * X[:, 1]
    P[:, :dim] *= w[:, None]
    P[:, dim] = w
    return ts, kv, P


def force_np(dim):
    def f(X):
        comps = [np.sin(2.0 * X[..., 0]) * X[..., 1], 0.5 + X[..., 0] * X[..., 1] ** 2]
        if dim == 3:
            comps.append(X[..., 2] - 0.3 * X[..., 0])
        return np.stack(comps, -1)
    return f


def build(deg, nel, mode):
    from tIGAr import EqualOrderSpline, ExtractedSpline
    from tIGAr.BSplines import ExplicitBSplineControlMesh
    dim = len(deg)
    ts, kv, P = patch(deg, nel)
    gen = EqualOrderSpline(dim, ExplicitBSplineControlMesh(deg, kv))
    sp = gen.getScalarSpline(0)
    for f in range(dim):                                   # clamp side 0 of direction 0
        gen.addZeroDofs(f, sp.getSideDofs(0, 0))
    gen.addZeroDofs(1, sp.getSideDofs(1, 1))                 # rollers for field 1 elsewhere
    spline = ExtractedSpline(gen, 2 * max(deg), mode=mode, controlNet=P)
    prob = OP.ElasticityProblem(deg, kv, P, MU, LAM, gen.zeroDofs)
    return spline, prob, ts.ncp


def forms(spline, u, v):
    from tIGAr import inner, sin, as_vector

    def eps(w):
        g = spline.grad(w)
        return 0.5 * (g + g.T)
    x = spline.spatialCoordinates()
    comps = [sin(2.0 * x[0]) * x[1], 0.5 + x[0] * x[1] ** 2]
    if len(x) == 3:
        comps.append(x[2] - 0.3 * x[0])
    a = (2.0 * MU * inner(eps(u), eps(v)) + LAM * spline.div(u) * spline.div(v)) * spline.dx
    L = inner(as_vector(comps), v) * spline.dx
    return a, L


@pytest.mark.parametrize("mode", ["fused", "csr"])
def test_elasticity_2d_matches_oracle(mode):
    from tIGAr import TrialFunction, TestFunction, Function, KrylovSolver
    deg, nel = [3, 3], [7, 6]
    spline, prob, n = build(deg, nel, mode)
    Uo = prob.solve(force_np(2))
    u, v = TrialFunction(spline.V), TestFunction(spline.V)
    a, L = forms(spline, u, v)
    MTAM, MTb = spline.assembleLinearSystem(a, L)
    assert relm(MTAM.to_scipy(), prob.C) < 1e-12
    assert rel(MTb.get_local(), prob.b) < 1e-12
    ks = KrylovSolver("cg", "jacobi")
    ks.parameters["relative_tolerance"] = 1e-13
    spline.setSolverOptions(linearSolver=ks)
    uh = Function(spline.V)
    U = spline.solveLinearSystem(MTAM, MTb, uh)
    assert rel(U.get_local(), Uo) < 1e-10
    assert rel(uh.comps[1].iga.cpu().numpy(), Uo[n:]) < 1e-10
    # no linearSolver = the reference's direct LU (common.py:1255-1256): band Cholesky on the
    # node-major interleaved block system
    spline.setSolverOptions()
    ud = Function(spline.V)
    Ud = spline.solveLinearSystem(MTAM, MTb, ud)
    assert spline.lastSolve["method"] == "direct"
    assert rel(Ud.get_local(), Uo) < 1e-11


def test_elasticity_3d_three_fields_fused_and_newton():
    """Three fields on a 3-D cubic patch (the sum-factorised 3-D kernels per block), solved
    once as a linear problem and once by Newton with J = derivative(R, u)."""
    from tIGAr import TrialFunction, TestFunction, Function, KrylovSolver, derivative
    deg, nel = [3, 3, 3], [3, 2, 3]
    spline, prob, n = build(deg, nel, "fused")
    Uo = prob.solve(force_np(3))
    u, v = TrialFunction(spline.V), TestFunction(spline.V)
    a, L = forms(spline, u, v)
    ks = KrylovSolver("cg", "jacobi")
    ks.parameters["relative_tolerance"] = 1e-13
    spline.setSolverOptions(maxIters=4, relativeTolerance=1e-9, linearSolver=ks)
    uh = Function(spline.V)
    U = spline.solveLinearVariationalProblem(a == L, uh)
    assert rel(U.get_local(), Uo) < 1e-10
    un = Function(spline.V)
    R = forms(spline, un, v)[0] - L
    spline.solveNonlinearVariationalProblem(R, derivative(R, un), un)
    assert rel(un.iga.cpu().numpy(), Uo) < 1e-8


@pytest.mark.parametrize("solver", ["jacobi", "direct"])
def test_kl_shell_scordelis_lo_roof_on_the_device(solver):
    """BASELINE configs[4] in small: the SVK Kirchhoff-Love shell residual of the
    reference's kl-shell-svk demo (tests/test_kl_shell_cpu.shell_forms, CPU-checked) on the
    cubic NURBS Scordelis-Lo roof, three fields, Newton with J = derivative(R, y) through
    ExtractedSpline.solveNonlinearVariationalProblem.  Checks the device tangent against
    the host-integrated one and the textbook mid-side displacement (0.3006)."""
    from tIGAr import (EqualOrderSpline, ExtractedSpline, Function, TestFunction, KrylovSolver)
    from tIGAr.NURBS import NURBSControlMesh, cylindrical_roof
    from test_kl_shell_cpu import shell_forms, Roof
    from oracle import assembly as OA
    scale, nel = 1e-3, [6, 6]
    host = Roof(nel, -90.0 * scale)
    gen = EqualOrderSpline(3, NURBSControlMesh(cylindrical_roof(3, nel)))
    sp = gen.getScalarSpline(0)
    for side in (0, 1):                                   # rigid diaphragms: u_x = u_z = 0
        gen.addZeroDofs(0, sp.getSideDofs(1, side))
        gen.addZeroDofs(2, sp.getSideDofs(1, side))
    n0 = sp.splines[0].ncp
    gen.addZeroDofs(1, [n0 // 2])                         # axial rigid-body translation
    spline = ExtractedSpline(gen, 6, mode="fused")
    n = host.n
    y = Function(spline.V)
    z = TestFunction(spline.V)
    W, res, dres = shell_forms(spline, y, z, -90.0 * scale)
    K = spline.assembleMatrix(dres, applyBCs=False)
    Kh = host.tangent(np.zeros(3 * n))
    assert relm(K.to_scipy(), Kh) < 1e-10
    ks = KrylovSolver("cg", "jacobi")
    ks.parameters["relative_tolerance"] = 1e-11
    spline.setSolverOptions(maxIters=6, relativeTolerance=1e-6,
                            linearSolver=ks if solver == "jacobi" else None)
    spline.solveNonlinearVariationalProblem(res, dres, y)
    assert spline.lastSolve["method"] == solver
    Uv = y.iga.cpu().numpy()
    s1 = host.ts.splines[1]
    span = int(s1.getKnotSpan(0.5))
    N = OA.bspline_ders(s1.ghostKnots, s1.p, span + s1.nGhost, 0.5, 0)[0]
    idx = (span - s1.p + np.arange(s1.p + 1)) * n0
    uz = (N * Uv[2 * n + idx]).sum() / (N * host.P[idx, 3]).sum()
    assert abs(abs(uz) / scale - 0.3006) < 0.02 * 0.3006


def test_fe_to_iga_round_trip_on_the_device():
    """SURVEY 8c KAT 5: FEtoIGA(M U) = U (common.py:968-993)."""
    from tIGAr import Function
    from tigar_b200 import dev
    spline, prob, n = build([2, 2], [6, 5], "csr")
    rng = np.random.RandomState(2)
    Uv = rng.rand(2 * n)
    w = Function(spline.V)
    w.set_iga(dev.from_np(Uv))
    assert np.abs(spline.FEtoIGA(w).get_local() - Uv).max() < 1e-9


@pytest.mark.parametrize("deg,nels", [([3, 3, 3], [5, 4, 9]), ([2, 2], [9, 8])])
def test_matrix_free_mode_on_the_device(deg, nels):
    """mode="matfree" (tigar_b200/matfree.py, SURVEY 7.2 hard part 1): operator action =
    assembled matrix times vector, slab-wise Jacobi diagonal = diagonal of the matrix,
    solve = oracle LU solution (1e-10)."""
    import torch
    from tIGAr import TrialFunction, TestFunction, Function, KrylovSolver, inner, sin
    from tigar_b200 import dev
    from gpu_util import make_pair
    kv = [OB.uniform_knots(p, 0.0, 1.0, n) for p, n in zip(deg, nels)]
    gen, fused, pr = make_pair(deg, kv, mode="fused")
    gen2, mf, _ = make_pair(deg, kv, mode="matfree")

    def forms(spline):
        u, v = TrialFunction(spline.V), TestFunction(spline.V)
        x = spline.spatialCoordinates()
        f = 1.0
        for d in range(len(x)):
            f = f * sin(math.pi * x[d])
        return (inner(spline.grad(u), spline.grad(v)) * spline.dx,
                inner(f * (len(x) * math.pi ** 2), v) * spline.dx)
    a, L = forms(fused)
    C = fused.assembleMatrix(a, diag=2.5)
    a2, L2 = forms(mf)
    op = mf.assembleMatrix(a2, diag=2.5)
    n = fused._patch.n_iga
    xv = dev.from_np(np.random.RandomState(0).rand(n))
    yref = dev.to_np(C.matvec(xv))
    # default: the one-kernel variant (jit.generate(..., op=...))
    assert rel(dev.to_np(op.matvec(xv)), yref) < 1e-12
    # two-kernel variant (Gauss-point pass + vector assembly), same operator
    os.environ["TIGAR_B200_MF_FUSED"] = "0"
    try:
        y2 = dev.to_np(op.matvec(xv))
    finally:
        del os.environ["TIGAR_B200_MF_FUSED"]
    assert rel(y2, yref) < 1e-12
    C0 = fused.assembleMatrix(a, applyBCs=False).to_scipy()
    op0 = mf.assembleMatrix(a2, applyBCs=False)
    d1 = dev.to_np(op0.jacobi_dinv(1))                    # explicit slab count: assembled
    assert rel(d1, 1.0 / C0.diagonal()) < 1e-13
    op0._dinv = None
    d2 = dev.to_np(op0.jacobi_dinv(2))
    assert rel(d2, d1) < 1e-13
    op0._dinv = None
    d3 = dev.to_np(op0.jacobi_dinv())                     # default: generated diagonal kernel
    assert rel(d3, d1) < 1e-12
    # assembleLinearSystem: load vector and diagonal from ONE Gauss-point pass (the march's
    # second vector job on product tables, engine.assemble_vector_and_diag)
    opS, bS = mf.assembleLinearSystem(a2, L2, applyBCs=False)
    assert opS._dinv is not None, "the shared pass did not run"
    assert rel(dev.to_np(opS._dinv), d1) < 1e-12
    assert rel(bS.get_local(), mf.assembleVector(L2, applyBCs=False).get_local()) < 1e-13
    ks = KrylovSolver("cg", "jacobi")
    ks.parameters["relative_tolerance"] = 1e-13
    mf.setSolverOptions(linearSolver=ks)
    uh = Function(mf.V)
    U = mf.solveLinearVariationalProblem(a2 == L2, uh)
    Uo = pr.run(lambda X: len(deg) * math.pi ** 2 * np.prod(np.sin(math.pi * X[..., :len(deg)]), axis=-1))
    assert rel(U.get_local(), Uo) < 1e-10
