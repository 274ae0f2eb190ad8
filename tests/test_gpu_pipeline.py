"""GPU parity, layers (ii)+(iii): Gauss-point assembly, M^T A M / M^T b, BCs
and the solve, against the oracle on the same inputs.

Tolerances (floating point; stated per north_star): matrices/vectors agree to
1e-12 relative (different summation order), solutions to 1e-10 relative in the
IGA DoF vector (BASELINE.json: ||u_new - u_ref|| / ||u_ref|| < 1e-10)."""
import math

import numpy as np
import pytest

from gpu_util import make_pair, uk, rel, relm

pytestmark = pytest.mark.gpu

PI = math.pi


def poisson_forms(spline, kind="sin"):
    from tIGAr import TrialFunction, TestFunction, inner, sin
    u, v = TrialFunction(spline.V), TestFunction(spline.V)
    x = spline.spatialCoordinates()
    if kind == "sin":
        f = 1.0
        for d in range(len(x)):
            f = f * sin(PI * x[d])
        f = f * (len(x) * PI ** 2)
    else:
        f = x[0] * x[0] + 1.0
    a = inner(spline.grad(u), spline.grad(v)) * spline.dx
    L = inner(f, v) * spline.dx
    return a, L


def f_np(kind, dim):
    if kind == "sin":
        return lambda X: dim * PI ** 2 * np.prod(np.sin(PI * X[..., :dim]), axis=-1)
    return lambda X: X[..., 0] ** 2 + 1.0


CASES = [([2, 2], [6, 5]), ([3, 3], [4, 5]), ([2, 3], [4, 4]), ([2, 2, 2], [3, 2, 3]),
         ([3, 3, 3], [2, 3, 2]), ([2], [7])]


@pytest.mark.parametrize("deg,nels", CASES)
def test_fe_assembly_and_ptap_match_oracle(deg, nels):
    from tigar_b200 import dev
    kv = [uk(p, n) for p, n in zip(deg, nels)]
    gen, spline, pr = make_pair(deg, kv, mode="csr")
    a, L = poisson_forms(spline, "poly")
    pr.extract()
    pr.assemble(f_np("poly", len(deg)))
    pr.ptap()
    # cpFuncs = M_control * P  (common.py:367-380)
    for i in range(spline.nsd + 1):
        assert np.abs(spline.cpFuncs[i].vector().get_local() - pr.cpn[:, i]).max() < 1e-14
    A = spline._assemble_kind(a, "fe")
    assert relm(A.to_scipy(), pr.Afe) < 1e-12
    b = spline._assemble_kind(L, "fe")
    assert rel(dev.to_np(b), pr.bfe) < 1e-12
    C0 = spline.extractMatrix(A, applyBCs=False)
    assert relm(C0.to_scipy(), pr.C0) < 1e-12
    MTb0 = spline.extractVector(b, applyBCs=False)
    assert rel(MTb0.get_local(), pr.b0) < 1e-12
    C = spline.assembleMatrix(a, diag=3.5)
    pr.ptap(diag=3.5)
    assert relm(C.to_scipy(), pr.C) < 1e-12
    Cd = C.to_scipy().toarray()
    z = np.unique(pr.zeroDofs)
    assert np.all(Cd[z, z] == 3.5)
    MTb = spline.assembleVector(L)
    assert rel(MTb.get_local(), pr.b) < 1e-12
    assert not MTb.get_local()[z].any()


@pytest.mark.parametrize("deg,nels", CASES)
def test_fused_equals_csr_path(deg, nels):
    kv = [uk(p, n) for p, n in zip(deg, nels)]
    gen, spline, pr = make_pair(deg, kv, mode="fused")
    a, L = poisson_forms(spline, "sin")
    pr.extract()
    pr.assemble(f_np("sin", len(deg)))
    pr.ptap()
    C = spline.assembleMatrix(a)
    assert relm(C.to_scipy(), pr.C) < 1e-12
    MTb = spline.assembleVector(L)
    assert rel(MTb.get_local(), pr.b) < 1e-12


@pytest.mark.parametrize("mode", ["csr", "fused"])
@pytest.mark.parametrize("deg,nels", [([2, 2], [16, 16]), ([3, 3], [10, 10]), ([3, 3, 3], [5, 4, 5]),
                                      ([2, 2, 2], [8, 8, 8])])
def test_poisson_solution_matches_oracle(deg, nels, mode):
    from tIGAr import Function, assemble
    kv = [uk(p, n) for p, n in zip(deg, nels)]
    gen, spline, pr = make_pair(deg, kv, mode=mode)
    a, L = poisson_forms(spline, "sin")
    u = Function(spline.V)
    MTU = spline.solveLinearVariationalProblem(a == L, u)
    Uo = pr.run(f_np("sin", len(deg)))             # oracle: sparse LU
    assert rel(MTU.get_local(), Uo) < 1e-10        # the north_star tolerance
    # FE representation u = M U (common.py:1259)
    assert rel(u.vector().get_local(), pr.M @ Uo) < 1e-10
    # L2 error functional (poisson.py:132)
    x = spline.spatialCoordinates()
    from tIGAr import sin
    soln = 1.0
    for d in range(len(deg)):
        soln = soln * sin(PI * x[d])
    err = math.sqrt(assemble(((u - soln) ** 2) * spline.dx))
    exact = lambda X: np.prod(np.sin(PI * X[..., :len(deg)]), axis=-1)
    erro = pr.error(Uo, "l2", exact)
    assert abs(err - erro) / erro < 1e-7


def test_biharmonic_matches_oracle():
    from tIGAr import TrialFunction, TestFunction, Function, inner, cos, assemble
    deg, nels = [4, 4], [8, 8]
    kv = [uk(p, n, -1.0, 1.0) for p, n in zip(deg, nels)]
    for mode in ("csr", "fused"):
        gen, spline, pr = make_pair(deg, kv, nLayers=2, form="biharmonic", mode=mode)
        u, v = TrialFunction(spline.V), TestFunction(spline.V)
        lap = lambda w: spline.div(spline.grad(w))
        x = spline.spatialCoordinates()
        cx, cy = cos(PI * x[0]), cos(PI * x[1])
        f = PI ** 4 * (cx * (cy + 1.0) + 2.0 * cx * cy + (cx + 1.0) * cy)
        res = inner(lap(u), lap(v)) * spline.dx - inner(f, v) * spline.dx
        uh = Function(spline.V)
        MTU = spline.solveLinearVariationalProblem(res, uh)
        fo = lambda X: PI ** 4 * (np.cos(PI * X[..., 0]) * (np.cos(PI * X[..., 1]) + 1)
                                  + 2 * np.cos(PI * X[..., 0]) * np.cos(PI * X[..., 1])
                                  + (np.cos(PI * X[..., 0]) + 1) * np.cos(PI * X[..., 1]))
        Uo = pr.run(fo)
        assert relm(spline.assembleMatrix(inner(lap(u), lap(v)) * spline.dx).to_scipy(), pr.C) < 1e-11
        assert rel(MTU.get_local(), Uo) < 1e-8      # cond ~ h^-4: looser than Poisson
        soln = (cx + 1.0) * (cy + 1.0)
        en = math.sqrt(assemble((lap(uh - soln) ** 2) * spline.dx))
        lapo = lambda X: -PI ** 2 * (np.cos(PI * X[..., 0]) * (np.cos(PI * X[..., 1]) + 1)
                                     + (np.cos(PI * X[..., 0]) + 1) * np.cos(PI * X[..., 1]))
        eo = pr.error(Uo, "energy", lapo)
        assert abs(en - eo) / eo < 1e-5


def test_manufactured_force_from_div_grad():
    """f = -div(grad(soln)) as the demo writes it (poisson.py:112-114) equals
    the closed form."""
    from tIGAr import TestFunction, inner, sin
    deg, nels = [3, 3], [6, 6]
    kv = [uk(p, n) for p, n in zip(deg, nels)]
    gen, spline, pr = make_pair(deg, kv, mode="csr")
    v = TestFunction(spline.V)
    x = spline.spatialCoordinates()
    soln = sin(PI * x[0]) * sin(PI * x[1])
    f = -spline.div(spline.grad(soln))
    b1 = spline.assembleVector(inner(f, v) * spline.dx, applyBCs=False).get_local()
    b2 = spline.assembleVector(inner(2 * PI ** 2 * soln, v) * spline.dx, applyBCs=False).get_local()
    assert rel(b1, b2) < 1e-9


def test_generic_csr_kernels_match_windowed():
    """tg_spmv / tg_zero_rows_cols / tg_solve_cg (general CSR) against the
    windowed variants and scipy."""
    import ctypes as C
    from tigar_b200 import dev
    from tigar_b200._lib import lib, check
    deg, nels = [2, 2], [7, 6]
    kv = [uk(p, n) for p, n in zip(deg, nels)]
    gen, spline, pr = make_pair(deg, kv, mode="csr")
    a, L = poisson_forms(spline, "sin")
    Cm = spline.assembleMatrix(a)
    b = spline.assembleVector(L)
    w = Cm.window
    rp, cols = w.rowptr(), w.columns()
    cv = Cm.csr_values()              # exact CSR order (the window may be SELL)
    n = w.nrows
    rng = np.random.RandomState(1)
    x = dev.from_np(rng.rand(n))
    y1, y2 = dev.empty(n), dev.empty(n)
    check(lib.tg_spmv(dev.ptr(rp), dev.ptr(cols), dev.ptr(cv), dev.ptr(x), dev.ptr(y1), n,
                      dev.stream()))
    Cm.matvec(x, y2)
    assert np.abs(dev.to_np(y1) - dev.to_np(y2)).max() < 1e-12
    assert np.abs(dev.to_np(y1) - Cm.to_scipy() @ dev.to_np(x)).max() < 1e-12
    sol = dev.zeros(n)
    work = dev.empty(4 * n + lib.tg_cg_scratch_len() + 8)
    its, relres = C.c_int32(0), C.c_double(0)
    check(lib.tg_solve_cg(dev.ptr(rp), dev.ptr(cols), dev.ptr(cv), dev.ptr(b.t), dev.ptr(sol),
                          n, 1e-13, 0.0, 10000, 10, dev.ptr(work), C.byref(its), C.byref(relres),
                          dev.stream()))
    import scipy.sparse.linalg as spla
    ref = spla.spsolve(Cm.to_scipy().tocsc(), b.get_local())
    assert rel(dev.to_np(sol), ref) < 1e-10 and its.value > 0
    # general-CSR zeroRowsColumns
    A2 = spline.assembleMatrix(a, applyBCs=False).csr_values()
    check(lib.tg_zero_rows_cols(dev.ptr(rp), dev.ptr(cols), dev.ptr(A2), n,
                                dev.ptr(spline._bc_mask()), 1.0, dev.stream()))
    assert np.array_equal(dev.to_np(A2), dev.to_np(cv))


def test_write_and_read_extraction(tmp_path):
    from tIGAr import ExtractedSpline, Function
    deg, nels = [2, 2], [5, 5]
    kv = [uk(p, n) for p, n in zip(deg, nels)]
    gen, spline, pr = make_pair(deg, kv, mode="csr")
    d = str(tmp_path / "extraction")
    gen.writeExtraction(d)
    sp2 = ExtractedSpline(d, 4, mode="csr")
    a, L = poisson_forms(spline, "sin")
    a2, L2 = poisson_forms(sp2, "sin")
    u1, u2 = Function(spline.V), Function(sp2.V)
    U1 = spline.solveLinearVariationalProblem(a == L, u1).get_local()
    U2 = sp2.solveLinearVariationalProblem(a2 == L2, u2).get_local()
    assert rel(U1, U2) < 1e-12


@pytest.mark.parametrize("size", [2, 3])
@pytest.mark.parametrize("deg,nels", [([2, 2, 2], [3, 4, 5]), ([3, 3, 3], [3, 2, 6]), ([2, 2], [5, 7])])
def test_slab_partitioned_assembly_matches_global(deg, nels, size):
    """Every rank's block of the row-distributed system (assembled with halo
    cell layers, rows it does not own skipped) equals the corresponding rows of
    the oracle's global M^T A M / M^T b, BCs included.  Ranks are emulated one
    after the other on one GPU (no communication is needed for assembly)."""
    from tIGAr import EqualOrderSpline, ExtractedSpline
    from tIGAr.BSplines import ExplicitBSplineControlMesh
    from tigar_b200.api import _Comm

    class FakeComm(_Comm):
        def __init__(self, r, n):
            _Comm.__init__(self, False)
            self._r, self._n = r, n
        rank = property(lambda self: self._r)
        size = property(lambda self: self._n)

    kv = [uk(p, n) for p, n in zip(deg, nels)]
    _, _, pr = make_pair(deg, kv, mode="fused")
    pr.extract()
    pr.assemble(f_np("sin", len(deg)))
    pr.ptap(diag=2.0)
    Cg = pr.C.tocsr()
    rows_seen = 0
    for r in range(size):
        gen = EqualOrderSpline(FakeComm(r, size), 1, ExplicitBSplineControlMesh(deg, kv))
        sp = gen.getScalarSpline(0)
        for d in range(len(deg)):
            for side in (0, 1):
                gen.addZeroDofs(0, sp.getSideDofs(d, side))
        spline = ExtractedSpline(gen, 2 * max(deg))
        assert spline.mode == "fused"
        a, L = poisson_forms(spline, "sin")
        patch = spline.patch()
        pp, pl = patch.pp, patch.plane
        A, b = spline.assembleLinearSystem(a, L)
        # same diag as the oracle run
        A2 = spline.assembleMatrix(a, diag=2.0)
        r0, r1, c0, c1 = pp["k0"] * pl, pp["k1"] * pl, pp["c0"] * pl, pp["c1"] * pl
        ref = Cg[r0:r1, c0:c1]
        assert Cg[r0:r1].nnz == ref.nnz
        assert relm(A2.to_scipy(), ref) < 1e-12
        assert rel(b.get_local(), pr.b[r0:r1]) < 1e-12
        b2 = spline.assembleVector(L)
        assert rel(b2.get_local(), pr.b[r0:r1]) < 1e-12
        rows_seen += r1 - r0
    assert rows_seen == Cg.shape[0]


@pytest.mark.parametrize("deg,nels", [([2], [9]), ([3, 2], [4, 5]), ([3, 3, 3], [3, 4, 3]), ([4, 4], [4, 3]),
                                      ([2, 3, 2], [3, 3, 4])])
def test_kronecker_ptap_equals_generic_ptap_and_scipy(deg, nels):
    """tg_ptap_kron_ap + tg_win_rowcombine (M never read) against the generic
    box-intersection kernels and scipy's M^T A M on a random windowed A."""
    from tigar_b200.engine import TensorPatch, WinMatrix
    from tigar_b200 import dev
    kv = [uk(p, n, -1.0, 2.0) for p, n in zip(deg, nels)]
    patch = TensorPatch(deg, kv)
    assert patch.kron_supported()
    A = WinMatrix(patch.window("A"))
    rng = np.random.RandomState(5)
    A.vals.copy_(dev.from_np(rng.randn(A.window.nnz)))
    M = patch.build_M()
    Ck = patch.ptap_kron(A)
    Cg, AP = patch.ptap(A, M, keep_AP=True)            # generic kernels
    As, Ms = A.to_scipy(), M.to_scipy()
    ref = (Ms.T @ As @ Ms).tocsr()
    assert relm(Ck.to_scipy(), ref) < 1e-13
    assert relm(Cg.to_scipy(), ref) < 1e-13
    assert relm(AP.to_scipy(), (As @ Ms).tocsr()) < 1e-13


@pytest.mark.parametrize("deg,nels", [([3, 2], [4, 5]), ([3, 3, 3], [3, 4, 3]), ([4, 4], [4, 3]),
                                      ([2, 3, 2], [3, 3, 4]), ([3, 3, 3], [9, 7, 8]),
                                      ([2, 2], [40, 37]), ([1, 1, 1], [5, 4, 6]),
                                      ([4, 4, 4], [3, 2, 3]), ([3, 3], [70, 3])])
def test_march_ptap_equals_scipy(deg, nels):
    """tg_ptap_march (two-sided per-direction passes, TMA-staged rows, sliding
    register accumulators) against scipy's M^T A M on a random windowed A, on
    non-uniform knot spacing, and against the row-wise Kronecker kernels."""
    from tigar_b200.engine import TensorPatch, WinMatrix
    from tigar_b200 import dev
    rng = np.random.RandomState(7)
    kv = []
    for p, n in zip(deg, nels):
        k = np.array(uk(p, n, -1.0, 2.0))
        inner = k[p + 1:-(p + 1)]
        if len(inner):
            k[p + 1:-(p + 1)] = inner + 0.3 * (3.0 / n) * (rng.rand(len(inner)) - 0.5)
        kv.append(list(k))
    patch = TensorPatch(deg, kv)
    assert patch._march_setup() is not None
    A = WinMatrix(patch.window("A"))
    A.vals.copy_(dev.from_np(rng.randn(A.window.nnz)))
    M = patch.build_M()
    Cm, stages = patch.ptap_march(A, keep=True)
    As, Ms = A.to_scipy(), M.to_scipy()
    ref = (Ms.T @ As @ Ms).tocsr()
    assert relm(Cm.to_scipy(), ref) < 1e-13
    assert relm(patch.ptap_kron(A).to_scipy(), ref) < 1e-13
    # segments along the march direction give the same rows
    dirs, passes = patch._march
    saved = [(P_["nseg"], P_["seg"], P_["nsegw"], P_["segw"]) for P_ in passes]
    for P_ in passes:
        n = patch.ncp[P_["d"]]
        P_["nseg"] = P_["nsegw"] = max(min(3, n), P_["nsegw"])
        P_["seg"] = P_["segw"] = dev.from_np(
            np.array([(n * k) // P_["nseg"] for k in range(P_["nseg"] + 1)], dtype=np.int32))
    C3 = patch.ptap_march(A)
    assert relm(C3.to_scipy(), ref) < 1e-13
    for P_, sv in zip(passes, saved):
        P_["nseg"], P_["seg"], P_["nsegw"], P_["segw"] = sv
    # the CTA-tiled variants (cp.async and TMA-staged rows) stay in the tree: same result
    old = patch.MARCH_VARIANT
    try:
        for v in (0, 1, 2):
            patch.MARCH_VARIANT = v
            assert relm(patch.ptap_march(A).to_scipy(), ref) < 1e-13
    finally:
        patch.MARCH_VARIANT = old


def test_jit_kernel_equals_interpreter():
    """NVRTC-compiled Gauss-point kernel vs the register-machine interpreter
    (tg_qp_eval) on the same forms: identical systems to rounding (the JIT may
    contract a*b+c into FMAs)."""
    import os
    deg, nels = [3, 3, 3], [3, 4, 3]
    kv = [uk(p, n) for p, n in zip(deg, nels)]
    out = {}
    for flag in ("jit", "interp"):
        if flag == "interp":
            os.environ["TIGAR_B200_NO_JIT"] = "1"
        try:
            gen, spline, pr = make_pair(deg, kv, mode="fused")
            a, L = poisson_forms(spline, "sin")
            A, b = spline.assembleLinearSystem(a, L)
            out[flag] = (A.to_scipy(), b.get_local())
        finally:
            os.environ.pop("TIGAR_B200_NO_JIT", None)
    assert relm(out["jit"][0], out["interp"][0]) < 1e-14
    assert rel(out["jit"][1], out["interp"][1]) < 1e-14


@pytest.mark.parametrize("deg,nels", [([2], [40]), ([3, 2], [37, 5]), ([3, 3, 3], [11, 3, 4]), ([2, 2, 2], [2, 2, 2])])
def test_sell_and_row_major_layouts_agree(deg, nels):
    """The IGA system matrix in SELL-H layout (default) and in row-major layout
    (TIGAR_B200_LAYOUT=0): identical CSR export, SpMV, BCs, diagonal and CG."""
    import os
    from tigar_b200 import dev
    kv = [uk(p, n) for p, n in zip(deg, nels)]
    res = {}
    for layout in ("1", "0"):
        os.environ["TIGAR_B200_LAYOUT"] = layout
        try:
            gen, spline, pr = make_pair(deg, kv, mode="fused")
            assert spline.patch().window("C").layout == int(layout)
            a, L = poisson_forms(spline, "sin")
            C0 = spline.assembleMatrix(a, applyBCs=False)
            C = spline.assembleMatrix(a, diag=2.5)
            b = spline.assembleVector(L)
            rng = np.random.RandomState(2)
            x = rng.rand(C.window.ncols)
            y = dev.to_np(C0.matvec(dev.from_np(x)))
            sol, its, relres = spline.patch().solve_cg(C, b.t, rtol=1e-13)
            res[layout] = (C0.to_scipy(), C.to_scipy(), y, dev.to_np(sol), x)
        finally:
            os.environ.pop("TIGAR_B200_LAYOUT", None)
    s, r = res["1"], res["0"]
    # same entries up to summation order (row-major: march kernels, SELL: element kernels)
    scale = abs(r[0]).max()
    assert abs(s[0] - r[0]).max() < 1e-13 * scale and abs(s[1] - r[1]).max() < 1e-13 * scale
    assert np.abs(s[2] - s[0] @ s[4]).max() < 1e-12 and np.abs(r[2] - r[0] @ r[4]).max() < 1e-12
    assert rel(s[3], r[3]) < 1e-10
    import scipy.sparse.linalg as spla
    assert (s[1].diagonal()[np.unique(spline.zeroDofs)] == 2.5).all()


def test_nonzero_bc_via_newton_and_projection():
    """The flow of poisson-nonzero-bc.py:92-105: L2-project the exact solution
    (nonzero on the boundary) into the spline space, then one Newton step with
    the Gateaux-derivative Jacobian keeps the boundary data; p=2, rate ~3."""
    from tIGAr import TestFunction, inner, cos, assemble, derivative
    errs = []
    for nel in (8, 16):
        deg = [2, 2]
        kv = [uk(p, nel) for p in deg]
        gen, spline, pr = make_pair(deg, kv, mode="fused")
        x = spline.spatialCoordinates()
        soln = cos(PI * x[0]) * cos(PI * x[1])
        f = -spline.div(spline.grad(soln))
        u = spline.project(soln, rationalize=False)
        v = TestFunction(spline.V)
        residual = (inner(spline.grad(u), spline.grad(v)) - inner(f, v)) * spline.dx
        jacobian = derivative(residual, u)
        spline.relativeTolerance = 1e-8
        spline.solveNonlinearVariationalProblem(residual, jacobian, u)
        errs.append(math.sqrt(assemble(((u - soln) ** 2) * spline.dx)))
    rate = math.log(errs[0] / errs[1]) / math.log(2.0)
    assert errs[1] < 2e-4 and 2.6 < rate < 3.6, (errs, rate)


def test_fe_vector_write_refreshes_the_iga_dofs():
    """``u.vector().set_local(...)`` (the reference idiom) writes FE coefficients;
    the element-fused path reads IGA DoFs -- they are re-derived (FEtoIGA) instead of being read
    stale."""
    from tIGAr import Function, TestFunction, inner
    deg, nels = [2, 2], [6, 5]
    kv = [uk(p, n) for p, n in zip(deg, nels)]
    gen, spline, pr = make_pair(deg, kv, mode="fused")
    rng = np.random.RandomState(7)
    Uv = rng.rand(spline.patch().n_iga)
    ref = Function(spline.V)
    ref.set_iga(dev_from(Uv))
    v = TestFunction(spline.V)
    b_ref = spline.assembleVector(inner(ref, v) * spline.dx, applyBCs=False).get_local()
    w = Function(spline.V)
    w.set_iga(dev_from(np.zeros_like(Uv)))            # stale IGA data ...
    w.vector().set_local(ref.vector().get_local())    # ... overwritten through the FE vector
    b = spline.assembleVector(inner(w, v) * spline.dx, applyBCs=False).get_local()
    assert rel(b, b_ref) < 1e-9


def dev_from(a):
    from tigar_b200 import dev
    return dev.from_np(np.asarray(a, dtype=np.float64))
