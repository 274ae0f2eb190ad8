"""CPU run of the multi-field API glue (ExtractedSpline.assembleMatrix / assembleVector /
extractMatrix / extractVector / solveLinearSystem / solveNonlinearVariationalProblem for
nFields > 1, multifield.BlockOps) with the C-ABI entry points it calls replaced by numpy
stand-ins working on the RAW POINTERS they are handed (CPU torch tensors have host
addresses), and the scalar assembly replaced by the host integrator of
test_multifield_cpu.py.  What is checked is the product's Python: block bookkeeping,
pointer offsets of the field slices, BC masks, call order of the CG building blocks.
The kernels themselves are covered by the ``-m gpu`` tests.
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import assembly as OA
from oracle import extraction as OX
from oracle import pipeline as OP
from test_multifield_cpu import (HostIntegrator, curved_patch, elasticity_forms, body_force,
                                 MU, LAM)


def _f64(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_double * int(n)).from_address(int(ptr)))


def _u8(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_uint8 * int(n)).from_address(int(ptr)))


class FakeWindow(object):
    def __init__(self, nr, nc):
        self.nrows, self.ncols = nr, nc
        self.nnz = nr * nc

    def ref(self):
        return self


class FakeWinMatrix(object):
    """Dense stand-in of engine.WinMatrix: ``vals`` row-major [nr*nc]."""

    def __init__(self, window, vals=None):
        self.window = window
        self.vals = torch.zeros(window.nnz, dtype=torch.float64) if vals is None else vals

    def dense(self):
        return self.vals.numpy().reshape(self.window.nrows, self.window.ncols)

    def matvec(self, x):
        return torch.from_numpy(self.dense() @ x.numpy())

    def to_scipy(self, drop_eps=None):
        import scipy.sparse as sp
        return sp.csr_matrix(self.dense())


class FakeLib(object):
    """numpy twins of the C-ABI calls the multi-field glue makes (include/tigar_b200.h)."""

    def __init__(self):
        self.calls = []

    def tg_cg_scratch_len(self):
        return 16

    def tg_win_zero_rows_cols(self, w, vals, rowmask, colmask, diag, col_shift, stream):
        self.calls.append("zero_rows_cols")
        A = _f64(vals, w.nnz).reshape(w.nrows, w.ncols)
        rm, cm = _u8(rowmask, w.nrows).astype(bool), _u8(colmask, w.ncols).astype(bool)
        A[rm, :] = 0.0
        A[:, cm] = 0.0
        idx = np.nonzero(rm)[0]
        A[idx, idx] = diag
        return 0

    def tg_win_spmv(self, w, vals, x, y, stream):
        self.calls.append("spmv")
        _f64(y, w.nrows)[:] = _f64(vals, w.nnz).reshape(w.nrows, w.ncols) @ _f64(x, w.ncols)
        return 0

    def tg_win_diag_inv(self, w, vals, col_shift, dinv, stream):
        _f64(dinv, w.nrows)[:] = 1.0 / np.diag(_f64(vals, w.nnz).reshape(w.nrows, w.ncols))
        return 0

    def tg_axpy(self, y, a, x, n, stream):
        self.calls.append("axpy")
        _f64(y, n)[:] += a * _f64(x, n)
        return 0

    def tg_dot(self, a, b, n, scratch, out, stream):
        _f64(out, 1)[0] = _f64(a, n) @ _f64(b, n)
        return 0

    def tg_zero_entries(self, b, mask, n, stream):
        _f64(b, n)[_u8(mask, n).astype(bool)] = 0.0
        return 0

    def tg_cg_init(self, b, y, dinv, r, p, n, scratch, out2, stream):
        rr = _f64(b, n) - _f64(y, n)
        _f64(r, n)[:] = rr
        _f64(p, n)[:] = _f64(dinv, n) * rr
        _f64(out2, 2)[:] = [rr @ (_f64(dinv, n) * rr), rr @ rr]
        return 0

    def tg_cg_axpy_dot(self, x, r, p, q, dinv, n, num, den, scratch, out2, stream):
        a = _f64(num, 1)[0] / _f64(den, 1)[0]
        _f64(x, n)[:] += a * _f64(p, n)
        rr = _f64(r, n)
        rr -= a * _f64(q, n)
        _f64(out2, 2)[:] = [rr @ (_f64(dinv, n) * rr), rr @ rr]
        return 0

    def tg_cg_xpby(self, p, r, dinv, n, num, den, stream):
        b = _f64(num, 1)[0] / _f64(den, 1)[0]
        pp = _f64(p, n)
        pp[:] = _f64(dinv, n) * _f64(r, n) + b * pp
        return 0


class FakePatch(object):
    """Stand-in of engine.TensorPatch: scalar assembly through the host integrator."""
    part = None

    def __init__(self, ts, P, nq, lib):
        self.ts, self.P, self.nq, self.lib = ts, P, nq, lib
        self.dim = ts.nvar
        self.n_iga = ts.ncp
        self.Ms = OX.build_M_kron(ts)
        self.n_fe = self.Ms.shape[0]
        self._H = {}

    def _integrator(self, kind, funcs):
        if kind not in self._H:
            H = HostIntegrator(self.ts, self.P, self.nq, {}, order=1)
            if kind == "fe":
                pf = self.ts.getDegree()
                H.tabs = [OA.tab_fe(s, pf, self.nq, 1) for s in self.ts.splines]
                ncell = int(np.prod([tb.T.shape[0] for tb in H.tabs]))
                H.blk = OA.CellBlock(H.tabs, np.arange(ncell), 1)
                H.n = self.n_fe
            self._H[kind] = H
        H = self._H[kind]
        H.funcs = {fid: funcs[fid].numpy() for fid in list(funcs.keys())}
        return H

    def window(self, name):
        n = {"A": self.n_fe, "C": self.n_iga}[name]
        return FakeWindow(n, n)

    def assemble_matrix(self, terms, funcs, kind="fe", out=None, cache=None):
        if cache is not None:
            cache["calls"] = cache.get("calls", 0) + 1
        K = self._integrator(kind, funcs).matrix(terms).toarray()
        return FakeWinMatrix(self.window("A" if kind == "fe" else "C"),
                             torch.from_numpy(np.ascontiguousarray(K).ravel()))

    def assemble_vector(self, terms, funcs, kind="fe", out=None, cache=None):
        if cache is not None:
            cache["calls"] = cache.get("calls", 0) + 1
        b = torch.from_numpy(self._integrator(kind, funcs).vector(terms))
        if out is None:
            return b
        out += b
        return out

    def build_M(self):
        return FakeWinMatrix(FakeWindow(*self.Ms.shape),
                             torch.from_numpy(self.Ms.toarray().ravel()))

    def mt_vec(self, M, b):
        return torch.from_numpy(M.dense().T @ b.numpy())

    def ptap(self, A, M=None):
        Md = M.dense()
        C = Md.T @ A.dense() @ Md
        return FakeWinMatrix(self.window("C"), torch.from_numpy(np.ascontiguousarray(C).ravel()))

    def bc_mask(self, zeroDofs):
        m = torch.zeros(self.n_iga, dtype=torch.uint8)
        m[torch.from_numpy(np.asarray(zeroDofs, dtype=np.int64))] = 1
        return m

    def apply_bcs_matrix(self, Cm, mask, diag=1.0):
        self.lib.tg_win_zero_rows_cols(Cm.window, Cm.vals.data_ptr(), mask.data_ptr(),
                                       mask.data_ptr(), float(diag), 0, None)
        return Cm

    def apply_bcs_vector(self, b, mask):
        self.lib.tg_zero_entries(b.data_ptr(), mask.data_ptr(), b.numel(), None)
        return b

    def solve_cg(self, Cm, b, x=None, rtol=1e-12, atol=0.0, maxit=100000, check_every=5):
        return torch.from_numpy(np.linalg.solve(Cm.dense(), b.numpy())), 1, 0.0

    def solve(self, Cm, b, x=None, rtol=1e-12, atol=0.0, maxit=100000, method="auto", mask=None,
              diag=1.0):
        return self.solve_cg(Cm, b, x, rtol, atol, maxit) + (method,)


@pytest.fixture
def cpu_backend(monkeypatch):
    from tigar_b200 import api as A
    from tigar_b200 import dev, _lib
    fake = FakeLib()
    monkeypatch.setattr(dev, "device", lambda: torch.device("cpu"))
    monkeypatch.setattr(dev, "stream", lambda: None)
    monkeypatch.setattr(_lib, "lib", fake)
    monkeypatch.setattr(_lib, "check", lambda rc: None if rc == 0 else (_ for _ in ()).throw(RuntimeError(rc)))
    monkeypatch.setattr(A, "lib", fake)
    monkeypatch.setattr(A, "check", _lib.check)
    monkeypatch.setattr(A, "WinMatrix", FakeWinMatrix)
    return fake


def make_spline(mode, p=2, nel=(5, 4)):
    ts, kv, P = curved_patch(p, nel)
    n = ts.ncp
    z = []
    for f in range(2):
        z += [f * n + d for d in ts.getSideDofs(0, 0, 1)]
    z += [n + d for d in ts.getSideDofs(1, 1, 1)]
    return ts, kv, P, n, z


def build(mode, fake, p=2, nel=(5, 4)):
    from tigar_b200 import api as A
    ts, kv, P, n, z = make_spline(mode, p, nel)
    spl = object.__new__(A.ExtractedSpline)
    spl._patch = FakePatch(ts, P, p + 1, fake)
    spl.nsd, spl.nFields, spl.generator, spl.mode = 2, 2, None, mode
    spl.comm = A.selfcomm
    spl.V = A.FunctionSpace(spl, 2)
    spl.V_control = A.FunctionSpace(spl, 1, control=True)
    spl.cpFuncs = []
    for i in range(3):
        f = A.Function(spl.V_control)
        f.set_iga(torch.from_numpy(P[:, i].copy()))
        spl.cpFuncs.append(f)
    spl.zeroDofs = A._sorted_unique(np.array(z, dtype=np.int64))
    spl._M = spl._patch.build_M()
    spl.genericSetup()
    prob = OP.ElasticityProblem([p, p], kv, P, MU, LAM, z)
    return spl, prob, n


@pytest.mark.parametrize("mode", ["fused", "csr"])
def test_multifield_linear_solve_through_the_api(cpu_backend, mode):
    from tigar_b200 import api as A
    spl, prob, n = build(mode, cpu_backend)
    Uo = prob.solve(body_force)
    u, v = A.TrialFunction(spl.V), A.TestFunction(spl.V)
    a, L = elasticity_forms(spl, u, v)
    MTAM, MTb = spl.assembleLinearSystem(a, L)
    assert MTAM.shape == (2 * n, 2 * n) and sorted(MTAM.blocks) == [(0, 0), (0, 1), (1, 0), (1, 1)]
    assert abs(MTAM.to_scipy() - prob.C).max() < 1e-12 * abs(prob.C).max()
    assert np.abs(MTb.get_local() - prob.b).max() < 1e-13 * max(1.0, np.abs(prob.b).max())
    ks = A.KrylovSolver("cg", "jacobi")
    ks.parameters["relative_tolerance"] = 1e-13
    spl.setSolverOptions(linearSolver=ks)
    uh = A.Function(spl.V)
    Uv = spl.solveLinearSystem(MTAM, MTb, uh)
    assert spl.lastSolve["relative_residual"] <= 1e-13
    assert np.linalg.norm(Uv.get_local() - Uo) < 1e-10 * np.linalg.norm(Uo)
    assert torch.equal(uh.comps[1].iga, Uv.t[n:])
    # one SpMV per block per iteration, one axpy per off-diagonal block
    its = spl.lastSolve["iterations"]
    assert cpu_backend.calls.count("spmv") >= 4 * its
    assert cpu_backend.calls.count("axpy") * 2 == cpu_backend.calls.count("spmv")
    # driver entry point with an Equation
    uh2 = A.Function(spl.V)
    U2 = spl.solveLinearVariationalProblem(a == L, uh2)
    assert np.linalg.norm(U2.get_local() - Uo) < 1e-10 * np.linalg.norm(Uo)


def test_multifield_newton_converges_in_one_step_on_the_linear_problem(cpu_backend, monkeypatch):
    """solveNonlinearVariationalProblem (common.py:1304-1348) with J = derivative(R, u)
    on a multi-field residual: one Newton step solves the linear problem.  Run with the
    opt-in program cache: every block's cache entry is hit once per Newton iteration."""
    from tigar_b200 import api as A
    from tigar_b200 import ufl_lite as U
    monkeypatch.setenv("TIGAR_B200_PROG_CACHE", "1")
    spl, prob, n = build("fused", cpu_backend)
    Uo = prob.solve(body_force)
    v = A.TestFunction(spl.V)
    uh = A.Function(spl.V)                       # never assigned: starts from zero
    a_u, L = elasticity_forms(spl, uh, v)
    R = a_u - L
    J = A.derivative(R, uh)
    ks = A.KrylovSolver("cg", "jacobi")
    ks.parameters["relative_tolerance"] = 1e-13
    spl.setSolverOptions(maxIters=4, relativeTolerance=1e-9, linearSolver=ks)
    spl.solveNonlinearVariationalProblem(R, J, uh)
    assert np.linalg.norm(uh.iga.numpy() - Uo) < 1e-9 * np.linalg.norm(Uo)
    calls = sorted((k[0][0], c["calls"]) for k, c in spl._prog_cache.items())
    assert len(calls) == 6 and all(c == 2 for _, c in calls), calls   # 4 blocks + 2 vectors, 2 its


def test_missing_diagonal_block_is_created_for_the_bc_diagonal(cpu_backend):
    from tigar_b200 import api as A
    spl, prob, n = build("fused", cpu_backend)
    u, v = A.TrialFunction(spl.V), A.TestFunction(spl.V)
    form = (u[0] * v[0] + u[1] * v[0]) * spl.dx          # nothing tested against v[1]
    Cm = spl.assembleMatrix(form, diag=7.0)
    assert sorted(Cm.blocks) == [(0, 0), (0, 1), (1, 1)]
    D = Cm.block(1, 1).dense()
    zl = spl.zeroDofs[spl.zeroDofs >= n] - n
    assert np.count_nonzero(D) == len(zl) and np.all(D[zl, zl] == 7.0)


def test_multifield_generalized_alpha_elastodynamics(cpu_backend):
    """The time loop of the reference's dynamic demos (dynamic-tspline.py:100-128, 247-293)
    on a two-field linear elastodynamics problem: generalized-alpha integrator on
    multi-field Functions, residual at the alpha levels, tangent by derivative(), Newton per
    step -- against the same recurrence written in numpy on the oracle's matrices."""
    from tigar_b200 import api as A
    from tigar_b200 import ufl_lite as U
    from tigar_b200.time_integration import GeneralizedAlphaIntegrator
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    spl, prob, n = build("fused", cpu_backend)
    rho_inf, dt, dens = 0.5, 0.05, 3.0
    y, y0, v0, a0 = (A.Function(spl.V) for _ in range(4))
    rng = np.random.RandomState(5)
    free = np.ones(2 * n, bool)
    free[spl.zeroDofs] = False
    V0 = rng.rand(2 * n) * free
    for f_, val in ((y, np.zeros(2 * n)), (y0, np.zeros(2 * n)), (v0, V0), (a0, np.zeros(2 * n))):
        f_.set_iga(torch.from_numpy(val.copy()))
    ti = GeneralizedAlphaIntegrator(rho_inf, dt, y, (y0, v0, a0))
    z = A.TestFunction(spl.V)
    a_alpha, L = elasticity_forms(spl, ti.x_alpha(), z)
    res = dens * U.inner(ti.xddot_alpha(), z) * spl.dx + a_alpha - L
    dres = A.derivative(res, y)
    ks = A.KrylovSolver("cg", "jacobi")
    ks.parameters["relative_tolerance"] = 1e-14
    spl.setSolverOptions(maxIters=5, relativeTolerance=1e-8, linearSolver=ks)

    # the same scheme in numpy on the oracle's matrices
    K, b = prob.direct_iga(body_force)
    tabs = [OA.tab_iga(s, 3, 1) for s in prob.ts.splines]
    Ms = OA.assemble(tabs, prob.P, "mass")[0]
    M = dens * sp.kron(sp.identity(2), Ms, format="csr")
    am, af = (2 - rho_inf) / (1 + rho_inf), 1 / (1 + rho_inf)
    g = 0.5 + am - af
    be = 0.25 * (1 + am - af) ** 2
    x0, vv, aa = np.zeros(2 * n), V0.copy(), np.zeros(2 * n)
    Aeff = (am / (be * dt * dt)) * M + af * K
    fr = np.nonzero(free)[0]
    lu = spla.splu(Aeff[fr][:, fr].tocsc())
    for step in range(3):
        spl.solveNonlinearVariationalProblem(res, dres, y)
        pred = x0 + dt * vv + 0.5 * dt * dt * (1 - 2 * be) * aa
        rhs = b - M @ ((1 - am) * aa - am * pred / (be * dt * dt)) - (1 - af) * (K @ x0)
        x1 = np.zeros(2 * n)
        x1[fr] = lu.solve(rhs[fr])
        a1 = (x1 - pred) / (be * dt * dt)
        v1 = vv + dt * ((1 - g) * aa + g * a1)
        assert np.linalg.norm(y.iga.numpy() - x1) < 1e-9 * np.linalg.norm(x1), step
        ti.advance()
        x0, vv, aa = x1, v1, a1
        assert np.linalg.norm(v0.iga.numpy() - vv) < 1e-8 * np.linalg.norm(vv)
        assert np.linalg.norm(a0.iga.numpy() - aa) < 1e-7 * np.linalg.norm(aa)
    assert abs(ti.t - (dt + 3 * dt)) < 1e-14


def test_fe_to_iga_round_trip(cpu_backend):
    """SURVEY 8c KAT 5: FEtoIGA(M U) = U (common.py:968-993), scalar and multi-field."""
    from tigar_b200 import api as A
    spl, prob, n = build("csr", cpu_backend)
    spl.cgRelativeTolerance = 1e-13
    rng = np.random.RandomState(2)
    Uv = rng.rand(2 * n)
    w = A.Function(spl.V)
    w.set_iga(torch.from_numpy(Uv.copy()))
    back = spl.FEtoIGA(w).get_local()
    assert np.abs(back - Uv).max() < 1e-10
    # an FE function that is NOT in the spline space comes back as its least-squares fit
    Ms = spl._patch.Ms
    fe = rng.rand(Ms.shape[0])
    g = A.Function(spl.V.sub(0))
    g._fe = torch.from_numpy(fe.copy())
    spl.nFields = 1
    ls = spl.FEtoIGA(g).get_local()
    ref = np.linalg.lstsq(Ms.toarray(), fe, rcond=None)[0]
    assert np.abs(ls - ref).max() < 1e-9


def test_matrix_free_operator_equals_the_assembled_matrix(cpu_backend, monkeypatch):
    """mode="matfree" (tigar_b200/matfree.py): the operator action assembled as a linear
    form equals C x, BCs in operator form equal zeroRowsColumns, and the CG solve through
    the API reproduces the oracle's LU solution."""
    from tigar_b200 import api as A
    from tigar_b200 import ufl_lite as U
    from tigar_b200.matfree import FormOperator
    monkeypatch.setenv("TIGAR_B200_MF_FUSED", "0")          # the two-kernel path
    spl, prob, n = build("fused", cpu_backend)
    spl.nFields = 1
    spl.V = A.FunctionSpace(spl, 1)
    spl.zeroDofs = spl.zeroDofs[spl.zeroDofs < n]
    u, v = A.TrialFunction(spl.V), A.TestFunction(spl.V)
    x = spl.spatialCoordinates()
    a = U.inner(spl.grad(u), spl.grad(v)) * spl.dx + 0.3 * u * v * spl.dx
    L = U.inner(U.sin(2.0 * x[0]) + x[1], v) * spl.dx
    C = spl.assembleMatrix(a, diag=2.5)                   # fused: the matrix itself
    C0 = spl.assembleMatrix(a, applyBCs=False)
    b = spl.assembleVector(L)
    spl.mode = "matfree"
    op = spl.assembleMatrix(a, diag=2.5)
    assert isinstance(op, FormOperator) and op.shape == (n, n)
    rng = np.random.RandomState(4)
    xv = torch.from_numpy(rng.rand(n))
    y = op.matvec(xv)
    assert np.abs(y.numpy() - C.dense() @ xv.numpy()).max() < 1e-12 * np.abs(C.dense()).max()
    op0 = spl.assembleMatrix(a, applyBCs=False)
    y0 = op0.matvec(xv)
    assert np.abs(y0.numpy() - C0.dense() @ xv.numpy()).max() < 1e-12 * np.abs(C0.dense()).max()
    # Jacobi diagonal from the (transiently) assembled matrix
    assert np.allclose(op0.jacobi_dinv().numpy(), 1.0 / np.diag(C0.dense()), rtol=1e-13)
    # solve through the API
    ks = A.KrylovSolver("cg", "jacobi")
    ks.parameters["relative_tolerance"] = 1e-13
    spl.setSolverOptions(linearSolver=ks)
    uh = A.Function(spl.V)
    Uv = spl.solveLinearVariationalProblem(a == L, uh).get_local()
    Cd = C.dense().copy()
    z = spl.zeroDofs
    Cd[z, z] = 1.0                                         # default diag of the driver
    ref = np.linalg.solve(Cd, b.get_local())
    assert np.linalg.norm(Uv - ref) < 1e-10 * np.linalg.norm(ref)
    assert not Uv[z].any()
    # the Gauss-point program is compiled once per operator, not once per iteration
    assert spl.lastSolve["iterations"] > 5
