"""CPU run of the SCALAR hot path's Python (the code bench.py and smoke() drive:
EqualOrderSpline -> ExtractedSpline(...) -> assembleLinearSystem -> solveLinearSystem,
error functionals, project, Newton, writeExtraction / ExtractedSpline(dirname)) with
``engine.TensorPatch`` replaced by the host stand-in of test_multifield_glue_cpu.py and the
C-ABI calls by numpy twins.  Guards the API layer against regressions between GPU runs; the
kernels themselves are covered by the ``-m gpu`` tests.
"""
import math

import numpy as np
import pytest
import torch

from oracle import bsplines as OB
from oracle import pipeline as OP
from test_multifield_glue_cpu import FakePatch, FakeWinMatrix, cpu_backend  # noqa: F401


class ScalarFakePatch(FakePatch):
    """+ the scalar-path entry points of engine.TensorPatch."""

    def __init__(self, degrees, kvecs, quadDeg=None, eps=None, splines=None, part=None, lib=None):
        deg = [s.p for s in splines]
        ts = OB.TensorSpline(deg, [np.asarray(s.knots, dtype=float) for s in splines])
        self.quadDeg = 2 * max(deg) if quadDeg is None else quadDeg
        FakePatch.__init__(self, ts, None, self.quadDeg // 2 + 1, lib)
        self.degrees, self.splines, self.eps = deg, splines, eps
        self.nel = [s.nel for s in ts.splines]
        self.order = 4

    def _integrator(self, kind, funcs):
        from test_multifield_cpu import HostIntegrator
        from oracle import assembly as OA
        if kind not in self._H:
            H = HostIntegrator(self.ts, None, self.nq, {}, order=self.order)
            if kind == "fe":
                pf = self.ts.getDegree()
                H.tabs = [OA.tab_fe(s, pf, self.nq, self.order) for s in self.ts.splines]
                ncell = int(np.prod([tb.T.shape[0] for tb in H.tabs]))
                H.blk = OA.CellBlock(H.tabs, np.arange(ncell), self.order)
                H.n = self.n_fe
            self._H[kind] = H
        H = self._H[kind]
        H.funcs = {fid: funcs[fid].numpy() for fid in list(funcs.keys())}
        return H

    def assemble_system(self, mterms, vterms, funcs, kind="fe"):
        return (self.assemble_matrix(mterms, funcs, kind),
                self.assemble_vector(vterms, funcs, kind))

    def assemble_scalar(self, node, funcs, kind="fe"):
        return float(self._integrator(kind, funcs)._eval([node])[0].sum())

    def fe_node_coords(self):
        from oracle import extraction as OX
        return OX.fe_node_coords(self.ts)

    @property
    def nfe(self):
        from oracle import extraction as OX
        return OX.n_fe_nodes(self.ts)


@pytest.fixture
def scalar_backend(cpu_backend, monkeypatch):  # noqa: F811
    from tigar_b200 import api as A
    monkeypatch.setattr(A, "TensorPatch",
                        lambda *a, **k: ScalarFakePatch(*a, lib=cpu_backend, **k))
    return cpu_backend


def make(deg, nels, mode, nLayers=1, lo=0.0, hi=1.0):
    from tIGAr import EqualOrderSpline, ExtractedSpline
    from tIGAr.BSplines import ExplicitBSplineControlMesh, uniformKnots
    kv = [uniformKnots(p, lo, hi, n) for p, n in zip(deg, nels)]
    gen = EqualOrderSpline(1, ExplicitBSplineControlMesh(deg, kv))
    sp = gen.getScalarSpline(0)
    for d in range(len(deg)):
        for side in (0, 1):
            gen.addZeroDofs(0, sp.getSideDofs(d, side, nLayers))
    spline = ExtractedSpline(gen, 2 * max(deg), mode=mode)
    return gen, spline, kv


@pytest.mark.parametrize("mode", ["fused", "csr", "matfree"])
def test_poisson_through_the_real_constructor_path(scalar_backend, mode, monkeypatch):
    """The flow of demos/poisson/poisson.py (and of bench.one_step / smoke())."""
    monkeypatch.setenv("TIGAR_B200_MF_FUSED", "0")          # matfree: the two-kernel path
    from tIGAr import (TrialFunction, TestFunction, Function, KrylovSolver, inner, sin, pi,
                       assemble)
    deg, nels = [2, 2], [6, 5]
    gen, spline, kv = make(deg, nels, mode)
    assert spline.mode == mode and list(spline.zeroDofs) == sorted(set(gen.zeroDofs))
    pr = OP.Problem(deg, [np.asarray(k) for k in kv])
    Uo = pr.run(lambda X: 2 * math.pi ** 2 * np.prod(np.sin(math.pi * X[..., :2]), axis=-1))
    u, v = TrialFunction(spline.V), TestFunction(spline.V)
    x = spline.spatialCoordinates()
    soln = sin(pi * x[0]) * sin(pi * x[1])
    f = -spline.div(spline.grad(soln))
    ks = KrylovSolver("cg", "jacobi")
    ks.parameters["relative_tolerance"] = 1e-13
    spline.setSolverOptions(linearSolver=ks)
    uh = Function(spline.V)
    U = spline.solveLinearVariationalProblem(
        inner(spline.grad(u), spline.grad(v)) * spline.dx == inner(f, v) * spline.dx, uh)
    assert np.linalg.norm(U.get_local() - Uo) < 1e-10 * np.linalg.norm(Uo)
    err = math.sqrt(assemble(((uh - soln) ** 2) * spline.dx))
    assert abs(err - pr.error(Uo, "l2", lambda X: np.prod(np.sin(math.pi * X[..., :2]), axis=-1))) < 1e-9
    if mode != "matfree":
        MTAM, MTb = spline.assembleLinearSystem(
            inner(spline.grad(u), spline.grad(v)) * spline.dx, inner(f, v) * spline.dx)
        assert abs(MTAM.to_scipy() - pr.C).max() < 1e-12 * abs(pr.C).max()
        assert np.abs(MTb.get_local() - pr.b).max() < 1e-12 * np.abs(pr.b).max()


def test_biharmonic_residual_form_and_two_bc_layers(scalar_backend):
    """demos/biharmonic/biharmonic.py: residual split with lhs/rhs, laplacians through
    div(grad()), two layers of zero DoFs."""
    from tIGAr import TrialFunction, TestFunction, Function, KrylovSolver, inner, cos, pi
    deg, nels = [3, 3], [5, 5]
    gen, spline, kv = make(deg, nels, "fused", nLayers=2, lo=-1.0, hi=1.0)
    u, v = TrialFunction(spline.V), TestFunction(spline.V)
    x = spline.spatialCoordinates()
    soln = (cos(pi * x[0]) + 1.0) * (cos(pi * x[1]) + 1.0)

    def lap(w):
        return spline.div(spline.grad(w))
    f = lap(lap(soln))
    res = inner(lap(u), lap(v)) * spline.dx - inner(f, v) * spline.dx
    ks = KrylovSolver("cg", "jacobi")
    ks.parameters["relative_tolerance"] = 1e-13
    spline.setSolverOptions(linearSolver=ks)
    uh = Function(spline.V)
    U = spline.solveLinearVariationalProblem(res, uh).get_local()
    pr = OP.Problem(deg, [np.asarray(k) for k in kv], form="biharmonic", nLayers=2)

    def f_np(X):
        cx, cy = np.cos(math.pi * X[..., 0]), np.cos(math.pi * X[..., 1])
        return math.pi ** 4 * (cx * (cy + 1) + 2 * cx * cy + (cx + 1) * cy)
    Uo = pr.run(f_np)
    assert np.linalg.norm(U - Uo) < 1e-8 * np.linalg.norm(Uo)


def test_newton_projection_and_extraction_round_trip(scalar_backend, tmp_path):
    from tIGAr import (TrialFunction, TestFunction, Function, KrylovSolver, ExtractedSpline,
                       inner, derivative, sin, pi)
    deg, nels = [2, 2], [5, 4]
    gen, spline, kv = make(deg, nels, "fused")
    ks = KrylovSolver("cg", "jacobi")
    ks.parameters["relative_tolerance"] = 1e-13
    spline.setSolverOptions(maxIters=8, relativeTolerance=1e-10, linearSolver=ks)
    v = TestFunction(spline.V)
    x = spline.spatialCoordinates()
    uh = Function(spline.V)
    # nonlinear reaction-diffusion residual, Newton with the Gateaux tangent
    res = (inner(spline.grad(uh), spline.grad(v)) + (uh + uh ** 3) * v
           - 10.0 * sin(pi * x[0]) * sin(pi * x[1]) * v) * spline.dx
    spline.solveNonlinearVariationalProblem(res, derivative(res, uh), uh)
    R = spline.assembleVector(res).get_local()
    assert np.abs(R).max() < 1e-8
    # L2 projection (common.py:1392-1433) reproduces a function of the spline space
    g = spline.project(x[0] * x[1] + 0.5, rationalize=False)
    w = Function(spline.V)
    w.set_iga(g.iga.clone())
    from tIGAr import assemble
    assert assemble(((w - (x[0] * x[1] + 0.5)) ** 2) * spline.dx) < 1e-20
    # lumped-mass projection (common.py:1416-1430): U = (M^T b) / (M^T m); exact for a constant
    gl = spline.project(0.0 * x[0] + 2.5, rationalize=False, lumpMass=True)
    assert np.allclose(gl.iga.numpy(), 2.5, rtol=1e-13)
    gq = spline.project(x[0] * x[1] + 0.5, rationalize=False, lumpMass=True)
    m = spline.assembleVector(1.0 * v * spline.dx, applyBCs=False).get_local()
    b = spline.assembleVector((x[0] * x[1] + 0.5) * v * spline.dx, applyBCs=False).get_local()
    assert np.allclose(gq.iga.numpy(), b / m, rtol=1e-13)
    # on-disk round trip
    d = str(tmp_path / "extraction")
    gen.writeExtraction(d)
    sp2 = ExtractedSpline(d, 4, mode="fused")
    assert list(sp2.zeroDofs) == list(spline.zeroDofs) and sp2.nsd == spline.nsd
    u2, v2 = TrialFunction(sp2.V), TestFunction(sp2.V)
    A1 = sp2.assembleMatrix(inner(sp2.grad(u2), sp2.grad(v2)) * sp2.dx)
    u, vv = TrialFunction(spline.V), TestFunction(spline.V)
    A0 = spline.assembleMatrix(inner(spline.grad(u), spline.grad(vv)) * spline.dx)
    assert abs(A1.to_scipy() - A0.to_scipy()).max() < 1e-13


def test_bench_one_step_runs_on_the_emulated_backend(scalar_backend, monkeypatch):
    """bench.one_step (the timed body of bench.py) end to end on the host stand-ins, with
    resident control-net columns and with the host control net; result against the
    oracle."""
    import bench as B

    class _Event(object):
        def __init__(self, enable_timing=True):
            pass

        def record(self):
            pass
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    nel = 3
    kv, cm, pinned = B.build_inputs(nel)
    cols = [pinned[:, i].contiguous() for i in range(4)]
    pr = OP.Problem([3] * 3, [OB.uniform_knots(3, 0.0, 1.0, nel)] * 3)
    Uo = pr.run(lambda X: 3 * math.pi ** 2 * np.prod(np.sin(math.pi * X), axis=-1))
    for net, to_host in ((cols, False), (pinned, True)):
        n, its, ev, res, MTAM = B.one_step(kv, cm, net, "fused", 1e-12, to_host)
        Uv = res if to_host else res.numpy()
        assert n == len(Uo) == (nel + 3) ** 3 and its > 0 and len(ev) == 5
        assert np.linalg.norm(Uv - Uo) < 1e-9 * np.linalg.norm(Uo)
        assert B.spmv_bytes.__name__ == "spmv_bytes" and hasattr(MTAM, "window")


REF_DEMOS = "/root/reference/demos"


@pytest.mark.skipif(not __import__("os").path.isdir(REF_DEMOS),
                    reason="the reference tree is only present in the build container")
@pytest.mark.parametrize("demo,rate", [("poisson/poisson.py", 4.0), ("biharmonic/biharmonic.py", 3.0)])
def test_reference_demos_run_unmodified_on_the_api_layer(demo, rate, tmp_path):
    """north_star: demos/poisson and demos/biharmonic run UNMODIFIED against the new
    backend.  Here the reference's own scripts are executed as they lie in the read-only
    reference tree against ``tIGAr/`` with the device layer replaced by host stand-ins
    (tests/run_emulated.py); the printed convergence rates are the ones the demos'
    comments promise (poisson.py:26-28, biharmonic.py:22-27)."""
    import os
    import re
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, os.path.join(here, "run_emulated.py"),
                          os.path.join(REF_DEMOS, demo)], cwd=str(tmp_path),
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    rates = [float(x) for x in re.findall(r"\(rate = ([0-9][0-9.eE+-]*)\)", out.stdout)]
    assert len(rates) == 2 and all(abs(r - rate) < 0.15 for r in rates), out.stdout[-1000:]


@pytest.mark.parametrize("script,args,pattern,check", [
    ("poisson_explicit.py", ["2", "4", "3"], r"rate ([0-9.]+)\)", lambda r: abs(r[-1] - 3.0) < 0.2),
    ("biharmonic.py", ["3", "4", "2"], r"rate ([0-9.]+)\)", lambda r: abs(r[-1] - 2.0) < 0.5),
    ("poisson_annulus.py", ["2", "4", "2"], r"rate ([0-9.]+)\)", lambda r: r[-1] > 2.5),
    ("elasticity.py", ["2", "4", "3"], r"rate ([0-9.]+)\)", lambda r: abs(r[-1] - 3.0) < 0.2),
    ("scordelis_lo.py", ["6", "0.001", "jacobi"], r"Relative norm: ([0-9.eE+-]+)", lambda r: r[-1] < 1e-6),
])
def test_examples_run_on_the_api_layer(script, args, pattern, check, tmp_path):
    """examples/*.py through tests/run_emulated.py at tiny sizes: the scripts stay in step
    with the API (their device runs are recorded in profiles/r1_examples.log)."""
    import os
    import re
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    ex = os.path.join(os.path.dirname(here), "examples", script)
    out = subprocess.run([sys.executable, os.path.join(here, "run_emulated.py"), ex] + args,
                         cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    vals = [float(x) for x in re.findall(pattern, out.stdout)]
    assert vals and check(vals), out.stdout[-1500:]


class _EmuBasis(object):
    """What matfree / jit.launch_op read from engine.Basis: ``c`` mirrors the tg_basis
    struct with HOST addresses of the oracle's tables."""

    def __init__(self, patch, nder):
        from oracle import assembly as OA

        class _C(object):
            pass
        self.nder = nder
        self.nloc = [s.p + 1 for s in patch.ts.splines]
        tabs = [OA.tab_iga(s, patch.nq, nder) for s in patch.ts.splines]
        self.keep = []
        c = self.c = _C()
        c.dim = patch.dim
        c.nq = [patch.nq] * patch.dim + [1] * (3 - patch.dim)
        c.tab, c.idx, c.wq, c.xq = [0] * 3, [0] * 3, [0] * 3, [0] * 3
        c.n, c.nel = [1] * 3, [1] * 3
        for d, tb in enumerate(tabs):
            arrs = [np.ascontiguousarray(tb.T, dtype=np.float64),
                    np.ascontiguousarray(tb.idx, dtype=np.int32),
                    np.ascontiguousarray(tb.w, dtype=np.float64),
                    np.ascontiguousarray(tb.x, dtype=np.float64)]
            self.keep += arrs
            c.tab[d], c.idx[d], c.wq[d], c.xq[d] = [a.ctypes.data for a in arrs]
            c.n[d], c.nel[d] = int(tb.n), int(tb.T.shape[0])


class _EmuJitLib(object):
    """tg_jit_compile / tg_jit_launch on host threads: the kernel handle is the host build
    of the generated source, the launch receives the very OpArgs block jit.launch_op
    filled."""

    def __init__(self):
        self.libs = []
        self.launches = 0

    def tg_jit_compile(self, src, name, handle_ref):
        import ctypes
        import cuda_emu
        from tigar_b200.jit import OpArgs
        lib = cuda_emu.build(src.decode(), name.decode())
        lib.emu_run.argtypes = [ctypes.POINTER(OpArgs), ctypes.c_int, ctypes.c_int]
        self.libs.append(lib)
        handle_ref._obj.value = len(self.libs)            # 1-based handle
        return 0

    def tg_jit_launch(self, h, grid, nth, smem, args_ref, nbytes, stream):
        self.launches += 1
        self.libs[h.value - 1].emu_run(args_ref, int(grid), int(nth))
        return 0


def test_matfree_generated_kernels_through_the_operator_glue(scalar_backend, monkeypatch):
    """TIGAR_B200_MF_FUSED=1: FormOperator drives the generated operator / diagonal kernels
    (program, jets, derivative order, test multi-indices, function pointers, colour stride).
    The kernels run on host threads (tests/cuda_emu.py) in place of the NVRTC module."""
    import ctypes
    import cuda_emu
    from tigar_b200 import jit
    from tIGAr import TrialFunction, TestFunction, Function, KrylovSolver, inner, sin
    monkeypatch.setenv("TIGAR_B200_MF_FUSED", "1")
    monkeypatch.setattr(ScalarFakePatch, "basis", lambda self, kind, nder: _EmuBasis(self, nder),
                        raising=False)
    emu = _EmuJitLib()
    monkeypatch.setattr(jit, "lib", emu)
    monkeypatch.setattr(jit, "check", lambda rc: None)
    monkeypatch.setattr(jit, "_cache", {})

    deg, nels = [2, 2], [5, 4]
    gen, spline, kv = make(deg, nels, "matfree")
    gen2, fused, _ = make(deg, nels, "fused")

    def forms(sp):
        u, v = TrialFunction(sp.V), TestFunction(sp.V)
        x = sp.spatialCoordinates()
        return ((inner(sp.grad(u), sp.grad(v)) + 0.5 * u * v) * sp.dx,
                inner(sin(3.0 * x[0]) + x[1], v) * sp.dx)
    a, L = forms(spline)
    af, Lf = forms(fused)
    C = fused.assembleMatrix(af, diag=1.0)
    op = spline.assembleMatrix(a, diag=1.0)
    xv = torch.from_numpy(np.random.RandomState(1).rand(C.window.nrows))
    assert np.abs(op.matvec(xv).numpy() - C.dense() @ xv.numpy()).max() < 1e-12 * np.abs(C.dense()).max()
    C0 = fused.assembleMatrix(af, applyBCs=False)
    op0 = spline.assembleMatrix(a, applyBCs=False)
    assert np.allclose(op0.jacobi_dinv().numpy(), 1.0 / np.diag(C0.dense()), rtol=1e-12)
    ks = KrylovSolver("cg", "jacobi")
    ks.parameters["relative_tolerance"] = 1e-13
    spline.setSolverOptions(linearSolver=ks)
    fused.setSolverOptions(linearSolver=ks)
    uh, uf = Function(spline.V), Function(fused.V)
    U1 = spline.solveLinearVariationalProblem(a == L, uh).get_local()
    U2 = fused.solveLinearVariationalProblem(af == Lf, uf).get_local()
    assert np.linalg.norm(U1 - U2) < 1e-10 * np.linalg.norm(U2)
    # 9 colours per application (stride p + 1 = 3 in two directions), every one launched
    assert emu.launches % 9 == 0 and emu.launches >= 9 * (spline.lastSolve["iterations"] + 2)
