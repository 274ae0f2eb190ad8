"""Numerics of the form compiler's CUDA back end on the CPU: the source produced by
tigar_b200.jit.generate is compiled unchanged for the host (tests/cuda_emu.py) and compared
with the host interpreter of the same Gauss-point program."""
import numpy as np
import pytest

import cuda_emu
from oracle import assembly as OA
from oracle import bsplines as OB
from tigar_b200 import jit
from tigar_b200 import symbolic as S
from test_multifield_cpu import HostIntegrator, symbolic_spline


def curved_net(ts, dim):
    P = OB.explicit_control_net(ts, 0).copy()
    X = P[:, :dim].copy()
    P[:, 0] = X[:, 0] + 0.07 * np.prod(np.sin(np.pi * X), axis=1)
    P[:, 1] = X[:, 1] + 0.05 * X[:, 0] * (1 - X[:, 0]) * np.cos(1.3 * X[:, 1])
    w = 1.0 + 0.2 * X[:, 0] * X[:, 1]
    P[:, :dim] *= w[:, None]
    P[:, dim] = w
    return P


@pytest.mark.parametrize("deg,nel", [([3, 3, 3], [2, 2, 2]), ([2, 2], [3, 4]), ([4, 4], [2, 2])])
def test_generated_poisson_system_kernel_equals_interpreter(deg, nel):
    """The Gauss-point kernel of the benchmark's integrand (grad u . grad v and f v on a
    rational map: 6 or 9 matrix slots + 1 vector slot, rationalised trial/test functions in
    the quartic case) -- generated CUDA on host threads vs the interpreter."""
    from tigar_b200 import api as A
    from tigar_b200 import ufl_lite as U
    dim = len(deg)
    kv = [OB.uniform_knots(p, 0.0, 1.0, n) for p, n in zip(deg, nel)]
    ts = OB.TensorSpline(deg, kv)
    P = curved_net(ts, dim)
    spl = symbolic_spline(dim, dim, 1, ts.ncp)
    u, v = A.TrialFunction(spl.V), A.TestFunction(spl.V)
    if deg[0] == 4:
        u, v = spl.rationalize(u), spl.rationalize(v)
    x = spl.spatialCoordinates()
    f = U.sin(2.0 * x[0]) * x[1] + 1.0
    a = U.inner(spl.grad(u), spl.grad(v)) * spl.dx
    L = U.inner(f, v) * spl.dx
    mt = spl._weighted(a.scalar())
    vt = spl._weighted(L.scalar())
    outputs = [mt[k] for k in sorted(mt)] + [vt[k] for k in sorted(vt)]
    prog = S.compile_program(outputs, dim)
    fids = sorted(set(j[0] for j in prog.jets))
    jets = [(fids.index(fid), c, al) for (fid, c, al) in prog.jets]
    nder = max(max(al) for (_, _, al) in prog.jets)
    nq = max(deg) + 1
    nloc = [p + 1 for p in deg] + [1] * (3 - dim)
    nqs = [nq] * dim + [1] * (3 - dim)
    src, nth = jit.generate(prog, dim, nloc, nqs, nder + 1, jets, len(fids))
    funcs = {fn.fid: P[:, i].copy() for i, fn in enumerate(spl.cpFuncs)}
    H = HostIntegrator(ts, P, nq, funcs, order=nder)
    tabs = [OA.tab_iga(s, nq, nder) for s in ts.splines]
    ncell = int(np.prod(nel))
    got = cuda_emu.run_qp_kernel(src, nth, tabs, [funcs[fid] for fid in fids], len(outputs),
                                 ncell)
    ref = np.stack(H._eval(outputs), axis=1)
    for s in range(len(outputs)):
        sc = np.abs(ref[:, s]).max()
        assert np.abs(got[:, s] - ref[:, s]).max() <= 1e-12 * max(sc, 1e-300), s


@pytest.mark.parametrize("deg,nel", [([3, 3, 3], [5, 2, 3]), ([2, 2, 2], [3, 3, 2]), ([3, 3], [5, 6])])
def test_fused_operator_kernel_applies_the_matrix(deg, nel):
    """jit.generate(..., op=...): the one-kernel matrix-free operator (Gauss-point program +
    test-function contraction + coloured scatter) on host threads equals C x with C
    integrated on the host from the same bilinear terms."""
    from tigar_b200 import api as A
    from tigar_b200 import ufl_lite as U
    dim = len(deg)
    kv = [OB.uniform_knots(p, 0.0, 1.0, n) for p, n in zip(deg, nel)]
    ts = OB.TensorSpline(deg, kv)
    P = curved_net(ts, dim)
    spl = symbolic_spline(dim, dim, 1, ts.ncp)
    u, v = A.TrialFunction(spl.V), A.TestFunction(spl.V)
    a = (U.inner(spl.grad(u), spl.grad(v)) + 0.7 * u * v) * spl.dx
    mt = spl._weighted(a.scalar())
    xf = A.Function(spl.V)
    vt = {}
    for (aT, aU), node in mt.items():
        term = S.mul(node, S.jet(xf.fid, 0, tuple(aU)))
        vt[aT] = S.add(vt[aT], term) if aT in vt else term
    keys = sorted(vt)
    outputs = [vt[k] for k in keys]
    prog = S.compile_program(outputs, dim)
    fids = sorted(set(j[0] for j in prog.jets))
    jets = [(fids.index(fid), c, al) for (fid, c, al) in prog.jets]
    nder = max(max(max(al) for (_, _, al) in prog.jets), max(max(k) for k in keys))
    nq = max(deg) + 1
    nloc = [p + 1 for p in deg] + [1] * (3 - dim)
    nqs = [nq] * dim + [1] * (3 - dim)
    src, nth = jit.generate(prog, dim, nloc, nqs, nder + 1, jets, len(fids), op=keys)
    assert jit.check_source(src) > 1000                      # NVRTC, sm_100a
    rng = np.random.RandomState(8)
    xv = rng.rand(ts.ncp)
    funcs = {fn.fid: P[:, i].copy() for i, fn in enumerate(spl.cpFuncs)}
    funcs[xf.fid] = xv
    H = HostIntegrator(ts, P, nq, funcs, order=nder)
    Cx = H.matrix({(k[0], k[1]): n for k, n in mt.items()}) @ xv
    tabs = [OA.tab_iga(s, nq, nder) for s in ts.splines]
    y = np.zeros(ts.ncp)
    launches = cuda_emu.run_op_kernel(src, nth, tabs, [funcs[f] for f in fids], y,
                                      [p + 1 for p in deg])
    assert launches == int(np.prod([min(p + 1, n) for p, n in zip(deg, nel)]))
    assert np.abs(y - Cx).max() < 1e-12 * np.abs(Cx).max()


@pytest.mark.parametrize("deg,nel", [([3, 3, 3], [4, 2, 3]), ([2, 2], [4, 5])])
def test_generated_diagonal_kernel_gives_the_matrix_diagonal(deg, nel):
    """jit.generate(..., op=pairs, diag=True): the Jacobi diagonal of the matrix-free
    operator straight from the bilinear terms, on host threads, against the diagonal of the
    host-integrated matrix."""
    from tigar_b200 import api as A
    from tigar_b200 import ufl_lite as U
    dim = len(deg)
    kv = [OB.uniform_knots(p, 0.0, 1.0, n) for p, n in zip(deg, nel)]
    ts = OB.TensorSpline(deg, kv)
    P = curved_net(ts, dim)
    spl = symbolic_spline(dim, dim, 1, ts.ncp)
    u, v = A.TrialFunction(spl.V), A.TestFunction(spl.V)
    a = (U.inner(spl.grad(u), spl.grad(v)) + 0.7 * u * v) * spl.dx
    mt = spl._weighted(a.scalar())
    keys = sorted(mt)
    prog = S.compile_program([mt[k] for k in keys], dim)
    fids = sorted(set(j[0] for j in prog.jets))
    jets = [(fids.index(fid), c, al) for (fid, c, al) in prog.jets]
    nder = max(max(max(al) for (_, _, al) in prog.jets), max(max(k[0] + k[1]) for k in keys))
    nq = max(deg) + 1
    nloc = [p + 1 for p in deg] + [1] * (3 - dim)
    src, nth = jit.generate(prog, dim, nloc, [nq] * dim + [1] * (3 - dim), nder + 1, jets,
                            len(fids), op=keys, diag=True)
    assert jit.check_source(src) > 1000
    funcs = {fn.fid: P[:, i].copy() for i, fn in enumerate(spl.cpFuncs)}
    H = HostIntegrator(ts, P, nq, funcs, order=nder)
    ref = H.matrix({(k[0], k[1]): n for k, n in mt.items()}).diagonal()
    tabs = [OA.tab_iga(s, nq, nder) for s in ts.splines]
    y = np.zeros(ts.ncp)
    cuda_emu.run_op_kernel(src, nth, tabs, [funcs[f] for f in fids], y, [p + 1 for p in deg])
    assert np.abs(y - ref).max() < 1e-12 * np.abs(ref).max()
