"""Global sum-factorised assembly (csrc/tg_gsf.cu) against the element kernels it replaces and
against the oracle: uniform / non-uniform / repeated knots (march shifts > 1), mixed degrees,
2-D and 3-D, both bases (extracted spline basis and the Lagrange FE basis of the csr mode),
second-derivative forms, and chunking of the last direction (partial sums carried across chunk
boundaries)."""
import math
import os

import numpy as np
import pytest

from gpu_util import make_pair, uk, rel, relm

pytestmark = pytest.mark.gpu
PI = math.pi


def _forms(spline, form):
    from tIGAr import TrialFunction, TestFunction, inner, sin, cos
    u, v = TrialFunction(spline.V), TestFunction(spline.V)
    x = spline.spatialCoordinates()
    f = 1.0
    for d in range(len(x)):
        f = f * sin(PI * x[d]) + 0.3 * cos(x[d])
    if form == "poisson":
        a = inner(spline.grad(u), spline.grad(v)) * spline.dx + (2.0 + x[0]) * inner(u, v) * spline.dx
    else:
        lap = lambda w: spline.div(spline.grad(w))
        a = inner(lap(u), lap(v)) * spline.dx
    return a, inner(f, v) * spline.dx


def _system(deg, kv, mode, form, nLayers, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        gen, spline, pr = make_pair(deg, kv, nLayers=nLayers, form=form, mode=mode)
        a, L = _forms(spline, form)
        C, b = spline.assembleLinearSystem(a, L)
        C2 = spline.assembleMatrix(a)
        b2 = spline.assembleVector(L)
        return C.to_scipy(), b.get_local(), C2.to_scipy(), b2.get_local()
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


CASES = [
    ([3, 3, 3], [uk(3, 5), uk(3, 4), uk(3, 9)], "poisson", 1),
    ([2, 3, 2], [uk(2, 6), uk(3, 3), uk(2, 5)], "poisson", 1),
    ([2, 2], [uk(2, 9), uk(2, 7)], "poisson", 1),
    ([4, 4], [uk(4, 7), uk(4, 6)], "biharmonic", 2),
    ([3, 3], [[0, 0, 0, 0, .2, .2, .5, .55, .55, .9, 1, 1, 1, 1], uk(3, 5)], "poisson", 1),
    ([3, 2, 3], [[0, 0, 0, 0, .3, .3, .3, .7, 1, 1, 1, 1], uk(2, 4),
                 [0, 0, 0, 0, .1, .4, .4, .8, 1, 1, 1, 1]], "poisson", 1),
    ([1, 1], [uk(1, 5), uk(1, 6)], "poisson", 1),
]


@pytest.mark.parametrize("mode", ["fused", "csr"])
@pytest.mark.parametrize("deg,kv,form,nLayers", CASES)
def test_gsf_equals_element_kernels(deg, kv, form, nLayers, mode):
    ref = _system(deg, kv, mode, form, nLayers, {"TIGAR_B200_GSF": "0"})
    new = _system(deg, kv, mode, form, nLayers, {"TIGAR_B200_GSF": "1"})
    # one cell layer per chunk: every partial sum crosses a chunk boundary
    chunked = _system(deg, kv, mode, form, nLayers, {"TIGAR_B200_GSF": "1",
                                                     "TIGAR_B200_GSF_GB": "1e-9"})
    for got in (new, chunked):
        assert relm(got[0], ref[0]) < 1e-12
        assert rel(got[1], ref[1]) < 1e-12
        assert relm(got[2], ref[2]) < 1e-12
        assert rel(got[3], ref[3]) < 1e-12


def test_gsf_matches_the_oracle_and_uses_the_march_kernels():
    from tigar_b200._lib import lib
    deg, nels = [3, 3, 3], [6, 5, 7]
    kv = [uk(p, n) for p, n in zip(deg, nels)]
    gen, spline, pr = make_pair(deg, kv, mode="fused")
    from tIGAr import TrialFunction, TestFunction, inner, sin
    u, v = TrialFunction(spline.V), TestFunction(spline.V)
    x = spline.spatialCoordinates()
    a = inner(spline.grad(u), spline.grad(v)) * spline.dx
    L = inner(3 * PI ** 2 * sin(PI * x[0]) * sin(PI * x[1]) * sin(PI * x[2]), v) * spline.dx
    l0 = lib.tg_launch_count()
    C, b = spline.assembleLinearSystem(a, L)
    launches = lib.tg_launch_count() - l0
    # qp kernel + 3 matrix stages + 3 vector stages + BCs: a handful, not thousands of colours
    assert launches < 20, launches
    pr.extract()
    pr.assemble(lambda X: 3 * PI ** 2 * np.prod(np.sin(PI * X), axis=-1))
    pr.ptap()
    assert relm(C.to_scipy(), pr.C) < 1e-12
    assert rel(b.get_local(), pr.b) < 1e-12


def test_discontinuous_spline_on_the_fused_path():
    """A spline with an interior knot of multiplicity p+1 is discontinuous there; the reference
    extracts it to a DG space (BSplines.py:419-427, common.py:332-351).  The element-fused path
    needs no FE space: system and solution against the oracle's direct IGA Galerkin assembly."""
    import scipy.sparse.linalg as spla
    from tIGAr import Function
    from oracle import pipeline as OP
    deg = [2, 2]
    kv = [[0, 0, 0, .25, .5, .5, .5, .75, 1, 1, 1], uk(2, 5)]
    gen, spline, pr = make_pair(deg, kv, mode=None)
    assert spline.mode == "fused" and gen.extractionElement() == "DG"
    a, L = _forms(spline, "poisson")
    C, b = spline.assembleLinearSystem(a, L)
    f = lambda X: (np.sin(PI * X[..., 0]) + 0.3 * np.cos(X[..., 0])) * np.sin(PI * X[..., 1]) \
        + 0.3 * np.cos(X[..., 1])
    Co, bo = pr.direct_iga(f)
    # the oracle's Poisson form has no mass term: assemble that variant for the matrix check
    from tIGAr import TrialFunction, TestFunction, inner
    u, v = TrialFunction(spline.V), TestFunction(spline.V)
    C1 = spline.assembleMatrix(inner(spline.grad(u), spline.grad(v)) * spline.dx, applyBCs=False)
    assert relm(C1.to_scipy(), Co.tocsr()) < 1e-12
    assert rel(spline.assembleVector(L, applyBCs=False).get_local(), bo) < 1e-12
    uh = Function(spline.V)
    U = spline.solveLinearSystem(C, b, uh)
    Uo = spla.spsolve(C.to_scipy().tocsc(), b.get_local())
    assert rel(U.get_local(), Uo) < 1e-10
    with pytest.raises(NotImplementedError):
        make_pair(deg, kv, mode="csr")
