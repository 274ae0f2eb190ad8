"""Known-answer tests that pin the FE-assembly / PtAP / solve layer of the
oracle.  The reference delegates this layer to
FEniCS/PETSc (absent), so these KATs -- not reference outputs -- are the pin."""
import math

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import assembly as A
from oracle import bsplines as B
from oracle import extraction as X
from oracle import pipeline as P


def _ts(deg, nels, lo=0.0, hi=1.0):
    return B.TensorSpline(deg, [B.uniform_knots(p, lo, hi, n) for p, n in zip(deg, nels)])


@pytest.mark.parametrize("p,nel,nnz", [(2, 8, 40), (3, 10, 109), (4, 7, 131)])
def test_m1d_partition_of_unity_and_nnz(p, nel, nnz):
    s = B.Spline1(p, B.uniform_knots(p, 0.0, 1.0, nel))
    M = X.m1d(s, p)
    assert M.nnz == nnz == (nel - 1) * p + 2 + nel * (p - 1) * (p + 1)
    assert np.abs(np.asarray(M.sum(axis=1)).ravel() - 1.0).max() < 1e-15 * 8
    assert M.data.min() > 0 and M.data.max() <= 1.0


@pytest.mark.parametrize("deg,nels", [([2, 2], [4, 3]), ([3, 2], [3, 4]), ([2, 1, 2], [2, 2, 2])])
def test_M_kron_equals_reference_loop(deg, nels):
    ts = _ts(deg, nels)
    Mk = X.build_M_kron(ts)
    Ml = X.build_M_loop(ts)
    assert (Mk != Ml).nnz == 0          # same pattern, bit-identical values
    assert np.array_equal(Mk.indices, Ml.indices) and np.array_equal(Mk.data, Ml.data)


def test_knot_boundary_rule():
    # node exactly on an interior knot: p entries, left span (BSplines.py:299-308)
    s = B.Spline1(3, B.uniform_knots(3, 0.0, 1.0, 4))
    M = X.m1d(s, 3).tocsr()
    row = 3            # node on the first interior knot
    assert M.indptr[row + 1] - M.indptr[row] == 3
    assert s.getKnotSpan(0.25) == 3    # left span (knots[3] = 0 .. knots[4]=0.25)


def test_linear_reproduction():
    ts = _ts([3, 2], [5, 4], -1.0, 2.0)
    M = X.build_M_kron(ts)
    Pn = B.explicit_control_net(ts)
    cp = X.control_funcs(M, Pn)
    xyz = X.fe_node_coords(ts)
    assert np.abs(cp[:, :2] - xyz).max() < 1e-14
    assert np.abs(cp[:, 2] - 1.0).max() < 1e-14


@pytest.mark.parametrize("deg,nels,form", [([2, 2], [5, 4], "poisson"), ([3, 3], [4, 4], "poisson"),
                                           ([4, 4], [3, 3], "biharmonic"), ([2, 2, 2], [2, 3, 2], "poisson"),
                                           ([3, 3], [3, 3], "mass")])
def test_ptap_equals_direct_iga(deg, nels, form):
    kv = [B.uniform_knots(p, 0.0, 1.0, n) for p, n in zip(deg, nels)]
    pr = P.Problem(deg, kv, form=form, nLayers=2 if form == "biharmonic" else 1)
    f = lambda x: np.sin(3.0 * x[..., 0]) + x[..., 1] ** 2
    pr.extract()
    pr.assemble(f)
    pr.ptap(applyBCs=False)
    Ad, bd = pr.direct_iga(f)
    nrm = spla.norm(Ad)
    assert spla.norm(pr.C0 - Ad) / nrm < 1e-12
    assert np.linalg.norm(pr.b0 - bd) / np.linalg.norm(bd) < 1e-12
    assert spla.norm(pr.C0 - pr.C0.T) / nrm < 1e-13


def test_nurbs_like_geometry_ptap_equals_direct():
    # non-trivial rational geometry: perturb control points and weights
    deg, nels = [2, 2], [4, 4]
    kv = [B.uniform_knots(p, 0.0, 1.0, n) for p, n in zip(deg, nels)]
    ts = B.TensorSpline(deg, kv)
    Pn = B.explicit_control_net(ts)
    rng = np.random.RandomState(7)
    w = 1.0 + 0.3 * rng.rand(ts.ncp)
    Pn[:, :2] += 0.03 * rng.randn(ts.ncp, 2)
    Pn[:, :2] *= w[:, None]
    Pn[:, 2] = w
    pr = P.Problem(deg, kv, form="poisson", P=Pn, rationalize=True)
    f = lambda x: x[..., 0] * x[..., 1]
    pr.extract()
    pr.assemble(f)
    pr.ptap(applyBCs=False)
    Ad, bd = pr.direct_iga(f)
    # numerator and weight are splines, hence represented EXACTLY in the FE
    # space (common.py:917-921): the rational map agrees point-wise
    assert spla.norm(pr.C0 - Ad) / spla.norm(Ad) < 1e-11
    assert np.linalg.norm(pr.b0 - bd) / np.linalg.norm(bd) < 1e-11


def test_bcs():
    deg, nels = [2, 2], [4, 4]
    kv = [B.uniform_knots(p, 0.0, 1.0, n) for p, n in zip(deg, nels)]
    pr = P.Problem(deg, kv)
    pr.extract()
    pr.assemble(lambda x: 1.0 + 0 * x[..., 0])
    pr.ptap(diag=7.0)
    z = np.unique(pr.zeroDofs)
    Cs = P.apply_bcs_matrix(pr.C0, pr.zeroDofs, 7.0)
    assert abs(Cs - pr.C).max() == 0.0
    C = pr.C.toarray()
    for i in z:
        row = C[i].copy()
        assert row[i] == 7.0
        row[i] = 0
        assert not row.any() and not np.delete(C[:, i], i).any()
    assert not pr.b[z].any()


def _poisson_l2(nel, p=3):
    kv = [B.uniform_knots(p, 0.0, 1.0, nel)] * 2
    pr = P.Problem([p, p], kv)
    u = lambda x: np.sin(math.pi * x[..., 0]) * np.sin(math.pi * x[..., 1])
    f = lambda x: 2 * math.pi ** 2 * u(x)
    U = pr.run(f)
    return pr.error(U, "l2", u)


def test_poisson_manufactured_rate():
    # poisson.py:26-28 "converge at optimal rates": p=3 -> 4
    e0, e1 = _poisson_l2(8), _poisson_l2(16)
    assert 3.7 < math.log(e0 / e1) / math.log(2.0) < 4.4


def test_biharmonic_manufactured_rate():
    # biharmonic.py:22-27: energy-norm rate p-1 = 3 for p=4
    errs = []
    for nel in (6, 12):
        kv = [B.uniform_knots(4, -1.0, 1.0, nel)] * 2
        pr = P.Problem([4, 4], kv, form="biharmonic", nLayers=2)
        pi = math.pi
        f = lambda x: pi ** 4 * (np.cos(pi * x[..., 0]) * (np.cos(pi * x[..., 1]) + 1)
                                 + 2 * np.cos(pi * x[..., 0]) * np.cos(pi * x[..., 1])
                                 + (np.cos(pi * x[..., 0]) + 1) * np.cos(pi * x[..., 1]))
        lap = lambda x: -pi ** 2 * (np.cos(pi * x[..., 0]) * (np.cos(pi * x[..., 1]) + 1)
                                   + (np.cos(pi * x[..., 0]) + 1) * np.cos(pi * x[..., 1]))
        U = pr.run(f)
        errs.append(pr.error(U, "energy", lap))
    assert 2.7 < math.log(errs[0] / errs[1]) / math.log(2.0) < 3.5


def test_cantilever_eigenfrequencies():
    # modal-analysis.py:60-85 with EI = mu = L = 1, clamped at x=0:
    # omega = 3.5160, 22.034, 61.697 (textbook)
    p, nel = 3, 40
    kv = [B.uniform_knots(p, 0.0, 1.0, nel)]
    prK = P.Problem([p], kv, form="biharmonic")
    prK.extract(); prK.assemble(None); prK.ptap(applyBCs=False)
    prM = P.Problem([p], kv, form="mass")
    prM.extract(); prM.assemble(None); prM.ptap(applyBCs=False)
    keep = np.arange(2, prK.ts.ncp)            # clamp: first two control points
    K = prK.C0.toarray()[np.ix_(keep, keep)]
    Mm = prM.C0.toarray()[np.ix_(keep, keep)]
    import scipy.linalg as sla
    lam = sla.eigh(K, Mm, eigvals_only=True)[:3]
    om = np.sqrt(lam)
    assert np.allclose(om, [3.5160, 22.034, 61.697], rtol=2e-4)


def test_jacobi_cg_matches_lu():
    kv = [B.uniform_knots(2, 0.0, 1.0, 6)] * 3
    pr = P.Problem([2, 2, 2], kv)
    f = lambda x: 1.0 + x[..., 0]
    Ulu = pr.run(f).copy()
    Ucg = pr.solve("cg", rtol=1e-13)
    assert np.linalg.norm(Ucg - Ulu) / np.linalg.norm(Ulu) < 1e-11
