"""a user-defined ``AbstractScalarBasis`` (not a tensor-product ``BSpline``
object) through the generic device path -- the reference's per-node extraction loop
(common.py:1497-1509) on the host once, then CSR extraction operator, FE assembly, M^T A M in
operator form and CG on the device (tigar_b200/generic.py) -- against the oracle."""
import math

import numpy as np
import pytest

from gpu_util import rel, uk
from oracle import bsplines as OB
from oracle import pipeline as OP
from oracle import extraction as OX

pytestmark = pytest.mark.gpu
PI = math.pi


def make_basis(deg, kvecs, perm):
    """A basis that evaluates itself in pure Python (Cox-de Boor through the oracle's 1-D
    routine) and numbers its functions through ``perm`` -- so its M has no tensor-product
    window structure the fast path could use."""
    from tIGAr import AbstractScalarBasis
    from tIGAr.BSplines import TensorMesh
    ts = OB.TensorSpline(deg, kvecs)

    class PermutedSplineBasis(AbstractScalarBasis):
        def getNodesAndEvals(self, xi):
            return [[int(perm[i]), v] for i, v in ts.getNodesAndEvals(list(xi))]

        def getNcp(self):
            return ts.ncp

        def generateMesh(self, comm=None):
            return TensorMesh([s.uniqueKnots for s in ts.splines])

        def getDegree(self):
            return max(deg)

        def needsDG(self):
            return False

        def useRectangularElements(self):
            return True

        def getPrealloc(self):
            return int(np.prod([p + 1 for p in deg]))
    return PermutedSplineBasis(), ts


@pytest.mark.parametrize("deg,nels", [([2, 2], [6, 5]), ([3, 2, 2], [3, 4, 3])])
def test_generic_basis_poisson_matches_oracle(deg, nels):
    from tIGAr import (AbstractControlMesh, EqualOrderSpline, ExtractedSpline, TrialFunction,
                       TestFunction, Function, inner, sin)
    dim = len(deg)
    kv = [uk(p, n) for p, n in zip(deg, nels)]
    rng = np.random.RandomState(4)
    ts0 = OB.TensorSpline(deg, kv)
    perm = rng.permutation(ts0.ncp)
    basis, ts = make_basis(deg, kv, perm)
    P0 = OB.explicit_control_net(ts, 0)
    Pp = np.empty_like(P0)
    Pp[perm] = P0

    class Mesh(AbstractControlMesh):
        def getScalarSpline(self):
            return basis

        def getHomogeneousCoordinate(self, node, direction):
            return float(Pp[node, direction])

        def getNsd(self):
            return dim
    gen = EqualOrderSpline(1, Mesh())
    assert not gen.isTensorProduct()
    pr = OP.Problem(deg, kv)
    gen.addZeroDofs(0, [int(perm[i]) for i in pr.zeroDofs])
    spline = ExtractedSpline(gen, 2 * max(deg))
    assert spline.mode == "csr"
    # M: the per-node loop against the oracle's Kronecker build, columns permuted
    Mo = OX.build_M_kron(ts).tocsr()
    Ms = spline.M.to_scipy()
    assert abs(Ms[:, perm] - Mo).max() < 1e-14
    u, v = TrialFunction(spline.V), TestFunction(spline.V)
    x = spline.spatialCoordinates()
    f = 1.0
    for d in range(dim):
        f = f * sin(PI * x[d])
    uh = Function(spline.V)
    U = spline.solveLinearVariationalProblem(
        inner(spline.grad(u), spline.grad(v)) * spline.dx == inner(dim * PI ** 2 * f, v) * spline.dx,
        uh)
    Uo = pr.run(lambda X: dim * PI ** 2 * np.prod(np.sin(PI * X[..., :dim]), axis=-1))
    assert rel(U.get_local()[perm], Uo) < 1e-10
    # the FE function u = M U is numbering-independent
    assert rel(uh.vector().get_local(), pr.M @ Uo) < 1e-10
    C = spline.assembleMatrix(inner(spline.grad(u), spline.grad(v)) * spline.dx).to_scipy()
    assert abs(C[perm][:, perm] - pr.C).max() < 1e-11 * abs(pr.C).max()
