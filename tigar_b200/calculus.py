"""
Module-level calculus helpers of the reference's ``tIGAr/calculusUtils.py``
(exported through ``from tIGAr import *``), written on the ``ufl_lite`` form
language: the metric, pseudo-inverse and volume element of a mapping and the
Cartesian gradient / divergence / curl in the deformed configuration
(calculusUtils.py:18-24, 56-69, 255-302), plus the 1-D Gauss rules used for
through-thickness integration (calculusUtils.py:412-470).

``grad`` is the derivative w.r.t. the parametric (mesh) coordinates, as in the
reference where FEniCS's spatial coordinates are the parametric ones; its
dimension is that of the most recent ``ExtractedSpline`` (``ufl_lite.grad``).
Curvilinear tensors, Christoffel symbols and the pushforwards of compatible
spaces are not built (outside the Poisson / biharmonic hot path).
"""
import numpy as np

from . import ufl_lite as U


def getMetric(F):
    """g = DF^T DF (calculusUtils.py:18-24)."""
    DF = U.grad(F)
    return U.dot(DF.T, DF)


def pinvD(F):
    """Moore-Penrose pseudo-inverse g^-1 DF^T of the derivative of the mapping
    (calculusUtils.py:56-63)."""
    DF = U.grad(F)
    return U.dot(U.inv(getMetric(F)), DF.T)


def volumeJacobian(g):
    """sqrt(det g) (calculusUtils.py:65-69)."""
    return U.sqrt(U.det(g))


def cartesianGrad(f, F):
    """Gradient in spatial Cartesian coordinates, grad(f) . pinvD(F)
    (calculusUtils.py:255-261); appends an axis of length nsd."""
    return U.dot(U.grad(f), pinvD(F))


def cartesianDiv(f, F):
    """Sums the last two indices of cartesianGrad(f, F) (calculusUtils.py:269-276)."""
    g = cartesianGrad(f, F)
    if g.a.ndim < 2:
        raise ValueError("div of a scalar")
    if g.a.shape[-1] != g.a.shape[-2]:
        raise ValueError("cartesianDiv: last index of f must have the spatial dimension")
    out = np.empty(g.a.shape[:-2], dtype=object)
    for i in np.ndindex(*g.a.shape[:-2]):
        tot = U.Scalar()
        for k in range(g.a.shape[-1]):
            tot = tot.add(g.a[i + (k, k)])
        out[i] = tot
    return U.Tensor(out) if out.shape else U.Tensor(out[()])


def cartesianCurl(f, F):
    """calculusUtils.py:278-302."""
    return U.curl_from_grad(f, cartesianGrad(f, F))


def getQuadRule(n):
    """n-point Gauss-Legendre rule on (-1, 1) as (points, weights), each a list
    of ``Constant`` (calculusUtils.py:412-457; the reference stops at n = 4)."""
    n = int(n)
    if n < 1:
        print("ERROR: invalid number of quadrature points requested.")
        raise SystemExit
    x, w = np.polynomial.legendre.leggauss(n)
    x = 0.5 * (x - x[::-1])                       # exact antisymmetry, 0.0 in the middle
    w = 0.5 * (w + w[::-1])
    return [U.Constant(float(v)) for v in x], [U.Constant(float(v)) for v in w]


def getQuadRuleInterval(n, L):
    """Rule for (-L/2, L/2) (calculusUtils.py:459-470)."""
    xi_hat, w_hat = getQuadRule(n)
    return [L * x / 2.0 for x in xi_hat], [L * w / 2.0 for w in w_hat]
