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
Also here, all symbolic: Christoffel symbols, ``CurvilinearTensor`` with raised /
lowered indices, covariant derivative, curvilinear gradient / divergence
(calculusUtils.py:26-36, 83-250), the mapped normal and surface element
(:38-54, 71-81) and the pushforwards of compatible spaces (:307-346).
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


# ---------------------------------------------------------------- curvilinear calculus
def _obj(shape):
    out = np.empty(shape, dtype=object)
    for i in np.ndindex(*shape):
        out[i] = U.Scalar()
    return out


def _wrap(a):
    return U.Tensor(a) if a.shape else U.Tensor(a[()])


def getChristoffel(g):
    """Gamma[a, d, c] = 1/2 g^{ab} (d_d g_{cb} + d_c g_{db} - d_b g_{dc}), first index
    raised (calculusUtils.py:26-36)."""
    g = U.as_tensor(g)
    n = g.a.shape[0]
    ginv, dg = U.inv(g), U.grad(g)                  # dg[i, j, k] = d_k g_ij
    out = _obj((n, n, n))
    for a in range(n):
        for d in range(n):
            for c in range(n):
                tot = U.Scalar()
                for b in range(n):
                    br = dg.a[c, b, d].add(dg.a[d, b, c]).add(dg.a[d, c, b], -1.0)
                    tot = tot.add(ginv.a[a, b].mul(br))
                out[a, d, c] = tot.mul(U.as_tensor(0.5).a[()])
    return U.Tensor(out)


class CurvilinearTensor(object):
    """Components ``T`` in the parametric coordinate chart of a manifold with metric
    ``g``; ``lowered[i]`` tells whether index i is covariant (default: all lowered)
    (calculusUtils.py:83-174)."""

    def __init__(self, T, g, lowered=None):
        self.T = U.as_tensor(T)
        self.g = U.as_tensor(g)
        self.lowered = [True] * self.T.a.ndim if lowered is None else list(lowered)
        if len(self.lowered) != self.T.a.ndim:
            raise ValueError("one raised/lowered flag per index")

    def __add__(self, other):
        return CurvilinearTensor(self.T + other.T, self.g, self.lowered)

    def __sub__(self, other):
        return CurvilinearTensor(self.T - other.T, self.g, self.lowered)

    def __rmul__(self, other):
        return CurvilinearTensor(other * self.T, self.g, self.lowered)

    def raiseLowerIndex(self, i):
        mat = U.inv(self.g) if self.lowered[i] else self.g
        a = self.T.a
        out = _obj(a.shape)
        for idx in np.ndindex(*a.shape):
            tot = U.Scalar()
            for k in range(a.shape[i]):
                src = idx[:i] + (k,) + idx[i + 1:]
                tot = tot.add(a[src].mul(mat.a[k, idx[i]]))
            out[idx] = tot
        low = list(self.lowered)
        low[i] = not low[i]
        return CurvilinearTensor(_wrap(out), self.g, low)

    def raiseIndex(self, i):
        return self.raiseLowerIndex(i) if self.lowered[i] else self

    def lowerIndex(self, i):
        return self if self.lowered[i] else self.raiseLowerIndex(i)

    def sharp(self):
        r = self
        for i in range(self.T.a.ndim):
            r = r.raiseIndex(i)
        return r

    def flat(self):
        r = self
        for i in range(self.T.a.ndim):
            r = r.lowerIndex(i)
        return r

    def rank(self):
        return self.T.a.ndim


def curvilinearInner(T, S_):
    """Full contraction with the metric inserted as the index positions require
    (calculusUtils.py:176-187)."""
    return U.inner(T.sharp().T, S_.flat().T)


def covariantDerivative(T):
    """New (last, lowered) index = the differentiation direction
    (calculusUtils.py:189-211)."""
    a = T.T.a
    n = a.ndim
    gam = getChristoffel(T.g).a
    D = U.grad(T.T).a
    dimc = D.shape[-1]
    out = _obj(D.shape)
    for idx in np.ndindex(*D.shape):
        c = idx[-1]
        tot = D[idx]
        for i in range(n):
            for m in range(a.shape[i]):
                src = idx[:i] + (m,) + idx[i + 1:n]
                if T.lowered[i]:
                    tot = tot.add(a[src].mul(gam[m, idx[i], c]), -1.0)
                else:
                    tot = tot.add(a[src].mul(gam[idx[i], m, c]))
        out[idx] = tot
    assert dimc == T.g.a.shape[0]
    return CurvilinearTensor(_wrap(out), T.g, T.lowered + [True])


def curvilinearGrad(T):
    """Covariant derivative with the new index raised (calculusUtils.py:213-226)."""
    n = T.T.a.ndim
    return covariantDerivative(T).raiseIndex(n)


def curvilinearDiv(T):
    """Contracts the covariant derivative's new index with the LAST raised index
    (calculusUtils.py:228-250)."""
    n = T.T.a.ndim
    raised = [i for i in range(n) if not T.lowered[i]]
    if not raised:
        print("ERROR: Divergence operator requires at least one raised index.")
        raise SystemExit
    j = raised[-1]
    D = covariantDerivative(T).T.a
    shp = D.shape[:j] + D.shape[j + 1:n]
    out = _obj(shp)
    for idx in np.ndindex(*shp):
        tot = U.Scalar()
        for k in range(D.shape[j]):
            tot = tot.add(D[idx[:j] + (k,) + idx[j:] + (k,)])
        out[idx] = tot
    return CurvilinearTensor(_wrap(out), T.g, T.lowered[:j] + T.lowered[j + 1:])


def mappedNormal(N, F, normalize=True):
    """DF g^-1 N, the deformed normal of the area element with parametric normal
    ``N`` (calculusUtils.py:38-54)."""
    n = U.dot(U.grad(F), U.dot(U.inv(getMetric(F)), U.as_tensor(N)))
    return n / U.sqrt(U.inner(n, n)) if normalize else n


def surfaceJacobian(g, N):
    """sqrt(det g  N . g^-1 N) (calculusUtils.py:71-81)."""
    N = U.as_tensor(N)
    return U.sqrt(U.det(g) * U.inner(N, U.dot(U.inv(g), N)))


def cartesianPushforwardN(u, F):
    """Curl-conserving pushforward DF^-T u, 3-D fields on 3-D domains
    (calculusUtils.py:307-318)."""
    return U.dot(U.inv(U.grad(F).T), U.as_tensor(u))


def cartesianPushforwardRT(v, F):
    """Div-conserving pushforward DF v / sqrt(det g) (calculusUtils.py:320-336)."""
    return U.dot(U.grad(F), U.as_tensor(v)) / volumeJacobian(getMetric(F))


def cartesianPushforwardW(phi, F):
    """Mass-conserving pushforward phi / sqrt(det g) (calculusUtils.py:338-346)."""
    return U.as_tensor(phi) / volumeJacobian(getMetric(F))
