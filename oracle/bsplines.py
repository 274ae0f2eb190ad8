"""
Oracle restatement of the reference's B-spline layer (numpy, CPU).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Each function cites the reference lines it follows (``/root/reference/tIGAr``).
Two flavours are given where it matters: a scalar one that follows the
reference statement by statement, and a vectorised one used for large sizes;
tests check they agree bit-for-bit.
"""
import numpy as np

DOLFIN_EPS = 3.0e-16                 # dolfin/common/constants.h
KNOT_NEAR_EPS = 10.0 * DOLFIN_EPS    # BSplines.py:42
IGNORE_EPS = 1e-15                   # common.py:56


def near(x, x0, eps=DOLFIN_EPS):
    """dolfin.near: x0-eps <= x <= x0+eps."""
    return (x0 - eps <= x) and (x <= x0 + eps)


def uniform_knots(p, start, end, N, periodic=False, continuityDrop=0):
    """BSplines.py:14-38."""
    if continuityDrop >= p:
        raise ValueError("Continuity drop too high for spline degree.")
    out = []
    if not periodic:
        out += [start] * (p - continuityDrop)
    h = (end - start) / float(N)
    for i in range(N + 1):
        out += [start + float(i) * h] * (continuityDrop + 1)
    if not periodic:
        out += [end] * (p - continuityDrop)
    return out


class Spline1(object):
    """Univariate B-spline, BSplines.py:164-351."""

    def __init__(self, p, knots):
        self.p = int(p)
        self.knots = np.array(knots, dtype=np.float64)
        k = self.knots
        # computeNel, BSplines.py:235-243
        self.nel = 0
        for i in range(1, len(k)):
            if not near(k[i], k[i - 1], KNOT_NEAR_EPS):
                self.nel += 1
        # unique knots / multiplicities, BSplines.py:183-193
        self.uniqueKnots = np.zeros(self.nel + 1)
        self.multiplicities = np.zeros(self.nel + 1, dtype=np.int32)
        ct = -1
        last = None
        for i in range(len(k)):
            if last is None or not near(k[i], last, KNOT_NEAR_EPS):
                ct += 1
                self.uniqueKnots[ct] = k[i]
            last = k[i]
            self.multiplicities[ct] += 1
        self.ncp = len(k) - int(self.multiplicities[0])        # :273-277
        self.nGhost = self.p + 1                                # :197
        self.ghostKnots = np.array(
            [self.getKnot(i) for i in range(-self.nGhost, len(k) + self.nGhost)])

    def getKnot(self, i):
        """BSplines.py:245-260."""
        k = self.knots
        if i < 0:
            ii = len(k) - int(self.multiplicities[-1]) + i
            return k[0] - (k[-1] - k[ii])
        elif i >= len(k):
            ii = i - len(k) + int(self.multiplicities[0])
            return k[-1] + (k[ii] - k[0])
        return k[i]

    def greville(self, i):
        """BSplines.py:262-271."""
        r = 0.0
        for j in range(i, i + self.p):
            r += self.getKnot(j + 1)
        return r / float(self.p)

    def isDiscontinuous(self):
        """BSplines.py:225-233."""
        for i in range(1, len(self.uniqueKnots) - 1):
            if self.multiplicities[i] > self.p:
                return True
        return False

    def getKnotSpan(self, u):
        """BSplines.py:285-308 (left span at an interior knot)."""
        nspans = len(self.knots) - 1
        span = int(np.searchsorted(self.knots, u)) - 1
        lo = int(self.multiplicities[0]) - 1
        hi = nspans - (int(self.multiplicities[-1]) - 1) - 1
        return min(max(span, lo), hi)

    def getNodes(self, u):
        """BSplines.py:310-319."""
        s = self.getKnotSpan(u)
        return [i % self.ncp for i in range(s - self.p, s + 1)]

    def basisFuncs(self, knotSpan, u):
        """BSplines.py:321-351 calling basisFuncsInner :73-120."""
        return basis_funcs_inner(self.ghostKnots, self.nGhost, u, self.p,
                                 knotSpan + 1)

    # ---- vectorised ----
    def spans_vec(self, u):
        u = np.asarray(u, dtype=np.float64)
        nspans = len(self.knots) - 1
        span = np.searchsorted(self.knots, u) - 1
        lo = int(self.multiplicities[0]) - 1
        hi = nspans - (int(self.multiplicities[-1]) - 1) - 1
        return np.clip(span, lo, hi).astype(np.int64)

    def basis_vec(self, span, u):
        return basis_funcs_vec(self.ghostKnots, self.nGhost,
                               np.asarray(u, dtype=np.float64), self.p,
                               np.asarray(span) + 1)

    def element_spans(self):
        """Knot-span index of each non-degenerate element (last of a run of
        repeated knots): consequence of BSplines.py:183-193, 285-308."""
        return np.cumsum(self.multiplicities[:-1]).astype(np.int64) - 1


def basis_funcs_inner(ghostKnots, nGhost, u, pl, i):
    """Statement-by-statement restatement of the reference's only native
    routine, ``basisFuncsInner`` (BSplines.py:73-120): Piegl-Tiller A2.2."""
    N = pl + 1
    ndu = np.zeros((N, N))
    left = np.zeros(N)
    right = np.zeros(N)
    ndu[0, 0] = 1.0
    for j in range(1, pl + 1):
        left[j] = u - ghostKnots[i - j + nGhost]
        right[j] = ghostKnots[i + j - 1 + nGhost] - u
        saved = 0.0
        for r in range(0, j):
            ndu[j, r] = right[r + 1] + left[j - r]
            temp = ndu[r, j - 1] / ndu[j, r]
            ndu[r, j] = saved + right[r + 1] * temp
            saved = left[j - r] * temp
        ndu[j, j] = saved
    return ndu[:, pl].copy()


def basis_funcs_vec(ghostKnots, nGhost, u, pl, i):
    """Same recurrence, same operation order, vectorised over points
    (bit-identical to ``basis_funcs_inner``: no reassociation)."""
    n = u.shape[0]
    N = pl + 1
    ndu = np.zeros((N, N, n))
    left = np.zeros((N, n))
    right = np.zeros((N, n))
    ndu[0, 0] = 1.0
    for j in range(1, pl + 1):
        left[j] = u - ghostKnots[i - j + nGhost]
        right[j] = ghostKnots[i + j - 1 + nGhost] - u
        saved = np.zeros(n)
        for r in range(0, j):
            ndu[j, r] = right[r + 1] + left[j - r]
            temp = ndu[r, j - 1] / ndu[j, r]
            ndu[r, j] = saved + right[r + 1] * temp
            saved = left[j - r] * temp
        ndu[j, j] = saved
    return ndu[:, pl, :].T.copy()          # [n, p+1]


# index helpers, BSplines.py:354-370
def ij2dof(i, j, M):
    return j * M + i


def ijk2dof(i, j, k, M, N):
    return k * (M * N) + j * M + i


def dof2ij(dof, M):
    return (dof % M, dof // M)


def dof2ijk(dof, M, N):
    ij = dof % (M * N)
    return (ij % M, ij // M, dof // (M * N))


class TensorSpline(object):
    """Uni-/bi-/tri-variate B-spline, BSplines.py:374-649."""

    def __init__(self, degrees, kvecs):
        self.nvar = len(degrees)
        assert 1 <= self.nvar <= 3
        self.splines = [Spline1(degrees[d], kvecs[d]) for d in range(self.nvar)]
        self.ncp = int(np.prod([s.ncp for s in self.splines]))
        self.nel = int(np.prod([s.nel for s in self.splines]))

    def getNcp(self):
        return self.ncp

    def getDegree(self):
        """BSplines.py:580-588 with useRect=True."""
        return max(s.p for s in self.splines)

    def needsDG(self):
        return any(s.isDiscontinuous() for s in self.splines)

    def getNodesAndEvals(self, xi):
        """BSplines.py:450-503, same nesting order."""
        sp = self.splines
        spans = [s.getKnotSpan(xi[d]) for d, s in enumerate(sp)]
        nodes = [s.getNodes(xi[d]) for d, s in enumerate(sp)]
        ders = [s.basisFuncs(spans[d], xi[d]) for d, s in enumerate(sp)]
        out = []
        if self.nvar == 1:
            for i in range(len(nodes[0])):
                out.append([nodes[0][i], ders[0][i]])
        elif self.nvar == 2:
            for i in range(len(nodes[0])):
                for j in range(len(nodes[1])):
                    out.append([ij2dof(nodes[0][i], nodes[1][j], sp[0].ncp),
                                ders[0][i] * ders[1][j]])
        else:
            for i in range(len(nodes[0])):
                for j in range(len(nodes[1])):
                    for k in range(len(nodes[2])):
                        out.append([ijk2dof(nodes[0][i], nodes[1][j], nodes[2][k],
                                            sp[0].ncp, sp[1].ncp),
                                    ders[0][i] * ders[1][j] * ders[2][k]])
        return out

    def getSideDofs(self, direction, side, nLayers=1):
        """BSplines.py:599-649."""
        offsetSign = 1 - 2 * side
        out = []
        for absOffset in range(nLayers):
            offset = absOffset * offsetSign
            i = 0 if side == 0 else self.splines[direction].ncp - 1
            i += offset
            M = self.splines[0].ncp
            if self.nvar == 1:
                out.append(i)
                continue
            N = self.splines[1].ncp
            if self.nvar == 2:
                if direction == 0:
                    out += [ij2dof(i, j, M) for j in range(N)]
                else:
                    out += [ij2dof(j, i, M) for j in range(M)]
                continue
            O = self.splines[2].ncp
            if direction == 0:
                out += [ijk2dof(i, j, k, M, N) for j in range(N) for k in range(O)]
            elif direction == 1:
                out += [ijk2dof(j, i, k, M, N) for j in range(M) for k in range(O)]
            else:
                out += [ijk2dof(j, k, i, M, N) for j in range(M) for k in range(N)]
        return out


def explicit_control_net(tspline, extraDim=0):
    """Homogeneous control net of ExplicitBSplineControlMesh
    (BSplines.py:910-963): Greville abscissae, zero extra dims, weight 1.
    Returns [ncp, nsd+1]."""
    nvar = tspline.nvar
    nsd = nvar + extraDim
    ncps = [s.ncp for s in tspline.splines]
    grev = [np.array([s.greville(i) for i in range(s.ncp)]) for s in tspline.splines]
    P = np.zeros((tspline.ncp, nsd + 1))
    idx = np.arange(tspline.ncp)
    i0 = idx % ncps[0]
    P[:, 0] = grev[0][i0]
    if nvar >= 2:
        i1 = (idx // ncps[0]) % ncps[1]
        P[:, 1] = grev[1][i1]
    if nvar == 3:
        i2 = idx // (ncps[0] * ncps[1])
        P[:, 2] = grev[2][i2]
    P[:, nsd] = 1.0
    return P
