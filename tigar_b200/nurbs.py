"""
NURBS geometry for the hot path.

``NURBSControlMesh`` mirrors the reference's ``tIGAr/NURBS.py:13-77``: it takes
an object with igakit's ``NURBS`` attributes (``degree``, ``knots``,
``control`` = homogeneous control net, first parametric index first) and
flattens the net with the first direction fastest.  igakit itself is not
available here, so a minimal ``NURBS`` class with the two operations the
reference's demos use (``refine`` = knot insertion, ``elevate`` = degree
elevation of a single-element patch) is provided, plus the synthetic
quarter-annulus patch of BASELINE configs[3].  Host logic only (geometry
*input*); PetIGA file I/O is not built.
"""
import math

import numpy as np

from .bsplines import (AbstractControlMesh, BSpline, USE_RECT_ELEM_DEFAULT)


def _insert_knot(p, U, Pw, u):
    """Boehm's algorithm along axis 0 of Pw (homogeneous points)."""
    U = np.asarray(U, dtype=np.float64)
    n = Pw.shape[0]
    k = int(np.searchsorted(U, u, side="right")) - 1
    k = min(k, len(U) - p - 2)
    Q = np.empty((n + 1,) + Pw.shape[1:])
    for i in range(n + 1):
        if i <= k - p:
            Q[i] = Pw[i]
        elif i >= k + 1:
            Q[i] = Pw[i - 1]
        else:
            a = (u - U[i]) / (U[i + p] - U[i])
            Q[i] = a * Pw[i] + (1.0 - a) * Pw[i - 1]
    return np.insert(U, k + 1, u), Q


def _elevate_bezier(p, Pw):
    """Degree elevation p -> p+1 of a single Bezier segment along axis 0."""
    n = Pw.shape[0]
    assert n == p + 1
    Q = np.empty((n + 1,) + Pw.shape[1:])
    Q[0], Q[n] = Pw[0], Pw[n - 1]
    for i in range(1, n):
        a = i / float(p + 1)
        Q[i] = a * Pw[i - 1] + (1.0 - a) * Pw[i]
    return Q


class NURBS(object):
    """Minimal stand-in for ``igakit.nurbs.NURBS``: ``NURBS(knots, control)``
    with ``control`` of shape (n0[,n1[,n2]], dim) (weights = 1) or
    (..., dim+1) homogeneous when ``homogeneous=True``."""

    def __init__(self, knots, control, weights=None, homogeneous=False):
        self.knots = [np.array(k, dtype=np.float64) for k in knots]
        c = np.array(control, dtype=np.float64)
        nvar = len(self.knots)
        if c.ndim != nvar + 1:
            raise ValueError("control net rank does not match the number of knot vectors")
        if not homogeneous:
            w = np.ones(c.shape[:-1]) if weights is None else np.array(weights, dtype=np.float64)
            c = np.concatenate([c * w[..., None], w[..., None]], axis=-1)
        self.control = c
        self.degree = [len(k) - c.shape[d] - 1 for d, k in enumerate(self.knots)]

    @property
    def dim(self):
        return len(self.knots)

    def refine(self, axis, new_knots):
        c = np.moveaxis(self.control, axis, 0)
        U = self.knots[axis]
        for u in np.sort(np.asarray(new_knots, dtype=np.float64)):
            U, c = _insert_knot(self.degree[axis], U, c, float(u))
        self.knots[axis] = U
        self.control = np.moveaxis(c, 0, axis)
        return self

    def elevate(self, axis, times=1):
        U = self.knots[axis]
        p = self.degree[axis]
        if len(np.unique(U)) != 2:
            raise NotImplementedError("elevate() is implemented for single-element patches")
        c = np.moveaxis(self.control, axis, 0)
        for _ in range(times):
            c = _elevate_bezier(p, c)
            p += 1
        self.knots[axis] = np.array([U[0]] * (p + 1) + [U[-1]] * (p + 1))
        self.degree[axis] = p
        self.control = np.moveaxis(c, 0, axis)
        return self


class NURBSControlMesh(AbstractControlMesh):
    """NURBS.py:13-77 (igakit object in, homogeneous net flattened i-fastest)."""

    def __init__(self, fname, useRect=USE_RECT_ELEM_DEFAULT, overRefine=0):
        if isinstance(fname, str):
            raise NotImplementedError("PetIGA files need igakit; pass a NURBS object instead")
        nrb = fname
        self.scalarSpline = BSpline(list(nrb.degree), [list(k) for k in nrb.knots], useRect,
                                    overRefine)
        c = np.asarray(nrb.control, dtype=np.float64)
        nvar = len(nrb.degree)
        # first parametric index fastest (ij2dof / ijk2dof, BSplines.py:354-358)
        self.bnet = c.transpose(tuple(range(nvar - 1, -1, -1)) + (nvar,)).reshape(-1, c.shape[-1])

    def getScalarSpline(self):
        return self.scalarSpline

    def getHomogeneousCoordinate(self, node, direction):
        return self.bnet[node, direction]

    def getNsd(self):
        return self.bnet.shape[1] - 1

    def controlNet(self):
        return self.bnet


def _curve_1d(knots, cw, p_target, new_knots):
    """Elevate a single-segment homogeneous curve to p_target, then insert knots."""
    n = NURBS([knots], cw, homogeneous=True)
    if n.degree[0] < p_target:
        n.elevate(0, p_target - n.degree[0])
    if len(new_knots):
        n.refine(0, new_knots)
    return n.knots[0], n.control


def quarter_annulus(p, nel, dim=2, r_in=1.0, r_out=2.0, height=1.0):
    """Cubic (degree p >= 2) NURBS quarter annulus r in [r_in, r_out], theta in
    [0, pi/2] (x [0, height] in 3-D): exact quadratic-rational arc, degree
    elevated, uniformly h-refined to ``nel`` elements per direction (BASELINE
    configs[3]).  Built as a tensor product of refined 1-D curves so that large
    patches cost O(n) per direction.  Parametric directions: (radial, angular
    [, axial])."""
    if isinstance(nel, int):
        nel = [nel] * dim
    s = 1.0 / math.sqrt(2.0)
    # unit quarter circle, homogeneous (wx, wy, w)
    arc = np.array([[1.0, 0.0, 1.0], [s, s, s], [0.0, 1.0, 1.0]])
    lin = np.array([[0.0, 1.0], [1.0, 1.0]])            # t in [0,1], homogeneous (t, 1)

    def inner(n):
        return np.arange(1, n) / float(n)
    kr, cr = _curve_1d([0, 0, 1, 1], lin, p, inner(nel[0]))
    ka, ca = _curve_1d([0, 0, 0, 1, 1, 1], arc, p, inner(nel[1]))
    rad = r_in + (r_out - r_in) * cr[:, 0] / cr[:, 1]
    if dim == 2:
        net = np.zeros((len(rad), ca.shape[0], 3))
        net[..., 0] = rad[:, None] * ca[None, :, 0]
        net[..., 1] = rad[:, None] * ca[None, :, 1]
        net[..., 2] = ca[None, :, 2]
        return NURBS([kr, ka], net, homogeneous=True)
    kz, cz = _curve_1d([0, 0, 1, 1], lin, p, inner(nel[2]))
    z = height * cz[:, 0] / cz[:, 1]
    net = np.zeros((len(rad), ca.shape[0], len(z), 4))
    net[..., 0] = rad[:, None, None] * ca[None, :, None, 0]
    net[..., 1] = rad[:, None, None] * ca[None, :, None, 1]
    net[..., 2] = z[None, None, :] * ca[None, :, None, 2]
    net[..., 3] = ca[None, :, None, 2]
    return NURBS([kr, ka, kz], net, homogeneous=True)


def cylindrical_roof(p, nel, R=25.0, L=50.0, half_angle_deg=40.0):
    """Cylindrical roof of BASELINE configs[4] (Scordelis-Lo data: R = 25, L = 50, 80
    degree arc; textbook benchmark, not in the reference): axis along y, crown at
    (0, *, R).  Exact quadratic-rational arc, degree elevated to ``p`` and uniformly
    h-refined to ``nel`` elements per direction.  Parametric directions: (angular,
    axial); control points homogeneous (wx, wy, wz, w) as igakit stores them
    (NURBS.py:43-77)."""
    if isinstance(nel, int):
        nel = [nel, nel]
    phi = math.radians(half_angle_deg)
    s, c = math.sin(phi), math.cos(phi)
    # arc in the (x, z) plane, homogeneous (wx, wz, w); middle weight cos(phi)
    arc = np.array([[-R * s, R * c, 1.0], [0.0, R, c], [R * s, R * c, 1.0]])
    lin = np.array([[0.0, 1.0], [1.0, 1.0]])

    def inner(n):
        return np.arange(1, n) / float(n)
    ka, ca = _curve_1d([0, 0, 0, 1, 1, 1], arc, p, inner(nel[0]))
    kl, cl = _curve_1d([0, 0, 1, 1], lin, p, inner(nel[1]))
    y = L * cl[:, 0] / cl[:, 1]
    net = np.zeros((ca.shape[0], len(y), 4))
    net[..., 0] = ca[:, None, 0]
    net[..., 1] = ca[:, None, 2] * y[None, :]
    net[..., 2] = ca[:, None, 1]
    net[..., 3] = ca[:, None, 2]
    return NURBS([ka, kl], net, homogeneous=True)
