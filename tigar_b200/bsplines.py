"""
B-spline layer of the tIGAr API, backed by the CUDA library.

Mirrors the names and semantics of the reference's ``tIGAr/BSplines.py``
(uniformKnots :14-38, BSpline1 :164-351, index helpers :354-370, BSpline
:374-649, ExplicitBSplineControlMesh :910-963).  Knot bookkeeping is host
logic; every basis-function evaluation -- the reference's native
``basisFuncsInner`` (:73-120) -- runs on the GPU through
``tg_bspline_eval_batch`` (bit-exact with the reference recurrence).
"""
import numpy as np

from . import dev
from ._lib import lib, check

DOLFIN_EPS = 3.0e-16
KNOT_NEAR_EPS = 10.0 * DOLFIN_EPS      # BSplines.py:42
USE_RECT_ELEM_DEFAULT = True           # common.py:81


def near(x, x0, eps=DOLFIN_EPS):
    return (x0 - eps <= x) and (x <= x0 + eps)


def uniformKnots(p, start, end, N, periodic=False, continuityDrop=0):
    """Open (or periodic) uniform knot vector, BSplines.py:14-38."""
    if continuityDrop >= p:
        print("ERROR: Continuity drop too high for spline degree.")
        raise SystemExit(1)
    h = (end - start) / float(N)
    reps = continuityDrop + 1
    body = [start + float(i) * h for i in range(N + 1) for _ in range(reps)]
    if periodic:
        return body
    pad = p - continuityDrop
    return [start] * pad + body + [end] * pad


def ij2dof(i, j, M):
    return j * M + i


def ijk2dof(i, j, k, M, N):
    return k * (M * N) + j * M + i


def dof2ij(dof, M):
    return (dof % M, dof // M)


def dof2ijk(dof, M, N):
    ij = dof % (M * N)
    return (ij % M, ij // M, dof // (M * N))


class DofList(list):
    """A Python list of DoF ids (what the reference's ``getSideDofs`` returns,
    BSplines.py:599-649) that is backed by the numpy arrays it was concatenated from and only
    becomes a real list of Python ints when list semantics are asked for: the 4e5 side DoFs of a
    256^3 patch cost 10 ms per step to box and 30 ms to unbox again.  ``len``, iteration, ``+``,
    ``+=`` and numpy conversion (``__array__``) work on the arrays; every other list method
    materialises first.  Any in-place edit other than ``+=`` drops the arrays."""

    MAX_CHUNKS = 64          # many tiny pieces: converting the list itself is cheaper

    def __init__(self, items=(), chunks=None, lazy=False):
        if lazy:
            list.__init__(self)
            self._chunks = list(chunks)
            self._mat = False
        else:
            list.__init__(self, items)
            self._chunks = chunks if chunks is not None else ([] if list.__len__(self) == 0
                                                              else None)
            # an empty list is trivially "not yet materialised": += of arrays stays lazy
            self._mat = list.__len__(self) > 0
        self._n = sum(c.size for c in self._chunks) if self._chunks is not None else 0

    @staticmethod
    def from_array(a):
        a = np.ascontiguousarray(a, dtype=np.int64).ravel()
        return DofList(chunks=[a], lazy=True)

    def _materialise(self):
        if not self._mat:
            items = self.asarray().tolist()
            self._mat = True
            list.extend(self, items)
        return self

    def _valid(self):
        return self._chunks is not None and (not self._mat or self._n == list.__len__(self))

    def asarray(self):
        """int64 array of the entries (duplicates and order kept)."""
        if self._valid():
            ch = self._chunks
            if not ch:
                return np.zeros(0, dtype=np.int64)
            return ch[0] if len(ch) == 1 else np.concatenate(ch)
        return np.array(list(list.__iter__(self)), dtype=np.int64)

    def __array__(self, dtype=None, copy=None):
        a = self.asarray()
        return a if dtype is None else a.astype(dtype)

    def __len__(self):
        return list.__len__(self) if self._mat else self._n

    def __bool__(self):
        return len(self) > 0

    def __iter__(self):
        if self._mat:
            return list.__iter__(self)
        return iter(self.asarray().tolist())

    def __iadd__(self, other):
        lazy_other = isinstance(other, DofList) and other._valid()
        if not self._mat and self._valid() and len(self._chunks) < self.MAX_CHUNKS:
            new = other._chunks if lazy_other else [np.array(list(other), dtype=np.int64).ravel()]
            self._chunks = self._chunks + list(new)
            self._n += sum(c.size for c in new)
            return self
        ok = self._valid() and len(self._chunks) < self.MAX_CHUNKS
        self._materialise()
        list.__iadd__(self, other)
        if ok:
            new = other._chunks if lazy_other else [np.array(list(other), dtype=np.int64).ravel()]
            self._chunks = self._chunks + list(new)
            self._n += sum(c.size for c in new)
        else:
            self._chunks = None
        return self

    def __add__(self, other):
        if not self._mat and self._valid():
            out = DofList(chunks=self._chunks, lazy=True)
        else:
            out = DofList(list(self), None if self._chunks is None else list(self._chunks))
        out += other
        return out

    def _read(name):                                  # noqa: N805
        def f(self, *a, **k):
            self._materialise()
            return getattr(list, name)(self, *a, **k)
        f.__name__ = name
        return f

    def _edit(name):                                  # noqa: N805
        def f(self, *a, **k):
            self._materialise()
            self._chunks = None
            return getattr(list, name)(self, *a, **k)
        f.__name__ = name
        return f

    for _n in ("__getitem__", "__contains__", "__eq__", "__ne__", "__repr__", "__str__", "index",
               "count", "copy", "__reversed__", "__mul__", "__rmul__", "__lt__", "__le__",
               "__gt__", "__ge__"):
        locals()[_n] = _read(_n)
    for _n in ("append", "extend", "insert", "pop", "remove", "sort", "reverse", "clear",
               "__setitem__", "__delitem__", "__imul__"):
        locals()[_n] = _edit(_n)
    del _n, _edit, _read
    __hash__ = None


def basisFuncsInner(ghostKnots, nGhost, u, pl, i, ndu, left, right, ders):
    """The reference's native routine (BSplines.py:73-120, 135-145) on the device: the ``pl+1``
    non-zero basis functions at ``u`` for index ``i = span+1`` are written into ``ders``
    (a numpy array, as in the reference).  ``ghostKnots``: numpy array or device tensor.
    ``ndu/left/right`` are the reference's scratch arrays; they are not touched (the reference
    passes ``ndu.flatten()``, a copy, so callers never see them filled either).  ``u`` and ``i``
    may be arrays of equal length (batched; ``ders`` then has shape (n, pl+1))."""
    g = ghostKnots if hasattr(ghostKnots, "data_ptr") else \
        dev.from_np(np.ascontiguousarray(ghostKnots, dtype=np.float64))
    uu = np.atleast_1d(np.asarray(u, dtype=np.float64)).ravel()
    ii = np.atleast_1d(np.asarray(i, dtype=np.int32)).ravel()
    if len(uu) != len(ii):
        raise ValueError("basisFuncsInner: u and i differ in length")
    if len(ii) and (ii.min() - pl + nGhost < 0 or ii.max() + pl - 1 + nGhost >= g.numel()):
        raise IndexError("basisFuncsInner: index i reaches outside the ghost knots")
    out = dev.empty(len(uu) * (pl + 1))
    du, di = dev.from_np(uu), dev.from_np(ii)      # (named: both must outlive the launch)
    check(lib.tg_basis_funcs_inner(dev.ptr(g), int(nGhost), int(pl), dev.ptr(du), dev.ptr(di),
                                   len(uu), dev.ptr(out), dev.stream()))
    np.asarray(ders).reshape(-1)[:] = dev.to_np(out)


class BSpline1(object):
    """Univariate B-spline (BSplines.py:164-351)."""

    def __init__(self, p, knots):
        self.p = int(p)
        self.knots = np.array(knots, dtype=np.float64)
        k = self.knots
        # runs of coincident knots (near() with KNOT_NEAR_EPS, :183-193, :235-243)
        new = np.ones(len(k), dtype=bool)
        new[1:] = ~((k[:-1] - KNOT_NEAR_EPS <= k[1:]) & (k[1:] <= k[:-1] + KNOT_NEAR_EPS))
        starts = np.flatnonzero(new)
        self.uniqueKnots = k[starts].copy()
        self.multiplicities = np.diff(np.append(starts, len(k))).astype(np.int32)
        self.nel = len(starts) - 1
        self.ncp = self.computeNcp()
        self.nGhost = self.p + 1
        self.ghostKnots = self.computeGhostKnots()
        self._dev = None

    # -- bookkeeping ---------------------------------------------------------
    def computeGhostKnots(self):
        n = len(self.knots)
        return np.array([self.getKnot(i) for i in range(-self.nGhost, n + self.nGhost)])

    def normalizeKnotVector(self):
        L = self.knots[-1] - self.knots[0]
        self.knots = (self.knots - self.knots[0]) / L
        self.uniqueKnots = (self.uniqueKnots - self.uniqueKnots[0]) / L
        self.ghostKnots = self.computeGhostKnots()
        self._dev = None

    def isDiscontinuous(self):
        return bool(np.any(self.multiplicities[1:-1] > self.p))

    def computeNel(self):
        return self.nel

    def getKnot(self, i):
        """Knot with possibly out-of-range index: ghosts mirror the other end
        of the vector (BSplines.py:245-260)."""
        k = self.knots
        n = len(k)
        if i < 0:
            return k[0] - (k[-1] - k[n - int(self.multiplicities[-1]) + i])
        if i >= n:
            return k[-1] + (k[i - n + int(self.multiplicities[0])] - k[0])
        return k[i]

    def greville(self, i):
        acc = 0.0
        for j in range(i, i + self.p):
            acc += self.getKnot(j + 1)
        return acc / float(self.p)

    def grevilleAll(self):
        return np.array([self.greville(i) for i in range(self.ncp)])

    def computeNcp(self):
        return len(self.knots) - int(self.multiplicities[0])

    def getNcp(self):
        return self.ncp

    def elementSpans(self):
        """Knot-span index of every non-degenerate element."""
        return (np.cumsum(self.multiplicities[:-1]) - 1).astype(np.int32)

    # -- device evaluation ---------------------------------------------------
    def deviceKnots(self):
        if self._dev is None:
            self._dev = (dev.from_np(self.knots), dev.from_np(self.ghostKnots))
        return self._dev

    def evalBatch(self, u):
        """(span[n], nodes[n,p+1], vals[n,p+1]) for parameters ``u`` -- device
        tensors; getKnotSpan + getNodes + basisFuncs of BSplines.py:285-351."""
        dk, dg = self.deviceKnots()
        if not hasattr(u, "data_ptr"):
            u = dev.from_np(np.asarray(u, dtype=np.float64).ravel())
        n = u.numel()
        span = dev.empty(n, dev.I32)
        nodes = dev.empty(n * (self.p + 1), dev.I32)
        vals = dev.empty(n * (self.p + 1))
        check(lib.tg_bspline_eval_batch(
            dev.ptr(dk), len(self.knots), dev.ptr(dg), self.nGhost, self.p, self.ncp,
            int(self.multiplicities[0]), int(self.multiplicities[-1]), dev.ptr(u), n,
            dev.ptr(span), dev.ptr(nodes), dev.ptr(vals), dev.stream()))
        return span, nodes.view(n, self.p + 1), vals.view(n, self.p + 1)

    # scalar API of the reference (one-point launches; use evalBatch for bulk)
    def getKnotSpan(self, u):
        return int(self.evalBatch([u])[0][0].item())

    def getNodes(self, u):
        return [int(i) for i in dev.to_np(self.evalBatch([u])[1][0])]

    def basisFuncs(self, knotSpan, u):
        """BSplines.py:321-351: the caller's span is used as given (not searched again)."""
        ders = np.zeros(self.p + 1)
        basisFuncsInner(self.deviceKnots()[1], self.nGhost, u, self.p, int(knotSpan) + 1,
                        None, None, None, ders)
        return ders


class AbstractScalarBasis(object):
    """common.py:1673-1759."""

    def getNodesAndEvals(self, xi):
        raise NotImplementedError

    def getNcp(self):
        raise NotImplementedError

    def generateMesh(self, comm=None):
        raise NotImplementedError

    def getDegree(self):
        raise NotImplementedError

    def needsDG(self):
        return True

    def useRectangularElements(self):
        return False

    def getPrealloc(self):
        return 500


class AbstractControlMesh(object):
    """common.py:1762-1791."""

    def getHomogeneousCoordinate(self, node, direction):
        raise NotImplementedError

    def getScalarSpline(self):
        raise NotImplementedError

    def getNsd(self):
        raise NotImplementedError


class TensorMesh(object):
    """One rectangular cell per non-degenerate knot span (what
    BSpline.generateMesh builds with DOLFIN, BSplines.py:505-569)."""

    def __init__(self, uniqueKnots):
        self.uniqueKnots = [np.asarray(u, dtype=np.float64) for u in uniqueKnots]
        self.dim = len(self.uniqueKnots)
        self.nel = [len(u) - 1 for u in self.uniqueKnots]

    def num_cells(self):
        return int(np.prod(self.nel))

    def coordinates(self):
        g = np.meshgrid(*self.uniqueKnots, indexing="ij")
        return np.stack([a.ravel(order="F") for a in g], axis=1)


class BSpline(AbstractScalarBasis):
    """Uni-, bi- or tri-variate tensor-product B-spline (BSplines.py:374-649)."""

    def __init__(self, degrees, kvecs, useRect=USE_RECT_ELEM_DEFAULT, overRefine=0):
        self.nvar = len(degrees)
        if self.nvar > 3 or self.nvar < 1:
            print("ERROR: Unsupported parametric dimension.")
            raise SystemExit(1)
        if not useRect or overRefine:
            raise NotImplementedError(
                "tigar_b200 extracts to rectangular (quad/hex) elements only")
        self.splines = [BSpline1(degrees[i], kvecs[i]) for i in range(self.nvar)]
        self.useRect = useRect
        self.overRefine = overRefine
        self.ncp = self.computeNcp()
        self.nel = self.computeNel()

    def normalizeKnotVectors(self):
        for s in self.splines:
            s.normalizeKnotVector()

    def needsDG(self):
        return any(s.isDiscontinuous() for s in self.splines)

    def useRectangularElements(self):
        return self.useRect

    def getPrealloc(self):
        n = 1
        for s in self.splines:
            n *= s.p + 1
        return n

    def getNodesAndEvals(self, xi):
        """[[globalIndex, value], ...] in the reference's nesting order (first
        direction outermost, BSplines.py:450-503); evaluated on the GPU."""
        nodes, vals = [], []
        for d, s in enumerate(self.splines):
            _, nd, vl = s.evalBatch([xi[d]])
            nodes.append([int(i) for i in dev.to_np(nd[0])])
            vals.append(dev.to_np(vl[0]))
        sp = self.splines
        out = []
        if self.nvar == 1:
            return [[nodes[0][i], float(vals[0][i])] for i in range(len(nodes[0]))]
        if self.nvar == 2:
            for i in range(len(nodes[0])):
                for j in range(len(nodes[1])):
                    out.append([ij2dof(nodes[0][i], nodes[1][j], sp[0].ncp),
                                float(vals[0][i] * vals[1][j])])
            return out
        for i in range(len(nodes[0])):
            for j in range(len(nodes[1])):
                for k in range(len(nodes[2])):
                    out.append([ijk2dof(nodes[0][i], nodes[1][j], nodes[2][k],
                                        sp[0].ncp, sp[1].ncp),
                                float(vals[0][i] * vals[1][j] * vals[2][k])])
        return out

    def generateMesh(self, comm=None):
        return TensorMesh([s.uniqueKnots for s in self.splines])

    def computeNcp(self):
        n = 1
        for s in self.splines:
            n *= s.getNcp()
        return n

    def getNcp(self):
        return self.ncp

    def getDegree(self):
        return max(s.p for s in self.splines)

    def computeNel(self):
        n = 1
        for s in self.splines:
            n *= s.nel
        return n

    def getSideDofs(self, direction, side, nLayers=1):
        """IGA DoFs of ``nLayers`` control-point layers on one face
        (BSplines.py:599-649), in the reference's order."""
        shape = [s.getNcp() for s in self.splines]
        out = []
        for layer in range(nLayers):
            i = layer if side == 0 else shape[direction] - 1 - layer
            if self.nvar == 1:
                out.append(np.array([i], dtype=np.int64))
            elif self.nvar == 2:
                M = shape[0]
                if direction == 0:
                    out.append(np.arange(shape[1], dtype=np.int64) * M + i)
                else:
                    out.append(i * M + np.arange(shape[0], dtype=np.int64))
            else:
                M, N, O = shape
                if direction == 0:      # j outer, k inner
                    jj, kk = np.arange(N, dtype=np.int64)[:, None], np.arange(O, dtype=np.int64)[None, :]
                    out.append((kk * (M * N) + jj * M + i).ravel())
                elif direction == 1:
                    jj, kk = np.arange(M, dtype=np.int64)[:, None], np.arange(O, dtype=np.int64)[None, :]
                    out.append((kk * (M * N) + i * M + jj).ravel())
                else:
                    jj, kk = np.arange(M, dtype=np.int64)[:, None], np.arange(N, dtype=np.int64)[None, :]
                    out.append((i * (M * N) + kk * M + jj).ravel())
        return DofList.from_array(out[0] if len(out) == 1 else np.concatenate(out))


class _UnitWeightColumns(list):
    """Device columns of a control net whose weights are all exactly 1."""
    unit_weights = True


class ExplicitBSplineControlMesh(AbstractControlMesh):
    """Physical == parametric space: Greville control points, unit weights
    (BSplines.py:910-963)."""

    def __init__(self, degrees, kvecs, extraDim=0, useRect=USE_RECT_ELEM_DEFAULT,
                 overRefine=0):
        self.scalarSpline = BSpline(degrees, kvecs, useRect=useRect, overRefine=overRefine)
        self.nvar = len(degrees)
        self.nsd = self.nvar + extraDim

    def getScalarSpline(self):
        return self.scalarSpline

    def getHomogeneousCoordinate(self, node, direction):
        if direction == self.nsd:
            return 1.0
        if direction >= self.nvar:
            return 0.0
        sp = self.scalarSpline.splines
        if self.nvar == 1:
            idx = node
        elif self.nvar == 2:
            idx = dof2ij(node, sp[0].getNcp())[direction]
        else:
            idx = dof2ijk(node, sp[0].getNcp(), sp[1].getNcp())[direction]
        return sp[direction].greville(idx)

    def getNsd(self):
        return self.nsd

    def controlNetDevice(self):
        """The homogeneous control net as nsd+1 device columns, generated on the device from
        the 1-D Greville abscissae (one coalesced kernel per column; the host loop of
        common.py:373-375 costs 1.4 s at 256^3)."""
        from . import dev
        from ._lib import lib, check
        sp = self.scalarSpline.splines
        n = [s.getNcp() for s in sp] + [1] * (3 - len(sp))
        ncp = int(np.prod(n))
        cols = _UnitWeightColumns()
        for d in range(self.nsd + 1):
            out = dev.empty(ncp)
            if d < self.nvar:
                g = dev.from_np(np.ascontiguousarray(sp[d].grevilleAll(), dtype=np.float64))
                check(lib.tg_tensor_column(dev.ptr(out), dev.ptr(g), n[0], n[1], n[2], d, 0.0,
                                           dev.stream()))
            else:
                check(lib.tg_tensor_column(dev.ptr(out), None, n[0], n[1], n[2], 0,
                                           1.0 if d == self.nsd else 0.0, dev.stream()))
            cols.append(out)
        return cols

    def controlNet(self):
        """Whole homogeneous control net [ncp, nsd+1] at once (bulk form of the
        per-node loop of common.py:373-375)."""
        sp = self.scalarSpline.splines
        shape = [s.getNcp() for s in sp]
        ncp = int(np.prod(shape))
        P = np.zeros((ncp, self.nsd + 1))
        idx = np.arange(ncp)
        stride = 1
        for d, s in enumerate(sp):
            P[:, d] = s.grevilleAll()[(idx // stride) % shape[d]]
            stride *= shape[d]
        P[:, self.nsd] = 1.0
        return P
