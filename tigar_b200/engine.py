"""
Host-side driver of the CUDA hot path on one tensor-product B-spline patch:

  extraction   M = M_w (x) M_v (x) M_u          (common.py:1460-1578)
  assembly     A_FE, b_FE on the Q_pf Lagrange mesh (common.py:1169,1215-1216)
  extraction   M^T A M, M^T b, zero BCs         (common.py:1142-1204)
  solve        Jacobi-CG                        (common.py:1236-1263)

and the element-fused variant  C = sum_e M_e^T K_e M_e  that never forms the
global A_FE / M (mandatory at 256^3 and beyond).

Everything numerical happens in libtigar_b200.so; this module only builds
descriptors (window ranges, multi-index lists, register programs) and
launches.  Matrices are "windowed CSR": each row stores the dense
tensor-product box of its columns, so no column index array exists (8 B/nnz).
"""
import ctypes as C
import os

import numpy as np

from . import dev
from . import symbolic as S
from . import march_tables
from ._lib import lib, check, tg_basis, tg_win, i32arr, vparr, c_vp
from .bsplines import BSpline1

IGNORE_EPS = 1e-15          # DEFAULT_BASIS_FUNC_IGNORE_EPS, common.py:56


# --------------------------------------------------------------------------
def gauss_rule01(nq):
    x, w = np.polynomial.legendre.leggauss(nq)
    return 0.5 * (x + 1.0), 0.5 * w


def lagrange_jets01(pf, t, nder):
    """d^k/dt^k of the equispaced degree-pf Lagrange basis on [0,1] at points
    t: [len(t), pf+1, nder+1].  Reference-element constants (FIAT's role)."""
    nodes = np.arange(pf + 1) / float(pf)
    out = np.zeros((len(t), pf + 1, nder + 1))
    for a in range(pf + 1):
        others = np.delete(nodes, a)
        poly = np.poly1d(others, r=True) / np.prod(nodes[a] - others)
        for k in range(nder + 1):
            out[:, a, k] = np.polyder(poly, k)(t) if k else poly(t)
    return out


class Window(object):
    """Tensor-product row windows of a sparse matrix + its device descriptor."""

    def __init__(self, nr, nc, lo, hi):
        self.dim = len(nr)
        self.nr = [int(x) for x in nr]
        self.nc = [int(x) for x in nc]
        self.lo = [np.ascontiguousarray(a, dtype=np.int32) for a in lo]
        self.hi = [np.ascontiguousarray(a, dtype=np.int32) for a in hi]
        for d in range(self.dim):
            assert len(self.lo[d]) == self.nr[d] and len(self.hi[d]) == self.nr[d]
            assert np.all(self.lo[d] >= 0) and np.all(self.hi[d] < self.nc[d])
            assert np.all(self.hi[d] >= self.lo[d])
        self.len = [(h - l + 1).astype(np.int64) for l, h in zip(self.lo, self.hi)]
        self.nrows = int(np.prod(self.nr))
        self.ncols = int(np.prod(self.nc))
        self.nnz = int(np.prod([int(l.sum()) for l in self.len]))
        self.row0 = [0] * self.dim
        self.col0 = [0] * self.dim
        self.layout = 0
        self.H = 1
        self._d = None

    def sell(self):
        """Switch to the SELL-H layout (lanes = rows; see tg_win.layout in
        include/tigar_b200.h).  Returns self."""
        assert self._d is None
        nr0 = self.nr[0]
        nchunk = (nr0 + 31) // 32
        self.H = (nr0 + nchunk - 1) // nchunk
        self.layout = 1
        w0 = int(self.len[0].max())
        self.bs0 = np.minimum(self.lo[0].astype(np.int64), self.nc[0] - w0).astype(np.int32)
        assert np.all(self.bs0 >= 0)
        return self

    def storage(self):
        """Doubles occupied by the value array in this layout."""
        if self.layout == 0:
            return self.nnz
        nchunk = (self.nr[0] + self.H - 1) // self.H
        rest = int(np.prod([int(l.sum()) for l in self.len[1:]])) if self.dim > 1 else 1
        return self.H * nchunk * int(self.len[0].max()) * rest

    def slab(self, k0, k1, c0, c1):
        """Rows [k0,k1) of the last direction, columns restricted to [c0,c1)
        (must contain every window of those rows): one rank's block of a
        row-distributed matrix."""
        L = self.dim - 1
        lo = list(self.lo)
        hi = list(self.hi)
        lo[L] = self.lo[L][k0:k1] - c0
        hi[L] = self.hi[L][k0:k1] - c0
        nr = list(self.nr)
        nc = list(self.nc)
        nr[L] = k1 - k0
        nc[L] = c1 - c0
        w = Window(nr, nc, lo, hi)
        w.row0[L] = k0
        w.col0[L] = c0
        if self.layout == 1:
            w.sell()
        return w

    def transpose(self):
        lo, hi = [], []
        for d in range(self.dim):
            a, b = self.lo[d].astype(np.int64), self.hi[d].astype(np.int64)
            if np.all(np.diff(a) >= 0) and np.all(np.diff(b) >= 0):
                # monotone windows (every spline / FE pattern): column c is covered by the
                # rows from the first one with hi >= c to the last one with lo <= c
                c = np.arange(self.nc[d], dtype=np.int64)
                l = np.searchsorted(b, c, side="left")
                h = np.searchsorted(a, c, side="right") - 1
                empty = l > h
                l[empty] = np.iinfo(np.int32).max
                h[empty] = -1
            else:
                l = np.full(self.nc[d], np.iinfo(np.int32).max, dtype=np.int64)
                h = np.full(self.nc[d], -1, dtype=np.int64)
                for rr in range(self.nr[d]):
                    l[a[rr]:b[rr] + 1] = np.minimum(l[a[rr]:b[rr] + 1], rr)
                    h[a[rr]:b[rr] + 1] = np.maximum(h[a[rr]:b[rr] + 1], rr)
            lo.append(l)
            hi.append(h)
        return Window(self.nc, self.nr, lo, hi)

    def compose(self, other):
        """Window of the product (self) * (other)."""
        lo, hi = [], []
        for d in range(self.dim):
            assert self.nc[d] == other.nr[d]
            ol, oh = other.lo[d], other.hi[d]
            if np.all(np.diff(ol) >= 0) and np.all(np.diff(oh) >= 0):
                l, h = ol[self.lo[d]], oh[self.hi[d]]          # monotone: ends of the range
            else:
                l = np.array([ol[a:b + 1].min() for a, b in zip(self.lo[d], self.hi[d])])
                h = np.array([oh[a:b + 1].max() for a, b in zip(self.lo[d], self.hi[d])])
            lo.append(l)
            hi.append(h)
        return Window(self.nr, other.nc, lo, hi)

    # device descriptor (lazily uploaded)
    def dev(self):
        if self._d is None:
            w = tg_win()
            w.dim = self.dim
            keep = []
            S_ptrs = []
            for d in range(3):
                if d < self.dim:
                    w.nr[d], w.nc[d] = self.nr[d], self.nc[d]
                    w.row0[d], w.col0[d] = self.row0[d], self.col0[d]
                    tl, th = dev.from_np(self.lo[d]), dev.from_np(self.hi[d])
                    ts = dev.from_np(np.concatenate([[0], np.cumsum(self.len[d])]).astype(np.int64))
                    keep += [tl, th, ts]
                    w.lo[d], w.hi[d] = dev.ptr(tl), dev.ptr(th)
                    w.S[d] = dev.ptr(ts)
                    S_ptrs.append(dev.ptr(ts))
                else:
                    w.nr[d], w.nc[d] = 1, 1
                    w.lo[d], w.hi[d] = None, None
                    w.S[d] = None
                    S_ptrs.append(0)
            rowptr = dev.empty(self.nrows + 1, dev.I64)
            w.rowptr = dev.ptr(rowptr)
            w.w0max = int(self.len[0].max())
            w.maxrow = int(np.prod([int(l.max()) for l in self.len]))
            w.layout, w.H = self.layout, self.H
            if self.layout == 1:
                tb = dev.from_np(self.bs0)
                keep.append(tb)
                w.bs0 = dev.ptr(tb)
            check(lib.tg_win_rowptr(C.byref(w), (c_vp * 3)(*S_ptrs), dev.ptr(rowptr),
                                    dev.stream()))
            keep.append(rowptr)
            self._d = (w, keep, rowptr)
        return self._d[0]

    def ref(self):
        return C.byref(self.dev())

    def rowptr(self):
        self.dev()
        return self._d[2]

    def columns(self):
        """Explicit CSR column array (only for export / generic CSR kernels)."""
        cols = dev.empty(self.nnz, dev.I32)
        check(lib.tg_win_fill_cols(self.ref(), dev.ptr(cols), dev.stream()))
        return cols


class WinMatrix(object):
    """Values on a Window.  ``to_scipy`` exports a standard CSR for checking."""

    def __init__(self, window, vals=None):
        self.window = window
        self.vals = dev.zeros(window.storage()) if vals is None else vals

    @staticmethod
    def from_csr_values(window, csr_vals):
        """Values given in exact row-major CSR order -> the window's layout."""
        if window.layout == 0:
            return WinMatrix(window, csr_vals)
        m = WinMatrix(window)
        check(lib.tg_win_import_vals(window.ref(), dev.ptr(csr_vals), dev.ptr(m.vals),
                                     dev.stream()))
        return m

    def csr_values(self):
        w = self.window
        if w.layout == 0:
            return self.vals
        out = dev.empty(w.nnz)
        check(lib.tg_win_export_vals(w.ref(), dev.ptr(self.vals), dev.ptr(out), dev.stream()))
        return out

    @property
    def shape(self):
        return (self.window.nrows, self.window.ncols)

    def matvec(self, x, y=None):
        y = dev.empty(self.window.nrows) if y is None else y
        check(lib.tg_win_spmv(self.window.ref(), dev.ptr(self.vals), dev.ptr(x), dev.ptr(y),
                              dev.stream()))
        return y

    def to_scipy(self, drop_eps=None):
        import scipy.sparse as sp
        w = self.window
        rp = dev.to_np(w.rowptr())
        cols = dev.to_np(w.columns())
        vals = dev.to_np(self.csr_values())
        A = sp.csr_matrix((vals, cols, rp), shape=self.shape)
        if drop_eps is not None:
            A.data[np.abs(A.data) <= drop_eps] = 0.0
            A.eliminate_zeros()
        A.sort_indices()
        return A

    def copy(self):
        return WinMatrix(self.window, self.vals.clone())


# --------------------------------------------------------------------------
class Dir1D(object):
    """One parametric direction: knot data, 1-D extraction rows, per-element
    tables at the Gauss points."""

    def __init__(self, spline1, pf, nq, eps=IGNORE_EPS):
        s = spline1
        self.s = s
        self.p, self.pf, self.nq = s.p, int(pf), int(nq)
        self.nel, self.ncp = s.nel, s.ncp
        self.nfe = s.nel * self.pf + 1
        if s.multiplicities[0] != s.p + 1 or s.multiplicities[-1] != s.p + 1:
            raise NotImplementedError(
                "tigar_b200 tensor-product fast path needs open (non-periodic) knot vectors")
        self.d_uk = dev.from_np(s.uniqueKnots)
        self.d_espan = dev.from_np(s.elementSpans())
        self.tables = {}
        # cell -> first basis function, and the 1-D window of the IGA system matrix (pairs of
        # functions that share a cell): needs no FE space
        self.first = s.elementSpans().astype(np.int64) - s.p
        lo = np.full(self.ncp, np.iinfo(np.int64).max, dtype=np.int64)
        hi = np.full(self.ncp, -1, dtype=np.int64)
        for a in range(s.p + 1):
            np.minimum.at(lo, self.first + a, self.first)
            np.maximum.at(hi, self.first + a, self.first + s.p)
        self.c_lo, self.c_hi = lo, hi
        self.discontinuous = bool(s.isDiscontinuous())
        if self.discontinuous:
            # a discontinuous spline needs a DG extraction space (BSplines.py:419-427); the
            # element-fused and matrix-free paths never form M or A_FE, so they work as they are
            self.m_first = self.m_vals = self.x_fe = None
            self.m_lo = self.m_hi = None
            return
        # 1-D extraction rows at the FE nodes (reference per-node evaluation)
        x = dev.empty(self.nfe)
        check(lib.tg_fe_nodes_1d(dev.ptr(self.d_uk), self.nel, self.pf, dev.ptr(x), dev.stream()))
        self.x_fe = x
        span, nodes, vals = s.evalBatch(x)
        self.m_first = (span - self.p).contiguous()          # int32 [nfe]
        self.m_vals = vals.contiguous()                      # [nfe, p+1]
        hv = dev.to_np(vals)
        hf = dev.to_np(self.m_first).astype(np.int64)
        keep = np.abs(hv) > eps
        first_kept = keep.argmax(axis=1)
        last_kept = self.p - keep[:, ::-1].argmax(axis=1)
        width = keep.sum(axis=1)
        if not np.all(width == last_kept - first_kept + 1):
            raise RuntimeError("non-contiguous 1-D extraction row")
        self.m_lo = hf + first_kept
        self.m_hi = hf + last_kept

    def table(self, nder):
        """Device tables tabulated up to derivative order ``nder``."""
        nder = int(nder)
        if nder not in self.tables:
            s = self.s
            tq, gw = gauss_rule01(self.nq)
            lag = lagrange_jets01(self.pf, tq, nder)
            dk, dg = s.deviceKnots()
            np1, nf, nd = self.p + 1, self.pf + 1, nder + 1
            t = dict(
                Me=dev.empty(self.nel * nf * np1),
                tabN=dev.empty(self.nel * self.nq * np1 * nd),
                idxN=dev.empty(self.nel * np1, dev.I32),
                tabL=dev.empty(self.nel * self.nq * nf * nd),
                idxL=dev.empty(self.nel * nf, dev.I32),
                wq=dev.empty(self.nel * self.nq),
                xq=dev.empty(self.nel * self.nq),
                lag=dev.from_np(lag), tq=dev.from_np(tq), gw=dev.from_np(gw))
            check(lib.tg_tabulate_1d(
                dev.ptr(dg), s.nGhost, self.p, self.ncp, dev.ptr(self.d_uk),
                dev.ptr(self.d_espan), self.nel, self.pf, self.nq, nder,
                dev.ptr(t["lag"]), dev.ptr(t["tq"]), dev.ptr(t["gw"]),
                dev.ptr(t["Me"]), dev.ptr(t["tabN"]), dev.ptr(t["idxN"]),
                dev.ptr(t["tabL"]), dev.ptr(t["idxL"]), dev.ptr(t["wq"]), dev.ptr(t["xq"]),
                dev.stream()))
            self.tables[nder] = t
        return self.tables[nder]


class Basis(object):
    """ctypes tg_basis + the tensors it points to."""

    def __init__(self, dirs, kind, nder):
        b = tg_basis()
        b.dim = len(dirs)
        b.nder = nder
        self.keep = []
        for d in range(3):
            if d < len(dirs):
                D = dirs[d]
                t = D.table(nder)
                self.keep.append(t)
                b.nel[d], b.nq[d] = D.nel, D.nq
                if kind == "fe":
                    b.n[d], b.nloc[d] = D.nfe, D.pf + 1
                    b.tab[d], b.idx[d] = dev.ptr(t["tabL"]), dev.ptr(t["idxL"])
                else:
                    b.n[d], b.nloc[d] = D.ncp, D.p + 1
                    b.tab[d], b.idx[d] = dev.ptr(t["tabN"]), dev.ptr(t["idxN"])
                b.wq[d], b.xq[d] = dev.ptr(t["wq"]), dev.ptr(t["xq"])
            else:
                b.n[d] = b.nel[d] = b.nloc[d] = b.nq[d] = 1
        self.c = b
        self.kind = kind
        self.nder = nder
        self.nloc = [b.nloc[d] for d in range(b.dim)]
        self.ntot = int(np.prod([b.n[d] for d in range(b.dim)]))
        self.nqp = int(np.prod([b.nq[d] for d in range(b.dim)]))
        self.nen = int(np.prod(self.nloc))

    def ref(self):
        return C.byref(self.c)


def pad3(alpha):
    a = tuple(int(x) for x in alpha)
    return a + (0,) * (3 - len(a))


class SlabVector(object):
    """Planes [p0, p1) of the last direction of an IGA vector whose other planes this rank never
    reads (a coefficient function on a slab partition: the control net of a NURBS patch).  The
    Gauss-point kernels index coefficient vectors GLOBALLY, so they are given the virtual base
    ``ptr - 8*p0*plane``; only the rank's own cell layers are launched and their basis
    functions lie inside the stored planes (``TensorPatch.coef_planes``)."""

    def __init__(self, t, p0, p1, plane):
        assert t.numel() == (p1 - p0) * plane
        self.t, self.p0, self.p1, self.plane = t, int(p0), int(p1), int(plane)

    def numel(self):
        return self.t.numel()


def coef_ptr(x):
    if isinstance(x, SlabVector):
        return dev.ptr(x.t) - 8 * x.p0 * x.plane
    return dev.ptr(x)


class TensorPatch(object):
    """A tensor-product B-spline patch and its Q_pf Lagrange background mesh."""

    def __init__(self, degrees, kvecs, quadDeg=None, eps=IGNORE_EPS, splines=None, part=None):
        """``part=(rank, size)``: this process owns one slab of IGA planes of
        the last parametric direction (fused path only); see multigpu.py."""
        dev.require_cuda()
        self.dim = len(degrees)
        self.splines = splines or [BSpline1(p, k) for p, k in zip(degrees, kvecs)]
        self.degrees = [s.p for s in self.splines]
        self.pf = max(self.degrees)                     # BSplines.py:580-588 (useRect)
        self.quadDeg = 2 * self.pf if quadDeg is None else int(quadDeg)
        self.nq = self.quadDeg // 2 + 1                 # Gauss-Legendre points / direction
        self.eps = eps
        self.dirs = []
        for s in self.splines:          # directions with the same knot vector share their tables
            twin = next((D for D in self.dirs if D.s.p == s.p and len(D.s.knots) == len(s.knots)
                         and np.array_equal(D.s.knots, s.knots)), None)
            self.dirs.append(twin if twin is not None else Dir1D(s, self.pf, self.nq, eps))
        self.nel = [D.nel for D in self.dirs]
        self.ncp = [D.ncp for D in self.dirs]
        self.nfe = [D.nfe for D in self.dirs]
        self.ncells = int(np.prod(self.nel))
        self.n_iga = int(np.prod(self.ncp))
        self.n_fe = int(np.prod(self.nfe))
        self._bases = {}
        self._win = {}
        self.launches = 0
        self.part = part
        self.slab_lo, self.slab_hi = 0, self.nel[-1]
        if part is not None and part[1] > 1:
            from .multigpu import plane_partition
            D = self.dirs[-1]
            plane = self.n_iga // D.ncp
            first = D.s.elementSpans().astype(np.int64) - D.p      # first function of a cell
            self.pp = plane_partition(D.ncp, first, first + D.p, self.window_global_C_last(),
                                      part[0], part[1])
            self.slab_lo, self.slab_hi = self.pp["cells"]
            self.plane = plane
            self.n_loc = plane * (self.pp["k1"] - self.pp["k0"])
            self.n_ext = plane * (self.pp["c1"] - self.pp["c0"])
            self.xoff = plane * (self.pp["k0"] - self.pp["c0"])
        else:
            self.part = None

    def coef_planes(self):
        """Planes of the last direction that the rank's cell layers (halo layers included)
        read from a coefficient function: [first function of the first cell, last function of
        the last cell], widened to the column range of the rank's matrix rows."""
        if self.part is None:
            return 0, self.ncp[-1]
        D = self.dirs[-1]
        first = D.s.elementSpans().astype(np.int64) - D.p
        lo = min(int(first[self.slab_lo]), self.pp["c0"])
        hi = max(int(first[self.slab_hi - 1]) + D.p + 1, self.pp["c1"])
        return lo, hi

    def window_global_C_last(self):
        """(lo, hi) of the global C window in the last direction."""
        w = self._global_window("C")
        return w.lo[-1], w.hi[-1]

    # ---- descriptors -------------------------------------------------------
    def basis(self, kind, nder):
        key = (kind, int(nder))
        if key not in self._bases:
            self._bases[key] = Basis(self.dirs, kind, int(nder))
        return self._bases[key]

    def window(self, name):
        """'M' FExIGA, 'MT' IGAxFE, 'A' FExFE, 'P' = A*M, 'PT', 'C' IGAxIGA.
        With a slab partition, 'C' is this rank's block of rows."""
        if name == "C" and getattr(self, "part", None) is not None:
            if "Cloc" not in self._win:
                pp = self.pp
                self._win["Cloc"] = self._system_window().slab(pp["k0"], pp["k1"], pp["c0"],
                                                               pp["c1"])
            return self._win["Cloc"]
        if name == "C":
            return self._system_window()
        return self._global_window(name)

    def _system_window(self):
        """The IGA system matrix window.  Row-major by default; TIGAR_B200_LAYOUT=1
        selects the SELL-H layout (lanes = rows).  Measured on B200 at 128^3/200^3
        cubic (tools/spmv_probe.py, profiles/r1_spmv_variants.txt): row-major
        61 % of HBM peak, SELL-29 29-46 % -- the coalesced layout loses because
        each warp keeps fewer bytes in flight; it stays available for study."""
        if "Csys" not in self._win:
            g = self._global_window("C")
            w = Window(g.nr, g.nc, g.lo, g.hi)
            if os.environ.get("TIGAR_B200_LAYOUT", "0") == "1":
                w.sell()
            self._win["Csys"] = w
        return self._win["Csys"]

    def _global_window(self, name):
        if name not in self._win:
            if name != "C" and any(D.discontinuous for D in self.dirs):
                raise NotImplementedError(
                    "discontinuous B-splines need DG extraction for the csr path (A_FE, M); "
                    "use mode='fused' or 'matfree'")
            if name == "M":
                w = Window(self.nfe, self.ncp, [D.m_lo for D in self.dirs],
                           [D.m_hi for D in self.dirs])
            elif name == "MT":
                w = self._global_window("M").transpose()
            elif name == "A":
                lo, hi = [], []
                for D in self.dirs:
                    g = np.arange(D.nfe)
                    e_lo = np.maximum((g - 1) // D.pf, 0)          # leftmost cell touching g
                    e_hi = np.minimum(g // D.pf, D.nel - 1)        # rightmost cell
                    lo.append(e_lo * D.pf)
                    hi.append((e_hi + 1) * D.pf)
                w = Window(self.nfe, self.nfe, lo, hi)
            elif name == "P":
                w = self._global_window("A").compose(self._global_window("M"))
            elif name == "PT":
                w = self._global_window("P").transpose()
            elif name == "C":
                # pairs of functions sharing a cell (= MT o A o M composed, without FE windows)
                w = Window(self.ncp, self.ncp, [D.c_lo for D in self.dirs],
                           [D.c_hi for D in self.dirs])
            else:
                raise KeyError(name)
            self._win[name] = w
        return self._win[name]

    # ---- (i) extraction ----------------------------------------------------
    def build_M(self):
        """Global extraction operator on its window (generateM)."""
        w = self.window("M")
        M = WinMatrix(w, dev.empty(w.nnz))
        firsts = vparr([dev.ptr(D.m_first) for D in self.dirs])
        vals = vparr([dev.ptr(D.m_vals) for D in self.dirs])
        check(lib.tg_m_fill(w.ref(), firsts, vals, i32arr(self.degrees), dev.ptr(M.vals),
                            dev.stream()))
        return M

    def mt_vec(self, M, b):
        """M^T b (multTranspose, common.py:97-109)."""
        from .generic import CsrMatrix
        if isinstance(M, CsrMatrix):
            return M.transpose().matvec(b)
        out = dev.empty(self.n_iga)
        check(lib.tg_mt_vec(self.window("M").ref(), self.window("MT").ref(), dev.ptr(M.vals),
                            dev.ptr(b), dev.ptr(out), dev.stream()))
        return out

    def fe_node_coords(self):
        """[n_fe, dim] FE node coordinates (tabulate_dof_coordinates)."""
        xs = [dev.to_np(D.x_fe) for D in self.dirs]
        g = np.meshgrid(*xs, indexing="ij")
        return np.stack([a.ravel(order="F") for a in g], axis=1)

    # ---- (ii) Gauss-point assembly ----------------------------------------
    def _cell_chunks(self, bytes_per_cell, budget=3 << 29):
        """Chunks of whole slabs of the last direction."""
        slab = self.ncells // self.nel[-1]
        per = max(1, int(budget // max(1, slab * bytes_per_cell)))
        k = self.slab_lo
        while k < self.slab_hi:
            n = min(per, self.slab_hi - k)
            yield k * slab, n * slab
            k += n

    def _qp_setup(self, outputs, funcs):
        """Compile ``outputs`` (symbolic Nodes) -> device program + jet table.
        funcs: {fid: device vector}."""
        prog = S.compile_program(outputs, self.dim)
        fids = sorted(set(j[0] for j in prog.jets))
        fpos = {f: i for i, f in enumerate(fids)}
        for f in fids:
            if f not in funcs:
                raise KeyError("coefficient function %r has no data in this basis" % (f,))
        jets = []
        for (f, comp, al) in prog.jets:
            jets += [fpos[f], comp] + list(pad3(al))
        nder = max([max(al) for (_, _, al) in prog.jets] + [0])
        P = dict(prog=prog, fids=fids, njets=len(prog.jets), jets=i32arr(jets),
                 coefs=vparr([coef_ptr(funcs[f]) for f in fids]),
                 keep=[funcs[f] for f in fids],
                 ncomp=i32arr([1] * len(fids)),
                 d_prog=dev.from_np(np.array(prog.prog, dtype=np.int32).reshape(-1, 4))
                 if prog.prog else dev.zeros(4, dev.I32),
                 d_consts=dev.from_np(np.array(prog.consts or [0.0], dtype=np.float64)),
                 outregs=i32arr(prog.outregs), nout=len(prog.outregs), nder=nder)
        return P

    def _setup_cached(self, cache, outputs, funcs):
        """``_qp_setup`` with an optional per-caller cache (a dict): the compiled program,
        its device copy and its JIT kernels are kept, the coefficient-function pointers are
        re-bound on every call (Functions may have been given new tensors).  ``outputs`` is
        a callable that builds the output nodes (only evaluated on a miss)."""
        if cache is None:
            return self._qp_setup(outputs(), funcs)
        P = cache.get("P")
        if P is None:
            P = self._qp_setup(outputs(), funcs)
            cache["P"] = P
        else:
            P["keep"] = [funcs[f] for f in P["fids"]]
            P["coefs"] = vparr([coef_ptr(t) for t in P["keep"]])
        return P

    def _qp_eval(self, B, P, cell0, ncells, out, gsf=None):
        """``gsf=(sstride, lc, elast0)``: store the outputs in the layout of the global
        sum-factorised assembly (generated kernels only)."""
        from . import jit
        if gsf is not None:
            key = ("gsf", B.nder)
            k = P.setdefault("jit", {}).get(key)
            if k is None:
                fpos = {f: i for i, f in enumerate(P["fids"])}
                jets = [(fpos[f], comp, pad3(al)) for (f, comp, al) in P["prog"].jets]
                nloc = B.nloc + [1] * (3 - self.dim)
                nq = [B.c.nq[d] for d in range(3)]
                k = jit.get_kernel(P["prog"], self.dim, nloc, nq, B.nder + 1, jets,
                                   len(P["fids"]), layout="gsf")
                P["jit"][key] = k
            jit.launch(k, B, [coef_ptr(t) for t in P["keep"]], cell0, ncells, out, gsf=gsf)
            return
        if jit.enabled() and len(P["fids"]) <= jit.MAXFUN:
            k = P.setdefault("jit", {}).get(B.nder)
            if k is None:
                fpos = {f: i for i, f in enumerate(P["fids"])}
                jets = [(fpos[f], comp, pad3(al)) for (f, comp, al) in P["prog"].jets]
                nloc = B.nloc + [1] * (3 - self.dim)
                nq = [B.c.nq[d] for d in range(3)]
                k = jit.get_kernel(P["prog"], self.dim, nloc, nq, B.nder + 1, jets,
                                   len(P["fids"]))
                P["jit"][B.nder] = k
            jit.launch(k, B, [coef_ptr(t) for t in P["keep"]], cell0, ncells, out)
            return
        check(lib.tg_qp_eval(B.ref(), len(P["fids"]), P["coefs"], P["ncomp"], P["njets"],
                             P["jets"], dev.ptr(P["d_prog"]), len(P["prog"].prog),
                             dev.ptr(P["d_consts"]), P["prog"].nreg, P["nout"], P["outregs"],
                             cell0, ncells, dev.ptr(out), dev.stream()))

    @staticmethod
    def jet_order(nodes):
        return S.max_order(nodes)

    @staticmethod
    def _sf_key(k):
        """Order terms so that those sharing the last-direction derivative
        orders are adjacent (the sum-factorised kernel runs its last stage once
        per such group)."""
        a, b = pad3(k[0]), pad3(k[1])
        return (a[2], b[2], a[1], b[1], a[0], b[0])

    # ---- global sum-factorised assembly (csrc/tg_gsf.cu) ---------------------------------
    def _gsf_ok(self, B, W, P):
        """The march kernels cover 2-D / 3-D patches, row-major windows, generated
        Gauss-point kernels and tables up to second derivatives."""
        from . import jit
        if os.environ.get("TIGAR_B200_GSF", "1") != "1" or self.dim not in (2, 3):
            return False
        if W is not None and W.layout != 0:
            return False
        if not jit.enabled() or len(P["fids"]) > jit.MAXFUN or B.nder > 2:
            return False
        return all(lib.tg_gsf_supported(B.nloc[d], int(B.c.nq[d])) for d in range(self.dim))

    @staticmethod
    def _dedupe(nodes):
        """(unique nodes, slot of every input node): hash-consed nodes are identical objects
        when two terms carry the same coefficient."""
        uniq, slot, seen = [], [], {}
        for n in nodes:
            k = seen.get(n.uid)
            if k is None:
                k = seen[n.uid] = len(uniq)
                uniq.append(n)
            slot.append(k)
        return uniq, slot

    @staticmethod
    def _gsf_plan(entries, dim, pair):
        """entries: [(input kind, ((al, be),)*dim)] -> per stage (nout, maxin, host plan).
        Terms whose derivative orders agree in all REMAINING directions share an output kind."""
        stages, cur = [], list(entries)
        for d in range(dim):
            outs, plan = {}, {}
            for kin, sig in cur:
                ko = outs.setdefault(sig[1:], len(outs))
                plan.setdefault(ko, []).append((kin, sig[0][0], sig[0][1] if pair else 0))
            maxin = max(len(v) for v in plan.values())
            tab = np.zeros((len(outs), 1 + 3 * maxin), dtype=np.int32)
            for ko, lst in plan.items():
                tab[ko, 0] = len(lst)
                tab[ko, 1:1 + 3 * len(lst)] = np.array(lst, dtype=np.int32).ravel()
            stages.append((len(outs), maxin, tab))
            cur = [(ko, rest) for rest, ko in outs.items()]
        assert stages[-1][0] == 1
        return stages

    def _gsf_dir(self, B, Wg, d, pair):
        """Per-direction device constants of the march: rowbase, F_d (pairs) or n_d."""
        key = ("gsf", B.kind, d, pair, id(Wg))
        hit = self._bases.get(key)
        if hit is None:
            if pair:
                S = np.concatenate([[0], np.cumsum(Wg.len[d])]).astype(np.int64)
                rb = dev.from_np((S[:-1] - Wg.lo[d].astype(np.int64)).astype(np.int64))
                hit = (rb, int(S[-1]))
            else:
                hit = (dev.zeros(1, dev.I64), int(B.c.n[d]))
            self._bases[key] = hit
        return hit

    def _gsf_perm_rows(self, W, F0):
        """Packed per-row constants of the permuted pair layout (tg_gsf_stage ``perm_rows``):
        for every row i1 of the second direction {S1[i1]*F0, len1[i1] | lo1[i1] << 32}."""
        key = ("gsf_prow", id(W), int(F0))
        hit = self._bases.get(key)
        if hit is None:
            ln = W.len[1].astype(np.int64)
            S1 = np.concatenate([[0], np.cumsum(ln)])[:-1]
            pk = np.empty((len(ln), 2), dtype=np.int64)
            pk[:, 0] = S1 * int(F0)
            pk[:, 1] = ln | (W.lo[1].astype(np.int64) << 32)
            hit = (dev.from_np(pk.ravel()), W)          # (keeps W alive: id() is the key)
            self._bases[key] = hit
        return hit[0]

    def _gsf_run(self, B, kind, P, nslots, mentries, ventries, A, b, vrow0, vnr, xvec=None):
        """One pass over the cell layers of this patch (or slab): Gauss-point kernel into the
        blocked coefficient layout, then the march stages of the matrix (``mentries``) and of
        the load vector (``ventries``).  A / b may be None.  ``xvec``: further vector jobs fed
        by the same Gauss-point pass, [(entries, out, [(tab, idx)]*dim, nd)], each with its own
        1-D tables (the matrix diagonal: ``assemble_vector_and_diag``)."""
        dim, L = self.dim, self.dim - 1
        nel = [int(B.c.nel[d]) for d in range(dim)]
        nq = [int(B.c.nq[d]) for d in range(dim)]
        nl = list(B.nloc)
        nqp = B.nqp
        nd = B.nder + 1
        Wg = self._global_window("A" if kind == "fe" else "C")
        mst = self._gsf_plan(mentries, dim, True) if A is not None else None
        vst = self._gsf_plan(ventries, dim, False) if b is not None else None
        F = [self._gsf_dir(B, Wg, d, True)[1] for d in range(dim)]
        n = [int(B.c.n[d]) for d in range(dim)]
        plane_cells = int(np.prod(nel[:L]))
        # bytes per layer of the last direction: coefficients + intermediates
        per = nslots * plane_cells * nqp
        tabs = [(B.c.tab[d], B.c.idx[d]) for d in range(dim)]
        jobs = []
        if mst is not None:
            jobs.append((1, mst, F, A, tabs, nd, "m"))
        if vst is not None:
            jobs.append((0, vst, n, b, tabs, nd, "v"))
        for i, (xe, xo, xt, xnd) in enumerate(xvec or []):
            jobs.append((0, self._gsf_plan(xe, dim, False), n, xo, xt, xnd, "x%d" % i))
        for _, st, G, _, _, _, _ in jobs:
            if dim == 3:
                per += st[0][0] * nel[1] * G[0] * nq[2] * nq[1] + st[1][0] * G[1] * G[0] * nq[2]
            else:
                per += st[0][0] * G[0] * nq[1]
        budget = float(os.environ.get("TIGAR_B200_GSF_GB", "20")) * 2 ** 30
        lc_max = max(1, min(self.slab_hi - self.slab_lo, int(budget // (8 * per))))
        W = A.window if A is not None else None
        bufs = {}

        def buf(name, nelem):
            t = bufs.get(name)
            if t is None or t.numel() < nelem:
                t = dev.empty(nelem)
                bufs[name] = t
            return t
        plans = {}

        def dplan(tag, tab):
            if tag not in plans:
                plans[tag] = dev.from_np(tab.ravel())
            return plans[tag]
        # Opt-in (TIGAR_B200_GSF_OVERLAP=1): the Gauss-point kernel of chunk k+1 on a second
        # stream while the march stages of chunk k run on the caller's (two coefficient buffers,
        # an event per chunk each way).  Measured on B200 at 256^3: no gain (250.8 vs 252.6 ms per
        # step -- the kernels time-slice, both are bound by instruction issue), so the default
        # keeps one stream and clean per-kernel timings.
        import torch
        chunks = []
        k = self.slab_lo
        while k < self.slab_hi:
            lc = min(lc_max, self.slab_hi - k)
            chunks.append((k, lc))
            k += lc
        # algorithmic flops of the Gauss-point kernel per point: the straight-line program (one
        # flop per operation; transcendentals counted once) + the three contraction stages of
        # every jet (one FMA chain of length nloc per stage and distinct derivative prefix)
        byf = {}
        for (f_, comp, al) in P["prog"].jets:
            byf.setdefault((f_, comp), []).append(pad3(al))
        jfma = 0
        for lst in byf.values():
            jfma += (len(set(a[0] for a in lst)) * nl[0]
                     + len(set((a[0], a[1]) for a in lst)) * (nl[1] if dim > 1 else 0)
                     + len(lst) * (nl[2] if dim > 2 else 0))
        qp_flops = float(len(P["prog"].prog) + 2 * jfma)
        overlap = len(chunks) > 1 and os.environ.get("TIGAR_B200_GSF_OVERLAP", "0") == "1"
        main = torch.cuda.current_stream()
        side = main
        if overlap:
            if getattr(self, "_side_stream", None) is None:
                self._side_stream = torch.cuda.Stream()
            side = self._side_stream
            side.wait_stream(main)           # the coefficient functions were produced on main
        xsize = nslots * max(c[1] for c in chunks) * plane_cells * nqp
        xbufs = [buf("X0a", xsize), buf("X0b", xsize) if overlap else None]
        qp_done = [None] * len(chunks)
        readers_done = [None] * len(chunks)

        def launch_qp(ci):
            kk, ll = chunks[ci]
            nc_ = ll * plane_cells
            Xc = xbufs[ci % 2] if overlap else xbufs[0]
            with torch.cuda.stream(side):
                if overlap and ci >= 2:      # the buffer is free once chunk ci-2 has been read
                    side.wait_event(readers_done[ci - 2])
                with dev.PROF.range("tigar_qp (generated Gauss-point kernel)",
                                    8 * nslots * nc_ * nqp, qp_flops * nc_ * nqp):
                    self._qp_eval(B, P, kk * plane_cells, nc_, Xc, gsf=(nc_ * nqp, ll, kk))
                if overlap:
                    qp_done[ci] = torch.cuda.Event()
                    qp_done[ci].record(side)
            return Xc

        Xnext = launch_qp(0)
        for ci, (k, lc) in enumerate(chunks):
            ncells = lc * plane_cells
            X0 = Xnext
            if overlap:
                main.wait_event(qp_done[ci])
                if ci + 1 < len(chunks):
                    Xnext = launch_qp(ci + 1)
            elif ci + 1 < len(chunks):
                pass                          # launched after this chunk's stages (same buffer)
            for pair, st, G, out, tabs, nd, tag in jobs:
                perm = 0
                rbs = [self._gsf_dir(B, Wg, d, bool(pair))[0] for d in range(dim)]
                outp = dev.ptr(out.vals) if pair else dev.ptr(out)
                if dim == 3:
                    nq2c = lc * nq[2]
                    # stage 1: march e0; inner = (e1, q2c, q1l)
                    n1o, mi1, t1 = st[0]
                    Y1 = buf(tag + "Y1", n1o * nel[1] * G[0] * nq2c * nq[1])
                    ninner = nel[1] * nq2c * nq[1]
                    nin_tot = int(t1[:, 0].sum())
                    with dev.PROF.range("k_gsf %s stage 1" % ("matrix" if pair else "vector"),
                                        8 * (nin_tot * ncells * nqp + n1o * nel[1] * G[0] * nq2c * nq[1]),
                                        2.0 * nin_tot * ncells * nqp * nl[0] * (nl[0] if pair else 1)):
                      check(lib.tg_gsf_stage(
                        dev.ptr(X0), ncells * nqp, ninner * nq[0], 0, 0, nel[0], nel[0], nl[0],
                        nq[0], nd, tabs[0][0], tabs[0][1], dev.ptr(rbs[0]),
                        dev.ptr(dplan((tag, 0), t1)), n1o, mi1, pair, ninner, nq2c, nq[1],
                        dev.ptr(Y1), nel[1] * G[0] * nq2c * nq[1], nq2c * nq[1],
                        G[0] * nq2c * nq[1], nq[1], 0, None, 0, 0, 0, None, 0, None, None,
                        dev.stream()))
                    # stage 2: march e1; inner = (f0, e2l, q2l).  Matrices: the pair index
                    # (f1, f0) is written in the thread order of the last stage (perm)
                    perm = 1 if pair and os.environ.get("TIGAR_B200_GSF_PERM", "1") == "1" else 0
                    n2o, mi2, t2 = st[1]
                    Y2 = buf(tag + "Y2", n2o * lc * G[1] * G[0] * nq[2])
                    ninner = G[0] * nq2c
                    nin_tot = int(t2[:, 0].sum())
                    y1sz = nel[1] * G[0] * nq2c * nq[1]
                    with dev.PROF.range("k_gsf %s stage 2" % ("matrix" if pair else "vector"),
                                        8 * (nin_tot * y1sz + n2o * lc * G[1] * G[0] * nq[2]),
                                        2.0 * nin_tot * y1sz * nl[1] * (nl[1] if pair else 1)):
                      check(lib.tg_gsf_stage(
                        dev.ptr(Y1), nel[1] * G[0] * nq2c * nq[1], ninner * nq[1], 0, 0, nel[1],
                        nel[1], nl[1], nq[1], nd, tabs[1][0], tabs[1][1], dev.ptr(rbs[1]),
                        dev.ptr(dplan((tag, 1), t2)), n2o, mi2, pair, ninner, lc, nq[2],
                        dev.ptr(Y2), lc * G[1] * G[0] * nq[2],
                        nq[2] if perm else G[0] * nq[2], nq[2],
                        G[1] * G[0] * nq[2], 0, None, G[0] if perm else 0, 0, 0, None, perm,
                        W.ref() if perm else None,
                        dev.ptr(self._gsf_perm_rows(W, G[0]))
                        if perm and os.environ.get("TIGAR_B200_GSF_PROW", "1") == "1" else None,
                        dev.stream()))
                    Xl, skl, scl, ninl = Y2, lc * G[1] * G[0] * nq[2], G[1] * G[0] * nq[2], G[1] * G[0]
                else:
                    # stage 1: march e0; inner = (e1c, q1l)
                    n1o, mi1, t1 = st[0]
                    Y1 = buf(tag + "Y1", n1o * lc * G[0] * nq[1])
                    ninner = lc * nq[1]
                    nin_tot = int(t1[:, 0].sum())
                    with dev.PROF.range("k_gsf %s stage 1" % ("matrix" if pair else "vector"),
                                        8 * (nin_tot * ncells * nqp + n1o * lc * G[0] * nq[1]),
                                        2.0 * nin_tot * ncells * nqp * nl[0] * (nl[0] if pair else 1)):
                      check(lib.tg_gsf_stage(
                        dev.ptr(X0), ncells * nqp, ninner * nq[0], 0, 0, nel[0], nel[0], nl[0],
                        nq[0], nd, tabs[0][0], tabs[0][1], dev.ptr(rbs[0]),
                        dev.ptr(dplan((tag, 0), t1)), n1o, mi1, pair, ninner, 1, nq[1],
                        dev.ptr(Y1), lc * G[0] * nq[1], nq[1], G[0] * nq[1], 0, 0, None, 0, 0, 0,
                        None, 0, None, None, dev.stream()))
                    Xl, skl, scl, ninl = Y1, lc * G[0] * nq[1], G[0] * nq[1], G[0]
                nlo, mil, tl = st[L]
                nin_tot = int(tl[:, 0].sum())
                if pair:       # bytes: the stage input once + the rows of C these layers finish
                    outb = 8.0 * W.nnz * lc / max(1, self.slab_hi - self.slab_lo)
                else:
                    outb = 8.0 * ninl * lc
                with dev.PROF.range("k_gsf %s stage %d (writes %s)" % (
                        "matrix" if pair else "vector", dim, "C" if pair else "b"),
                        8 * nin_tot * skl + outb,
                        2.0 * nin_tot * skl * nl[L] * (nl[L] if pair else 1)):
                  check(lib.tg_gsf_stage(
                    dev.ptr(Xl), skl, scl, k, k, k + lc, nel[L], nl[L], nq[L], nd, tabs[L][0],
                    tabs[L][1], dev.ptr(rbs[L]), dev.ptr(dplan((tag, L), tl)), nlo, mil, pair,
                    ninl, 1, 1, None, 0, 0, 0, 0, 1, W.ref() if pair else None, G[0],
                    vrow0, vnr, outp, perm if (pair and dim == 3) else 0,
                    W.ref() if (pair and dim == 3 and perm) else None, None, dev.stream()))
            if overlap:
                readers_done[ci] = torch.cuda.Event()
                readers_done[ci].record(main)
            elif ci + 1 < len(chunks):
                Xnext = launch_qp(ci + 1)
        if overlap:
            main.wait_stream(side)
        self.launches += 1

    def assemble_vector_and_diag(self, vterms, mterms, funcs):
        """(b, diag C) from ONE Gauss-point pass: the load vector of ``vterms`` and the diagonal
        of the extracted matrix of ``mterms`` (its Jacobi preconditioner / the data the FD
        preconditioner is fitted to) without the matrix.  N_i is a tensor product, so
            C_ii = sum_q c(q) prod_d D^{a_d}N_{i_d}(q_d) D^{b_d}N_{i_d}(q_d)
        is a load-vector assembly against 1-D tables of PRODUCTS D^s N D^t N: a second vector
        job of the march, fed by the same coefficient buffer.  None if the march does not cover
        the case (more than 3 product columns in a direction, slab partition, 1-D)."""
        dim = self.dim
        if dim not in (2, 3) or self.part is not None:
            return None
        vkeys = sorted(vterms)
        groups = {}
        for k in sorted(mterms):          # (s,t) and (t,s) with one coefficient: one slot
            a, c = pad3(k[0])[:3], pad3(k[1])[:3]
            groups.setdefault((min(a, c), max(a, c), mterms[k].uid), []).append(k)
        dnodes, dpairs = [], []
        for (a, c, _), ks in sorted(groups.items()):
            node = mterms[ks[0]]
            dnodes.append(node if len(ks) == 1 else S.mul(S.const(float(len(ks))), node))
            dpairs.append((a, c))
        cols = [sorted(set((min(a[d], c[d]), max(a[d], c[d])) for a, c in dpairs))
                for d in range(dim)]
        if any(len(cl) > 3 for cl in cols):
            return None
        uniq, slot = self._dedupe([vterms[k] for k in vkeys] + dnodes)
        P = self._qp_setup(uniq, funcs)
        nder = max([P["nder"]] + [max(pad3(k)[:3]) for k in vkeys]
                   + [max(a + c) for a, c in dpairs])
        B = self.basis("iga", nder)
        if not self._gsf_ok(B, None, P):
            return None
        nd = B.nder + 1
        key = ("ptab", nder, tuple(map(tuple, cols)))
        ptabs = self._bases.get(key)
        if ptabs is None:                  # 1-D tables (nel x nq x nloc x 3): host, once
            ptabs = []
            for d in range(dim):
                t = dev.to_np(B.keep[d]["tabN"]).reshape(-1, nd)
                pt = np.zeros((t.shape[0], 3))
                for j, (s_, t_) in enumerate(cols[d]):
                    pt[:, j] = t[:, s_] * t[:, t_]
                ptabs.append(dev.from_np(pt.ravel()))
            self._bases[key] = ptabs
        vent = [(slot[i], tuple((pad3(k)[d], 0) for d in range(dim))) for i, k in enumerate(vkeys)]
        dent = [(slot[len(vkeys) + j],
                 tuple((cols[d].index((min(a[d], c[d]), max(a[d], c[d]))), 0) for d in range(dim)))
                for j, (a, c) in enumerate(dpairs)]
        b, dg = dev.empty(B.ntot), dev.empty(B.ntot)
        xt = [(dev.ptr(ptabs[d]), B.c.idx[d]) for d in range(dim)]
        self._gsf_run(B, "iga", P, len(uniq), None, vent, None, b, 0, int(B.c.n[dim - 1]),
                      xvec=[(dent, dg, xt, 3)])
        return b, dg

    def assemble_matrix(self, terms, funcs, kind="fe", out=None, cache=None):
        """terms: {(alphaTest, alphaTrial): Node} (coefficient already includes
        J and the quadrature weight).  kind 'fe' -> A_FE on the Lagrange
        basis; kind 'iga' -> sum_e M_e^T K_e M_e on the spline basis.
        ``cache``: see ``_setup_cached`` (same terms on every call)."""
        alS = sorted(set(pad3(k[0]) for k in terms))
        alT = sorted(set(pad3(k[1]) for k in terms))
        nS, nT = len(alS), len(alT)
        W = self.window("A" if kind == "fe" else "C")
        A = WinMatrix(W) if out is None else out
        order = max(max(max(a) for a in alS + alT), self.jet_order(list(terms.values())))
        B = self.basis(kind, order)
        stride = i32arr([2] * self.dim if kind == "fe" else B.nloc)
        if out is None and self.dim in (2, 3):
            keys = sorted(terms, key=self._sf_key)
            # terms with the same coefficient (symmetric forms: (a,b) and (b,a)) share a slot
            uniq, slot = self._dedupe([terms[k] for k in keys])
            gcache = None if cache is None else cache.setdefault("gsf", {})
            P = self._setup_cached(gcache, lambda: uniq, funcs)
            if self._gsf_ok(B, W, P):
                ent = [(slot[i], tuple((pad3(k[0])[d], pad3(k[1])[d]) for d in range(self.dim)))
                       for i, k in enumerate(keys)]
                A = WinMatrix(W, dev.empty(W.storage()))
                self._gsf_run(B, kind, P, len(uniq), ent, None, A, None, 0, 0)
                return A
        if lib.tg_assemble_sf_supported(B.ref()):
            # sum-factorised kernel: one coefficient slot per non-zero term
            keys = sorted(terms, key=self._sf_key)
            P = self._setup_cached(cache, lambda: [terms[k] for k in keys], funcs)
            tl = []
            for i, k in enumerate(keys):
                tl += [i] + list(pad3(k[0])) + list(pad3(k[1]))
            h_terms = i32arr(tl)
            nslots = len(keys)
            buf = None
            for cell0, nc in self._cell_chunks(nslots * B.nqp * 8):
                if buf is None or buf.numel() < nc * nslots * B.nqp:
                    buf = dev.empty(nc * nslots * B.nqp)
                self._qp_eval(B, P, cell0, nc, buf)
                check(lib.tg_assemble_matrix_terms(B.ref(), W.ref(), nslots, h_terms, nslots,
                                                   stride, dev.ptr(buf), cell0, nc,
                                                   dev.ptr(A.vals), dev.stream()))
            return A
        grid = [[S.ZERO] * nT for _ in range(nS)]
        for (a, b), node in terms.items():
            grid[alS.index(pad3(a))][alT.index(pad3(b))] = node
        P = self._setup_cached(cache, lambda: [grid[s][t] for s in range(nS) for t in range(nT)],
                               funcs)
        aS = i32arr([x for a in alS for x in a])
        aT = i32arr([x for a in alT for x in a])
        per_cell = nS * nT * B.nqp * 8
        buf = None
        for cell0, nc in self._cell_chunks(per_cell):
            if buf is None or buf.numel() < nc * nS * nT * B.nqp:
                buf = dev.empty(nc * nS * nT * B.nqp)
            self._qp_eval(B, P, cell0, nc, buf)
            check(lib.tg_assemble_matrix_ex(B.ref(), W.ref(), nS, aS, nT, aT, stride,
                                            dev.ptr(buf), cell0, nc, dev.ptr(A.vals),
                                            dev.stream()))
        return A

    def assemble_system(self, mterms, vterms, funcs, kind="fe"):
        """Matrix and vector of one linear system from a single Gauss-point
        pass (the geometry sub-expressions are shared by hash-consing).
        Returns (WinMatrix, vector)."""
        alS = sorted(set(pad3(k) for k in vterms))
        mk = sorted(mterms, key=self._sf_key)
        nodes = [mterms[k] for k in mk] + [vterms[a] for a in
                                           sorted(vterms, key=lambda a: pad3(a))]
        order = max([max(pad3(k[0]) + pad3(k[1])) for k in mk] + [max(a) for a in alS]
                    + [self.jet_order(nodes)])
        B = self.basis(kind, order)
        W = self.window("A" if kind == "fe" else "C")
        P = None
        if self.dim in (2, 3) and (self.part is None or kind == "iga"):
            uniq, slot = self._dedupe(nodes)      # symmetric terms share a coefficient slot
            Pg = self._qp_setup(uniq, funcs)
            if self._gsf_ok(B, W, Pg):
                dim = self.dim
                nm = len(mk)
                ment = [(slot[i], tuple((pad3(k[0])[d], pad3(k[1])[d]) for d in range(dim)))
                        for i, k in enumerate(mk)]
                vent = [(slot[nm + i], tuple((a[d], 0) for d in range(dim)))
                        for i, a in enumerate(alS)]
                A = WinMatrix(W, dev.empty(W.storage()))
                if self.part is not None:
                    b = dev.empty(self.n_loc)
                    r0, nr = self.pp["k0"], self.pp["k1"] - self.pp["k0"]
                else:
                    b = dev.empty(B.ntot)
                    r0, nr = 0, int(B.c.n[dim - 1])
                self._gsf_run(B, kind, Pg, len(uniq), ment, vent, A, b, r0, nr)
                return A, b
        if not lib.tg_assemble_sf_supported(B.ref()):
            return (self.assemble_matrix(mterms, funcs, kind),
                    self.assemble_vector(vterms, funcs, kind))
        A = WinMatrix(W)
        if self.part is not None:
            if kind != "iga":
                raise NotImplementedError("slab partition exists for the fused path only")
            b = dev.zeros(self.n_loc)
            vrow0 = i32arr([0] * (self.dim - 1) + [self.pp["k0"]])
            vnr = i32arr(self.ncp[:-1] + [self.pp["k1"] - self.pp["k0"]])
        else:
            b = dev.zeros(B.ntot)
            vrow0 = i32arr([0] * self.dim)
            vnr = i32arr([B.c.n[d] for d in range(self.dim)])
        if P is None:
            P = self._qp_setup(nodes, funcs)
        nm, nv = len(mk), len(alS)
        nslots = nm + nv
        tl = []
        for i, k in enumerate(mk):
            tl += [i] + list(pad3(k[0])) + list(pad3(k[1]))
        h_terms = i32arr(tl)
        aS = i32arr([x for a in alS for x in a])
        vslots = i32arr([nm + i for i in range(nv)])
        stride = i32arr([2] * self.dim if kind == "fe" else B.nloc)
        buf = None
        for cell0, nc in self._cell_chunks(nslots * B.nqp * 8):
            if buf is None or buf.numel() < nc * nslots * B.nqp:
                buf = dev.empty(nc * nslots * B.nqp)
            self._qp_eval(B, P, cell0, nc, buf)
            check(lib.tg_assemble_matrix_terms(B.ref(), W.ref(), nm, h_terms, nslots, stride,
                                               dev.ptr(buf), cell0, nc, dev.ptr(A.vals),
                                               dev.stream()))
            check(lib.tg_assemble_vector_part(B.ref(), nv, aS, vslots, nslots, stride, vrow0, vnr,
                                              dev.ptr(buf), cell0, nc, dev.ptr(b),
                                              dev.stream()))
        return A, b

    def assemble_vector(self, terms, funcs, kind="fe", out=None, cache=None):
        """terms: {alphaTest: Node}.  ``cache``: a dict that keeps the compiled
        Gauss-point program between calls with the SAME terms and the same function
        tensors (the matrix-free operator calls this once per CG iteration)."""
        alS = sorted(set(pad3(k) for k in terms))
        nS = len(alS)
        def outputs():
            out_nodes = [S.ZERO] * nS
            for a, node in terms.items():
                out_nodes[alS.index(pad3(a))] = node
            return out_nodes
        P = self._setup_cached(cache, outputs, funcs)
        nder = max(P["nder"], max(max(a) for a in alS))
        B = self.basis(kind, nder)
        if self.dim in (2, 3) and (self.part is None or kind == "iga") and \
                self._gsf_ok(B, None, P):
            dim = self.dim
            vent = [(i, tuple((a[d], 0) for d in range(dim))) for i, a in enumerate(alS)]
            if self.part is not None:
                b = dev.empty(self.n_loc) if out is None else out
                r0, nr = self.pp["k0"], self.pp["k1"] - self.pp["k0"]
            else:
                b = dev.empty(B.ntot) if out is None else out
                r0, nr = 0, int(B.c.n[dim - 1])
            self._gsf_run(B, kind, P, nS, None, vent, None, b, r0, nr)
            return b
        if self.part is not None:
            if kind != "iga":
                raise NotImplementedError("slab partition exists for the fused path only")
            b = dev.zeros(self.n_loc) if out is None else out
            vrow0 = i32arr([0] * (self.dim - 1) + [self.pp["k0"]])
            vnr = i32arr(self.ncp[:-1] + [self.pp["k1"] - self.pp["k0"]])
        else:
            b = dev.zeros(B.ntot) if out is None else out
            vrow0 = i32arr([0] * self.dim)
            vnr = i32arr([B.c.n[d] for d in range(self.dim)])
        stride = i32arr([2] * self.dim if kind == "fe" else B.nloc)
        aS = i32arr([x for a in alS for x in a])
        slots = i32arr(list(range(nS)))
        buf = None
        for cell0, nc in self._cell_chunks(nS * B.nqp * 8):
            if buf is None or buf.numel() < nc * nS * B.nqp:
                buf = dev.empty(nc * nS * B.nqp)
            self._qp_eval(B, P, cell0, nc, buf)
            check(lib.tg_assemble_vector_part(B.ref(), nS, aS, slots, nS, stride, vrow0, vnr,
                                              dev.ptr(buf), cell0, nc, dev.ptr(b),
                                              dev.stream()))
        return b

    def assemble_scalar(self, node, funcs, kind="fe"):
        """sum over all Gauss points of ``node`` (functional assembly)."""
        P = self._qp_setup([node], funcs)
        B = self.basis(kind, P["nder"])
        tot = 0.0
        acc = dev.zeros(1)
        buf = None
        saved = (self.slab_lo, self.slab_hi)
        if self.part is not None:      # disjoint cell layers per rank, then all-reduce
            r, n = self.part
            self.slab_lo, self.slab_hi = (self.nel[-1] * r) // n, (self.nel[-1] * (r + 1)) // n
        try:
            for cell0, nc in self._cell_chunks(B.nqp * 8):
                if buf is None or buf.numel() < nc * B.nqp:
                    buf = dev.empty(nc * B.nqp)
                self._qp_eval(B, P, cell0, nc, buf)
                check(lib.tg_sum(dev.ptr(buf), nc * B.nqp, dev.ptr(acc), dev.stream()))
                tot += float(acc.item())
        finally:
            self.slab_lo, self.slab_hi = saved
        if self.part is not None:
            import torch
            import torch.distributed as dist
            t = torch.tensor([tot], dtype=torch.float64, device=dev.device())
            dist.all_reduce(t)
            tot = float(t.item())
        return tot

    # ---- (iii) triple product, BCs, solve -----------------------------------
    KRON_KB = 10

    def _kron_tabs(self):
        """Per-direction blocks tab[I][J-loA(I)][j-loP(I)] = M_d[J,j] for
        tg_ptap_kron_ap, and host copies of the 1-D extraction rows."""
        if getattr(self, "_ktabs", None) is None:
            KB = self.KRON_KB
            wA, wP = self._global_window("A"), self._global_window("P")
            tabs = []
            for d, D in enumerate(self.dirs):
                first = dev.to_np(D.m_first).astype(np.int64)
                vals = dev.to_np(D.m_vals)
                t = np.zeros((D.nfe, KB, KB))
                for I in range(D.nfe):
                    for J in range(wA.lo[d][I], wA.hi[d][I] + 1):
                        for k in range(D.p + 1):
                            j = first[J] + k
                            if D.m_lo[J] <= j <= D.m_hi[J]:
                                t[I, J - wA.lo[d][I], j - wP.lo[d][I]] = vals[J, k]
                tabs.append(dev.from_np(t))
            self._ktabs = tabs
        return self._ktabs

    def kron_supported(self):
        KB = self.KRON_KB
        wA, wP = self._global_window("A"), self._global_window("P")
        return all(int(wA.len[d].max()) <= KB and int(wP.len[d].max()) <= KB
                   for d in range(self.dim))

    def ptap_kron(self, A, keep=False):
        """C = M^T A M without reading the global M: tg_ptap_kron_ap (one pass
        over A) followed by one tg_win_rowcombine per direction."""
        wA, wP, wC, wMT = (self._global_window(k) for k in ("A", "P", "C", "MT"))
        tabs = self._kron_tabs()
        box = int(np.prod([max(int(wA.len[d].max()), int(wP.len[d].max()))
                           for d in range(self.dim)]))
        X = WinMatrix(wP, dev.empty(wP.nnz))
        check(lib.tg_ptap_kron_ap(wA.ref(), dev.ptr(A.vals), vparr([dev.ptr(t) for t in tabs]),
                                  wP.ref(), dev.ptr(X.vals), box, dev.stream()))
        stages = [X]
        for d, D in enumerate(self.dirs):
            key = "K%d" % d
            if key not in self._win:
                wX = stages[-1].window
                nr = list(wX.nr)
                nr[d] = D.ncp
                lo = list(wX.lo)
                hi = list(wX.hi)
                lo[d], hi[d] = wC.lo[d], wC.hi[d]
                self._win[key] = Window(nr, wX.nc, lo, hi)
            wY = self._win[key]
            Y = WinMatrix(wY, dev.empty(wY.nnz))
            if not hasattr(D, "d_slo"):
                D.d_slo = dev.from_np(wMT.lo[d].astype(np.int32))
                D.d_shi = dev.from_np(wMT.hi[d].astype(np.int32))
            check(lib.tg_win_rowcombine(stages[-1].window.ref(), dev.ptr(stages[-1].vals),
                                        wY.ref(), dev.ptr(Y.vals), d, dev.ptr(D.m_first),
                                        dev.ptr(D.m_vals), D.p + 1, dev.ptr(D.d_slo),
                                        dev.ptr(D.d_shi), dev.stream()))
            stages.append(Y)
            if not keep and len(stages) > 2:
                stages[-2] = None
        Cw = stages[-1]
        assert Cw.window.nnz == wC.nnz        # same pattern as the global C window
        out = WinMatrix.from_csr_values(self.window("C"), Cw.vals)
        return (out, [s for s in stages if s is not None]) if keep else out

    # ---- two-sided march passes (tg_ptap_march) --------------------------------
    MARCH_THREADS = 256
    MARCH_VARIANT = int(os.environ.get("TIGAR_B200_MARCH_VARIANT", "2"))

    def _march_setup(self):
        """Per-direction tables, intermediate windows and CTA tilings of the
        march PtAP; None when the patch does not have the structure the kernel
        assumes (then ptap() uses the row-wise Kronecker kernels)."""
        if getattr(self, "_march", 0) != 0:
            return self._march
        self._march = None
        if self.dim < 2:
            return None
        wA, wC, wMT = (self._global_window(k) for k in ("A", "C", "MT"))
        dirs = []
        for d, D in enumerate(self.dirs):
            p = D.p
            first = dev.to_np(D.m_first).astype(np.int64)
            T = march_tables.dir_tables(p, D.pf, first, dev.to_np(D.m_vals), D.m_lo, D.m_hi,
                                        wA.lo[d], wA.hi[d], wC.lo[d], wC.hi[d],
                                        self.MARCH_RMAX)
            if T is None:
                return None
            mrow, tabc, KA, SX = T["mrow"], T["tabc"], T["KA"], T["SX"]
            irec, jrec, cpad = T["irec"], T["jrec"], T["cpad"]
            grp, gidx, GMAX = T["grp"], T["gidx"], T["GMAX"]
            dirs.append(dict(p=p, KA=KA, first=dev.from_np(first.astype(np.int32)),
                             irec=dev.from_np(irec.astype(np.uint32).view(np.int32)),
                             jrec=dev.from_np(jrec.astype(np.uint32).view(np.int32)),
                             Sx=dev.from_np(SX[:-1].copy()), cpad=dev.from_np(cpad),
                             grp=dev.from_np(grp.astype(np.int32)), h_grp=grp, h_gidx=gidx,
                             GMAX=GMAX,
                             loX=wA.lo[d].astype(np.int64), hiX=wA.hi[d].astype(np.int64),
                             h_slo=wMT.lo[d].astype(np.int64), h_shi=wMT.hi[d].astype(np.int64),
                             mrow=dev.from_np(mrow), tabc=dev.from_np(tabc),
                             slo=dev.from_np(wMT.lo[d].astype(np.int32)),
                             shi=dev.from_np(wMT.hi[d].astype(np.int32))))

        def groups(lens, cap):
            g, tot = [0], 0
            for r, L in enumerate(lens):
                if tot + L > cap and tot > 0:
                    g.append(r)
                    tot = 0
                tot += int(L)
            g.append(len(lens))
            return g

        passes = []
        wX = wA
        for d in range(self.dim):
            nr = [self.ncp[k] if k <= d else self.nfe[k] for k in range(self.dim)]
            lo = [wC.lo[k] if k <= d else wA.lo[k] for k in range(self.dim)]
            hi = [wC.hi[k] if k <= d else wA.hi[k] for k in range(self.dim)]
            wY = Window(nr, nr, lo, hi)
            others = [k for k in range(self.dim) if k != d]
            cap = 16 if self.dim == 3 else self.MARCH_THREADS
            ga = groups(wX.len[others[0]], cap)
            gb = groups(wX.len[others[1]], cap) if self.dim == 3 else [0, 1]

            def gmax(g, lens):
                return (max(int(lens[g[k]:g[k + 1]].sum()) for k in range(len(g) - 1)),
                        max(g[k + 1] - g[k] for k in range(len(g) - 1)))
            Fa, na = gmax(ga, wX.len[others[0]])
            Fb, nb = gmax(gb, wX.len[others[1]]) if self.dim == 3 else (1, 1)
            assert Fa * Fb <= self.MARCH_THREADS
            KAmax = int(wX.len[d].max())
            CW = 2 * dirs[d]["p"] + 1
            maxlines = na * nb
            stage = (KAmax * Fa * Fb + 3 * maxlines + 1) & ~1
            outd = CW * Fa * Fb
            nseg = int(min(max(1, self.ncp[d] // 16),
                           max(1, -(-4 * 148 // ((len(ga) - 1) * (len(gb) - 1))))))
            seg = [(self.ncp[d] * k) // nseg for k in range(nseg + 1)]
            for w_ in (wX, wY):     # 32-bit line constants of tg_ptap_march_w
                tot = [int(l.sum()) for l in w_.len]
                if 16 * tot[0] * (tot[1] if self.dim > 1 else 1) >= 2 ** 32 or max(tot) >= 2 ** 31:
                    return None
            tasks = self._march_tasks(wX.len[others[0]],
                                      wX.len[others[1]] if self.dim == 3 else np.ones(1, np.int64))
            nsegw = int(min(max(1, self.ncp[d] // 16),
                            max(1, -(-16 * 148 * 2 // len(tasks)))))
            Dd = dirs[d]
            while True:        # bound the per-segment tables held in shared memory
                segw = [(self.ncp[d] * k) // nsegw for k in range(nsegw + 1)]
                gA = [int(Dd["h_gidx"][Dd["h_slo"][segw[k]]]) for k in range(nsegw)]
                gB = [int(Dd["h_gidx"][Dd["h_shi"][segw[k + 1] - 1]]) for k in range(nsegw)]
                nodes = [int(Dd["hiX"][Dd["h_grp"][b_ + 1] - 1] - Dd["loX"][Dd["h_grp"][a_]] + 1)
                         for a_, b_ in zip(gA, gB)]
                # shared-memory budget of two resident CTAs (tg_ptap_march_w: NS = 3 ring
                # stages of 32*GMAX doubles per warp, 8 warps -- 4 for p = 4)
                stgd = 32 * Dd["GMAX"] + 10 * int(tasks[:, 0].max()) + 4
                warpb = (3 * stgd + 64 + 3 + 48 + 1) * 8
                pernode = (Dd["p"] + 4) * 8 + 20
                # 16 warps per CTA (one CTA per SM) share one table copy: longer segments;
                # needs room for the ring of 16 warps plus a useful table
                wide = (Dd["p"] <= 3 and os.environ.get("TIGAR_B200_MARCH_WPC", "16") != "8"
                        and 16 * warpb + 100 * pernode + 4096 <= 220 * 1024)
                wpc, ncta = (4, 2) if Dd["p"] >= 4 else ((16, 1) if wide else (8, 2))
                budget = min(220 * 1024, (233472 - 1024 * ncta) // ncta) - wpc * warpb - 2048
                nodemax = max(32, min(self.MARCH_NODEMAX, budget // pernode))
                if max(nodes) <= nodemax or nsegw >= self.ncp[d]:
                    break
                nsegw += 1
            maxnodes = max(nodes)
            maxgroups = max(b_ - a_ + 1 for a_, b_ in zip(gA, gB))
            maxrows = max(segw[k + 1] - segw[k] for k in range(nsegw))
            passes.append(dict(wX=wX, wY=wY, d=d, KAmax=KAmax, maxlines=maxlines, stage=stage,
                               tasks=dev.from_np(tasks.ravel()), ntask=len(tasks), nsegw=nsegw,
                               maxnodes=maxnodes, maxrows=maxrows, maxgroups=maxgroups,
                               maxpieces=int(tasks[:, 0].max()), wpc=wpc,
                               segw=dev.from_np(np.array(segw, dtype=np.int32)),
                               outd=outd, nga=len(ga) - 1, ngb=len(gb) - 1, nseg=nseg,
                               ga=dev.from_np(np.array(ga, dtype=np.int32)),
                               gb=dev.from_np(np.array(gb, dtype=np.int32)),
                               seg=dev.from_np(np.array(seg, dtype=np.int32))))
            wX = wY
        assert passes[-1]["wY"].nnz == wC.nnz
        self._march = (dirs, passes)
        return self._march

    MARCH_MAXSUB = 8
    MARCH_NODEMAX = 400       # FE nodes of one march segment whose tables sit in shared memory
    MARCH_RMAX = 4            # rows per march group (TGW_RMAX)

    @classmethod
    def _march_tasks(cls, lena, lenb):
        """Warp tasks of tg_ptap_march_w: every line (ra, rb) of the two
        non-march directions has la*lb fibres; lines wider than a warp are cut
        along the second direction, narrow ones are packed up to MARCH_MAXSUB to a
        warp.  Returns int32 [ntask][4*MAXSUB+4]: npieces,0,0,0, {ra,rb,cb0,ncb}*."""
        MS = cls.MARCH_MAXSUB
        tasks, open_ = [], []          # open_: [lanes used, [pieces]]
        minp = int(min(lena.min() * 1, 32))
        for rb, lb in enumerate(lenb):
            for ra, la in enumerate(lena):
                la, lb = int(la), int(lb)
                assert la <= 32
                ncbmax = max(1, 32 // la)
                nsplit = -(-lb // ncbmax)
                cuts = [(lb * k) // nsplit for k in range(nsplit + 1)]
                for k in range(nsplit):
                    cb0, ncb = cuts[k], cuts[k + 1] - cuts[k]
                    n = la * ncb
                    for T in open_:
                        if T[0] + n <= 32 and len(T[1]) < MS:
                            T[0] += n
                            T[1].append((ra, rb, cb0, ncb))
                            break
                    else:
                        T = [n, [(ra, rb, cb0, ncb)]]
                        open_.append(T)
                    if T[0] + minp > 32 or len(T[1]) == MS:
                        open_.remove(T)
                        tasks.append(T[1])
                    elif len(open_) > 4:
                        tasks.append(open_.pop(0)[1])
        tasks += [T[1] for T in open_]
        out = np.zeros((len(tasks), 4 * MS + 4), dtype=np.int32)
        for t, pcs in enumerate(tasks):
            out[t, 0] = len(pcs)
            for k, pc in enumerate(pcs):
                out[t, 4 + 4 * k:8 + 4 * k] = pc
        return out

    def ptap_march(self, A, keep=False):
        """C = M^T A M by one two-sided march pass per direction (A read once,
        two shrinking intermediates, M never formed)."""
        dirs, passes = self._march_setup()
        X = A.vals
        stages = []
        for P_ in passes:
            D = dirs[P_["d"]]
            Y = dev.empty(P_["wY"].nnz)
            if self.MARCH_VARIANT == 2:
                check(lib.tg_ptap_march_w(P_["wX"].ref(), dev.ptr(X), P_["wY"].ref(), dev.ptr(Y),
                                          P_["d"], D["p"], D["GMAX"], dev.ptr(D["irec"]),
                                          dev.ptr(D["Sx"]), dev.ptr(D["jrec"]),
                                          dev.ptr(D["cpad"]), dev.ptr(D["grp"]), dev.ptr(D["slo"]),
                                          dev.ptr(D["shi"]), dev.ptr(P_["tasks"]), P_["ntask"],
                                          dev.ptr(P_["segw"]), P_["nsegw"], P_["maxnodes"],
                                          P_["maxrows"], P_["maxgroups"], P_["maxpieces"], P_["wpc"], dev.stream()))
                X = Y
                if keep:
                    stages.append(WinMatrix(P_["wY"], Y))
                continue
            check(lib.tg_ptap_march(P_["wX"].ref(), dev.ptr(X), P_["wY"].ref(), dev.ptr(Y),
                                    P_["d"], D["p"], D["KA"], P_["KAmax"], dev.ptr(D["first"]),
                                    dev.ptr(D["mrow"]), dev.ptr(D["tabc"]), dev.ptr(D["slo"]),
                                    dev.ptr(D["shi"]), dev.ptr(P_["ga"]), P_["nga"],
                                    dev.ptr(P_["gb"]), P_["ngb"], dev.ptr(P_["seg"]),
                                    P_["nseg"], P_["stage"], P_["outd"], P_["maxlines"],
                                    self.MARCH_VARIANT, dev.stream()))
            X = Y
            if keep:
                stages.append(WinMatrix(P_["wY"], Y))
        out = WinMatrix.from_csr_values(self.window("C"), X)
        return (out, stages) if keep else out

    def ptap(self, A, M=None, keep_AP=False):
        """C = M^T A M on windowed operands (MatPtAP, common.py:1194-1195)."""
        from .generic import CsrMatrix, GenericPtAP
        if isinstance(M, CsrMatrix):
            return GenericPtAP(A, M)
        if not keep_AP and not os.environ.get("TIGAR_B200_PTAP_GENERIC"):
            if os.environ.get("TIGAR_B200_PTAP", "march") == "march" \
                    and self._march_setup() is not None:
                return self.ptap_march(A)
            if self.kron_supported():
                return self.ptap_kron(A)
        if M is None:
            M = self.build_M()
        wA, wM, wMT = self.window("A"), self.window("M"), self.window("MT")
        wP, wPT, wC = self.window("P"), self.window("PT"), self._global_window("C")
        AP = dev.empty(wP.nnz)
        check(lib.tg_ptap_ap(wA.ref(), dev.ptr(A.vals), wM.ref(), dev.ptr(M.vals), wMT.ref(),
                             wP.ref(), dev.ptr(AP), dev.stream()))
        Cm = WinMatrix(wC, dev.empty(wC.nnz))
        check(lib.tg_ptap_c(wM.ref(), dev.ptr(M.vals), wMT.ref(), wP.ref(), dev.ptr(AP),
                            wPT.ref(), wC.ref(), dev.ptr(Cm.vals), dev.stream()))
        Cm = WinMatrix.from_csr_values(self.window("C"), Cm.vals)
        if keep_AP:
            return Cm, WinMatrix(wP, AP)
        return Cm

    def bc_mask(self, zeroDofs):
        """0/1 mask over the IGA DoFs, built on the device from the zeroDofs list."""
        m = dev.zeros(self.n_iga, dev.U8)
        z = np.ascontiguousarray(zeroDofs, dtype=np.int64).ravel()
        if z.size:
            if z.min() < 0 or z.max() >= self.n_iga:
                raise IndexError("zero DoF outside [0, %d)" % self.n_iga)
            dz = dev.from_np(z)
            check(lib.tg_mask_set(dev.ptr(m), dev.ptr(dz), z.size, self.n_iga, dev.stream()))
        return m

    def mask_planes(self, mask):
        """(hp, exact): hp[d] = device uint8 flags of the hyperplanes of direction d that are
        constrained as a whole; exact = the mask is exactly their union (side DoFs).  Cached
        per mask tensor."""
        hit = getattr(self, "_hp_cache", None)
        if hit is not None and hit[0] is mask:
            return hit[1], hit[2]
        shape = tuple(reversed(self.ncp))
        m = mask.view(shape) != 0
        hp, nfree = [], 1
        for d in range(self.dim):
            ax = self.dim - 1 - d
            other = tuple(a for a in range(self.dim) if a != ax)
            full = (m.all(dim=other) if other else m).to(dev.U8).contiguous()
            hp.append(full)
            nfree *= self.ncp[d] - int(full.sum().item())
        exact = int(m.sum().item()) == self.n_iga - nfree
        self._hp_cache = (mask, hp, exact)
        return hp, exact

    def apply_bcs_matrix(self, Cm, mask, diag=1.0):
        Cm.bc_mask, Cm.bc_diag = mask, float(diag)      # read by the preconditioned solvers
        if Cm.window is None:                           # operator form (generic basis)
            return Cm
        if Cm.window.layout == 0 and os.environ.get("TIGAR_B200_BC_HP", "1") == "1":
            hp, exact = self.mask_planes(mask)
            if exact:
                P = [dev.ptr(hp[d]) if d < self.dim else None for d in range(3)]
                # local row coordinates per direction whose window reaches a constrained plane
                W = Cm.window
                sels = []
                for d in range(self.dim):
                    hflag = dev.to_np(hp[d]).astype(bool)
                    cs = np.concatenate([[0], np.cumsum(hflag)])
                    lo = W.lo[d].astype(np.int64) + W.col0[d]
                    hi = W.hi[d].astype(np.int64) + W.col0[d]
                    rows = np.arange(W.nr[d]) + W.row0[d]
                    need = (cs[hi + 1] - cs[lo] > 0) | hflag[rows]
                    sels.append(dev.from_np(np.flatnonzero(need).astype(np.int32)))
                check(lib.tg_win_zero_rows_cols_hp(
                    W.ref(), dev.ptr(Cm.vals), P[0], P[1], P[2], float(diag),
                    vparr([dev.ptr(t) or 0 for t in sels] + [0] * (3 - self.dim)),
                    i32arr([t.numel() for t in sels] + [0] * (3 - self.dim)), dev.stream()))
                return Cm
        if self.part is not None:
            pp, pl = self.pp, self.plane
            rowmask = mask[pp["k0"] * pl:pp["k1"] * pl]
            colmask = mask[pp["c0"] * pl:pp["c1"] * pl]
            check(lib.tg_win_zero_rows_cols(Cm.window.ref(), dev.ptr(Cm.vals), dev.ptr(rowmask),
                                            dev.ptr(colmask), float(diag),
                                            pp["k0"] - pp["c0"], dev.stream()))
            return Cm
        check(lib.tg_win_zero_rows_cols(Cm.window.ref(), dev.ptr(Cm.vals), dev.ptr(mask),
                                        dev.ptr(mask), float(diag), 0, dev.stream()))
        return Cm

    def apply_bcs_vector(self, b, mask):
        if self.part is not None:
            mask = mask[self.pp["k0"] * self.plane:self.pp["k1"] * self.plane]
        check(lib.tg_zero_entries(dev.ptr(b), dev.ptr(mask), b.numel(), dev.stream()))
        return b

    def solve(self, Cm, b, x=None, rtol=1e-12, atol=0.0, maxit=100000, method="auto",
              mask=None, diag=1.0):
        """solve() of common.py:1255-1258 on one windowed system.  ``method``:
        "direct" (band Cholesky), "fd" (CG preconditioned by fast diagonalisation),
        "jacobi" (Jacobi-CG), or "auto": the reference's default is a direct LU, so the band
        solver is used while it is affordable (2-D patches, small 3-D ones), FD-CG beyond.
        Returns (x, iterations, relative residual, method used)."""
        from . import solvers
        method = os.environ.get("TIGAR_B200_SOLVER", method)
        if self.part is not None and method in ("auto", "direct"):
            method = "fd"           # row-distributed system: FD-preconditioned CG over NCCL
        if method == "auto":
            import torch
            free, _ = torch.cuda.mem_get_info()
            method = "direct" if solvers.direct_affordable(Cm.window, free) else "fd"
        if method == "direct":
            bc = solvers.BandCholesky(Cm).factor()
            xs = bc.solve(b)
            # iterative refinement with the true residual (one SpMV + one pair of band solves
            # per pass): removes the round-off the factorisation accumulates over a wide band,
            # which is what limits the answer when cond ~ h^-4 (biharmonic, configs[2])
            bb = float(b.norm())
            r = Cm.matvec(xs)
            r.neg_().add_(b)
            rel = float(r.norm()) / bb if bb > 0 else 0.0
            for _ in range(3):
                if rel < 1e-15:
                    break
                dx = bc.solve(r)
                check(lib.tg_axpy(dev.ptr(dx), 1.0, dev.ptr(xs), dx.numel(), dev.stream()))
                r2 = Cm.matvec(dx)
                r2.neg_().add_(b)
                rel2 = float(r2.norm()) / bb if bb > 0 else 0.0
                if not rel2 < rel:
                    break
                xs, r, rel = dx, r2, rel2
            return xs, 1, rel, "direct"
        if method == "fd":
            if self.part is not None:
                from .multigpu import solve_fd_pcg_dist
                xs, its, rel = solve_fd_pcg_dist(self, Cm, b, mask, diag, rtol, atol, maxit)
                return xs, its, rel, "fd"
            xs, its, rel, _ = solvers.solve_fd_pcg(self, Cm, b, mask, diag, x, rtol, atol, maxit)
            return xs, its, rel, "fd"
        if method != "jacobi":
            raise ValueError("unknown solver method %r" % (method,))
        xs, its, rel = self.solve_cg(Cm, b, x, rtol, atol, maxit)
        return xs, its, rel, "jacobi"

    def solve_cg(self, Cm, b, x=None, rtol=1e-12, atol=0.0, maxit=100000, check_every=5):
        if self.part is not None:
            from .multigpu import DeviceOps, dist_cg
            ops = DeviceOps(self, Cm)
            xl, its, rel = dist_cg(ops, b, rtol, atol, maxit, check_every)
            return xl, its, rel
        n = Cm.window.nrows
        x = dev.zeros(n) if x is None else x
        work = dev.empty(4 * n + lib.tg_cg_scratch_len() + 8)
        its = C.c_int32(0)
        rel = C.c_double(0.0)
        check(lib.tg_win_solve_cg(Cm.window.ref(), dev.ptr(Cm.vals), dev.ptr(b), dev.ptr(x),
                                  float(rtol), float(atol), int(maxit), int(check_every),
                                  dev.ptr(work), C.byref(its), C.byref(rel), dev.stream()))
        return x, its.value, rel.value
