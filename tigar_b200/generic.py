"""
Generic ``AbstractScalarBasis`` on the device.

For a user basis the reference evaluates ``getNodesAndEvals`` at every FE node in a Python
loop and fills ``M`` row by row (common.py:1497-1509, interface :1683-1692).  That loop is the
user's code, so it stays on the host -- once, at extraction time -- and everything after it
runs on the device:

* ``CsrMatrix``      -- the extraction operator as a general CSR matrix in HBM
  (``tg_spmv`` for ``M x``, the device-built transpose for ``M^T b``),
* the FE side (``A_FE``, ``b_FE`` on the CG Q_p Lagrange mesh of the basis' tensor-product
  extraction mesh) by the same kernels as the csr mode of tensor-product splines,
* ``GenericPtAP``    -- ``M^T A_FE M`` in operator form (three SpMVs; the global triple product
  of an unstructured ``M`` is never needed by the Krylov solve), homogeneous BCs as
  ``P C P + diag (I - P)`` (zeroRowsColumns, common.py:1199-1200), CG on the C-ABI vector
  kernels.

Restriction: the basis' extraction mesh must be a tensor-product mesh (``generateMesh``
returns a ``TensorMesh``); disconnected element soups (T-splines, multi-patch,
RhinoTSplines.py / MultiBSpline) are a different mesh model and stay out of scope.
"""
import numpy as np

from . import dev
from ._lib import lib, check


class CsrMatrix(object):
    """General CSR matrix on the device (int64 row pointers, int32 columns, FP64 values)."""

    def __init__(self, rowptr, cols, vals, shape):
        self.rowptr, self.cols, self.vals = rowptr, cols, vals
        self.shape = (int(shape[0]), int(shape[1]))
        self._T = None

    @staticmethod
    def from_rows(rows, ncols, eps):
        """rows: per-row lists of [column, value] pairs (``getNodesAndEvals`` output);
        entries with |value| <= eps are dropped (common.py:1507-1509)."""
        rp = np.zeros(len(rows) + 1, dtype=np.int64)
        cols, vals = [], []
        for i, r in enumerate(rows):
            for c, v in r:
                if abs(v) > eps:
                    cols.append(int(c))
                    vals.append(float(v))
            rp[i + 1] = len(cols)
        cols = np.array(cols, dtype=np.int32)
        if cols.size and (cols.min() < 0 or cols.max() >= ncols):
            raise IndexError("basis function index outside [0, %d)" % ncols)
        return CsrMatrix(dev.from_np(rp), dev.from_np(cols),
                         dev.from_np(np.array(vals, dtype=np.float64)), (len(rows), ncols))

    def matvec(self, x, y=None):
        y = dev.empty(self.shape[0]) if y is None else y
        check(lib.tg_spmv(dev.ptr(self.rowptr), dev.ptr(self.cols), dev.ptr(self.vals),
                          dev.ptr(x), dev.ptr(y), self.shape[0], dev.stream()))
        return y

    def transpose(self):
        """CSR of the transpose, built on the device (stable sort of the COO triplets)."""
        if self._T is None:
            import torch
            counts = self.rowptr[1:] - self.rowptr[:-1]
            rows = torch.repeat_interleave(torch.arange(self.shape[0], device=self.cols.device,
                                                        dtype=torch.int64), counts)
            order = torch.argsort(self.cols.to(torch.int64), stable=True)
            tcols = rows[order].to(torch.int32).contiguous()
            tvals = self.vals[order].contiguous()
            tcount = torch.bincount(self.cols.to(torch.int64), minlength=self.shape[1])
            trp = torch.zeros(self.shape[1] + 1, dtype=torch.int64, device=self.cols.device)
            trp[1:] = torch.cumsum(tcount, 0)
            self._T = CsrMatrix(trp, tcols, tvals, (self.shape[1], self.shape[0]))
            self._T._T = self
        return self._T

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((dev.to_np(self.vals), dev.to_np(self.cols), dev.to_np(self.rowptr)),
                             shape=self.shape)


class GenericPtAP(object):
    """``M^T A M`` (+ homogeneous BCs) as an operator: extractMatrix of common.py:1176-1204 for
    an extraction operator without tensor-product structure."""

    window = None

    def __init__(self, A, M):
        self.A, self.M, self.MT = A, M, M.transpose()
        self.n = M.shape[1]
        self.shape = (self.n, self.n)
        self.bc_mask, self.bc_diag = None, 1.0
        self._t1 = dev.empty(M.shape[0])
        self._t2 = dev.empty(M.shape[0])
        self._xm = dev.empty(self.n)

    def apply_raw(self, x, y=None):
        """y = M^T A M x (no BCs)."""
        self.M.matvec(x, self._t1)
        self.A.matvec(self._t1, self._t2)
        return self.MT.matvec(self._t2, y)

    def matvec(self, x, y=None):
        y = dev.empty(self.n) if y is None else y
        m = self.bc_mask
        if m is None:
            return self.apply_raw(x, y)
        st = dev.stream()
        check(lib.tg_masked_copy(dev.ptr(self._xm), dev.ptr(x), dev.ptr(m), self.n, st))
        self.apply_raw(self._xm, y)
        check(lib.tg_zero_entries(dev.ptr(y), dev.ptr(m), self.n, st))
        # + diag (I - P) x
        check(lib.tg_axpy(dev.ptr(y), self.bc_diag, dev.ptr(x), self.n, st))
        check(lib.tg_axpy(dev.ptr(y), -self.bc_diag, dev.ptr(self._xm), self.n, st))
        return y

    def to_scipy(self, drop_eps=None):
        """Explicit matrix (checking only): host triple product of the downloaded operands."""
        import scipy.sparse as sp
        Ms = self.M.to_scipy()
        C = (Ms.T @ self.A.to_scipy() @ Ms).tocsr()
        if self.bc_mask is not None:
            z = np.flatnonzero(dev.to_np(self.bc_mask))
            keep = np.ones(self.n)
            keep[z] = 0.0
            D = sp.diags(keep)
            C = (D @ C @ D + sp.diags((1.0 - keep) * self.bc_diag)).tocsr()
        return C

    def solve(self, b, x0=None, rtol=1e-12, atol=0.0, maxit=100000):
        """CG on the operator (unpreconditioned: the diagonal of an unassembled triple product
        is not available)."""
        from . import solvers
        scratch = dev.empty(lib.tg_cg_scratch_len())
        s = dev.zeros(2)

        def spmv_dot(p, q):
            self.matvec(p, q)
            check(lib.tg_dot(dev.ptr(p), dev.ptr(q), self.n, dev.ptr(scratch), dev.ptr(s),
                             dev.stream()))
            return float(s[0].item())

        def ident(r, z):
            z.copy_(r)
            return z
        return solvers.pcg(spmv_dot, ident, b, x0, rtol, atol, maxit)


def carrier_splines(mesh, degree):
    """Maximal-continuity B-splines of the FE degree on the basis' extraction mesh: they only
    carry the mesh (unique knots) and the Q_p Lagrange tables into ``engine.TensorPatch``."""
    from .bsplines import BSpline1
    out = []
    for uk in mesh.uniqueKnots:
        uk = np.asarray(uk, dtype=np.float64)
        kn = np.concatenate([[uk[0]] * degree, uk, [uk[-1]] * degree])
        out.append(BSpline1(degree, kn))
    return out
