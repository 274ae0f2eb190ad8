"""
Oracle: the reference pipeline end to end on the CPU (numpy/scipy):
extract (M, cpFuncs) -> assemble (A_FE, b_FE) -> M^T A M, M^T b, BCs -> solve.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Also serves as the timed
CPU baseline ("port") for bench.py.

Follows common.py:1142-1160 (extractVector), :1176-1204 (extractMatrix:
PtAP then zeroRowsColumns(zeroDofs, diag)), :1236-1263 (solveLinearSystem;
DOLFIN default = sparse LU), :1266-1290 (driver).
"""
import time
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import bsplines as B
from . import extraction as X
from . import assembly as A


def ptap(Afe, M):
    """MatPtAP stand-in (common.py:1194-1195): row-wise CSR products."""
    C = (M.T @ Afe @ M).tocsr()
    C.sort_indices()
    return C


def apply_bcs_matrix(C, zeroDofs, diag=1.0):
    """zeroRowsColumns (common.py:1199-1200)."""
    C = C.tolil(copy=True)
    z = np.unique(np.asarray(zeroDofs, dtype=np.int64))
    C[z, :] = 0.0
    C[:, z] = 0.0
    C[z, z] = diag
    C = C.tocsr()
    return C


def apply_bcs_matrix_fast(C, zeroDofs, diag=1.0):
    """Same result as apply_bcs_matrix, vectorised (keeps the pattern)."""
    C = C.tocsr(copy=True)
    n = C.shape[0]
    mask = np.zeros(n, dtype=bool)
    mask[np.asarray(zeroDofs, dtype=np.int64)] = True
    rows = np.repeat(np.arange(n), np.diff(C.indptr))
    kill = mask[rows] | mask[C.indices]
    isdiag = (rows == C.indices) & mask[rows]
    C.data[kill] = 0.0
    C.data[isdiag] = diag
    return C


def apply_bcs_vector(b, zeroDofs):
    """common.py:1154-1158."""
    b = b.copy()
    b[np.asarray(zeroDofs, dtype=np.int64)] = 0.0
    return b


def jacobi_cg(C, b, rtol=1e-12, maxit=100000, x0=None):
    """Jacobi-preconditioned CG (the Krylov option of common.py:1257-1258)."""
    dinv = 1.0 / C.diagonal()
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = b - C @ x
    z = dinv * r
    p = z.copy()
    rz = r @ z
    bn = np.linalg.norm(b)
    it = 0
    if bn == 0.0:
        return x, 0, 0.0
    while it < maxit:
        Ap = C @ p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        it += 1
        rn = np.linalg.norm(r)
        if rn <= rtol * bn:
            break
        z = dinv * r
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, it, rn / bn


class Problem(object):
    """One synthetic tensor-product patch problem."""

    def __init__(self, degrees, kvecs, form="poisson", P=None, nLayers=1,
                 quadDeg=None, rationalize=False, extraDim=0):
        self.ts = B.TensorSpline(degrees, kvecs)
        self.form = form
        self.P = B.explicit_control_net(self.ts, extraDim) if P is None else np.asarray(P, float)
        self.quadDeg = 2 * max(degrees) if quadDeg is None else quadDeg
        self.nq = self.quadDeg // 2 + 1
        self.rationalize = rationalize
        z = []
        for d in range(self.ts.nvar):
            for side in (0, 1):
                z += self.ts.getSideDofs(d, side, nLayers)
        self.zeroDofs = z
        self.times = {}

    # -- stage 1: extraction (common.py:321-383)
    def extract(self):
        t = time.perf_counter()
        self.M = X.build_M_kron(self.ts)
        self.cpn = X.control_funcs(self.M, self.P)
        self.times["extract"] = time.perf_counter() - t

    # -- stage 2: FE assembly (common.py:1215-1216, 1169)
    def assemble(self, f):
        t = time.perf_counter()
        pf = self.ts.getDegree()
        nder = max(A.FORM_ORDER[self.form], 1)
        if self.rationalize and self.form == "biharmonic":
            nder = 2
        self.tabs_fe = [A.tab_fe(s, pf, self.nq, nder) for s in self.ts.splines]
        self.Afe, self.bfe = A.assemble(self.tabs_fe, self.cpn, self.form, f,
                                        rationalize=self.rationalize)
        self.times["assemble"] = time.perf_counter() - t

    # -- stage 3: extraction of the system (common.py:1142-1204)
    def ptap(self, applyBCs=True, diag=1.0):
        t = time.perf_counter()
        self.C0 = ptap(self.Afe, self.M)
        self.b0 = self.M.T @ self.bfe
        if applyBCs:
            self.C = apply_bcs_matrix_fast(self.C0, self.zeroDofs, diag)
            self.b = apply_bcs_vector(self.b0, self.zeroDofs)
        else:
            self.C, self.b = self.C0, self.b0
        self.times["ptap"] = time.perf_counter() - t

    # -- stage 4: solve (common.py:1236-1263)
    def solve(self, method="lu", rtol=1e-12):
        t = time.perf_counter()
        if method == "lu":
            self.U = spla.spsolve(self.C.tocsc(), self.b)
            self.iters = 0
        else:
            self.U, self.iters, self.relres = jacobi_cg(self.C, self.b, rtol)
        self.times["solve"] = time.perf_counter() - t
        return self.U

    def run(self, f, method="lu", rtol=1e-12):
        self.extract()
        self.assemble(f)
        self.ptap()
        return self.solve(method, rtol)

    # -- independent path: direct IGA Galerkin (SURVEY 8c KAT 3)
    def direct_iga(self, f):
        nder = max(A.FORM_ORDER[self.form], 1)
        if self.rationalize and self.form == "biharmonic":
            nder = 2
        tabs = [A.tab_iga(s, self.nq, nder) for s in self.ts.splines]
        return A.assemble(tabs, self.P, self.form, f, rationalize=self.rationalize)

    def error(self, U, kind, exact):
        """L2 / energy error of the IGA solution, integrated in the B-spline
        basis (equal to the FE-side integral of poisson.py:132)."""
        nder = 2 if kind == "energy" else 1
        tabs = [A.tab_iga(s, self.nq, nder) for s in self.ts.splines]
        return np.sqrt(A.functional(tabs, self.P, U, kind, exact,
                                    rationalize=self.rationalize))


class ElasticityProblem(object):
    """Equal-order multi-field patch problem: nsd displacement fields
    on the control mesh's spline (EqualOrderSpline, common.py:1891-1945).  The
    multi-field extraction operator (generateM, common.py:1546-1573) is block
    diagonal, ``I_nf (x) M_scalar``; IGA DoFs are field-major (globalDof,
    common.py:254-262); homogeneous BCs on the listed global DoFs
    (common.py:1154-1158, 1199-1200)."""

    def __init__(self, degrees, kvecs, P, mu, lam, zeroDofs, quadDeg=None):
        self.ts = B.TensorSpline(degrees, kvecs)
        self.P = np.asarray(P, float)
        self.nf = self.P.shape[1] - 1
        self.mu, self.lam = mu, lam
        self.quadDeg = 2 * max(degrees) if quadDeg is None else quadDeg
        self.nq = self.quadDeg // 2 + 1
        self.zeroDofs = np.unique(np.asarray(zeroDofs, dtype=np.int64))

    def fe_path(self, f):
        """Reference-faithful: A_FE, b_FE on the Lagrange mesh, then M^T A M, M^T b."""
        Ms = X.build_M_kron(self.ts)
        M = sp.kron(sp.identity(self.nf), Ms, format="csr")
        cpn = X.control_funcs(Ms, self.P)
        pf = self.ts.getDegree()
        tabs = [A.tab_fe(s, pf, self.nq, 1) for s in self.ts.splines]
        Afe, bfe = A.assemble_elasticity(tabs, cpn, self.mu, self.lam, f)
        return ptap(Afe, M), M.T @ bfe

    def direct_iga(self, f):
        tabs = [A.tab_iga(s, self.nq, 1) for s in self.ts.splines]
        return A.assemble_elasticity(tabs, self.P, self.mu, self.lam, f)

    def solve(self, f, direct=True):
        C0, b0 = self.direct_iga(f) if direct else self.fe_path(f)
        self.C = apply_bcs_matrix_fast(C0, self.zeroDofs, 1.0)
        self.b = apply_bcs_vector(b0, self.zeroDofs)
        self.U = spla.spsolve(self.C.tocsc(), self.b)
        return self.U
