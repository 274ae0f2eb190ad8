"""
Linear solvers behind ``ExtractedSpline.solveLinearSystem`` (common.py:1236-1263).

The reference hands the extracted system to ``dolfin.solve`` -- a sparse direct LU through
PETSc by default (common.py:1255-1256) -- or to a user ``linearSolver`` (PETSc KSP + PC).
Here the same role is played by three device solvers, all FP64, all on the C-ABI kernels:

* ``BandCholesky``  -- direct: blocked band Cholesky (tg_band.cu).  In the reference's DoF
  numbering the IGA matrix of a 2-D patch (or a small 3-D one) is banded with half-bandwidth
  ``sum_d p_d * stride_d``; used whenever the band fits (every 2-D BASELINE config: the
  biharmonic system has cond ~ h^-4 and defeats point-Jacobi CG).
* ``FastDiag`` + ``pcg`` -- CG preconditioned by fast diagonalisation of the tensor-product
  operator ``sigma M(x)M(x)M + sum_d c_d K_d (x) M (x) M`` on the parametric domain (exact inverse
  for an affine geometry, h- and p-robust otherwise): three dense mode products with the 1-D
  generalised eigenvectors, a scaling, three mode products back (tg_dgemm_batched,
  tg_fd_scale).  ``c_d`` and ``sigma`` are a least-squares fit to diag(C), so anisotropic /
  curved geometries are followed without looking at the form.
* Jacobi-CG (tg_win_solve_cg) -- what a user ``KrylovSolver("cg", "jacobi")`` asks for.

Host side: only 1-D quantities (the n_d x n_d generalised eigenproblems, like the Gauss rules
and knot tables elsewhere) and the CG scalars.
"""
import ctypes as C
import math
import os

import numpy as np

from . import dev
from ._lib import lib, check


class SolverBreakdown(RuntimeError):
    """p.Ap <= 0 or a non-positive pivot: the matrix is not symmetric positive definite."""


# ------------------------------------------------------------------------------------------
def _dir_matrices(D):
    """1-D mass and stiffness matrices of one parametric direction on the parametric domain,
    from the device tables of the hot path (exact: nq = p+1 Gauss points per span)."""
    t = D.table(1)
    np1, nq, nel = D.p + 1, D.nq, D.nel
    tab = dev.to_np(t["tabN"]).reshape(nel, nq, np1, 2)
    idx = dev.to_np(t["idxN"]).reshape(nel, np1).astype(np.int64)
    wq = dev.to_np(t["wq"]).reshape(nel, nq)
    Me = np.einsum("eq,eqi,eqj->eij", wq, tab[..., 0], tab[..., 0])
    Ke = np.einsum("eq,eqi,eqj->eij", wq, tab[..., 1], tab[..., 1])
    n = D.ncp
    M = np.zeros((n, n))
    K = np.zeros((n, n))
    I = np.repeat(idx[:, :, None], np1, axis=2)
    J = np.repeat(idx[:, None, :], np1, axis=1)
    np.add.at(M, (I, J), Me)
    np.add.at(K, (I, J), Ke)
    return M, K


_eig_cache = {}


def _gen_eig(D, free):
    """(lam, U) of K u = lam M u restricted to the ``free`` indices, embedded in n x n:
    constrained rows/columns of U are zero and carry lam = +inf."""
    key = (D.p, D.nq, tuple(np.asarray(D.s.knots, dtype=np.float64).tolist()), free.tobytes())
    hit = _eig_cache.get(key)
    if hit is not None:
        return hit
    import scipy.linalg as sla
    M, K = _dir_matrices(D)
    n = D.ncp
    f = np.flatnonzero(free)
    lam = np.full(n, np.inf)
    U = np.zeros((n, n))
    if f.size:
        w, V = sla.eigh(K[np.ix_(f, f)], M[np.ix_(f, f)])
        lam[f] = np.maximum(w, 0.0)
        U[np.ix_(f, f)] = V
    out = (lam, U, np.diag(M).copy(), np.diag(K).copy())
    if len(_eig_cache) > 16:
        _eig_cache.clear()
    _eig_cache[key] = out
    return out


def choose_fd_weights(sums, dim):
    """Direction weights (c_d, sigma) of the FD surrogate from the 15 sums of
    ``tg_fd_fit_rel`` (relative-error least squares against diag(C)).  Three nested models are
    fitted -- stiffness only, mass only, both -- and the simplest one whose residual is within
    a factor 2 of the best admissible (non-negative) one wins: on curved or rational geometry
    the full model buys a slightly smaller residual with a large spurious mass term that
    doubles the iteration count (annulus: 52 iterations against 20), while a genuine
    reaction term or a pure mass matrix leaves the stiffness-only model far behind."""
    sums = np.asarray(sums, dtype=np.float64)
    G = np.zeros((4, 4))
    k = 0
    for a in range(4):
        for b in range(a, 4):
            G[a, b] = G[b, a] = sums[k]
            k += 1
    r, N = sums[10:14], max(float(sums[14]), 1.0)

    def fit(idx):
        Gi = G[np.ix_(idx, idx)]
        try:
            x = np.linalg.solve(Gi, r[idx])
        except np.linalg.LinAlgError:
            return None
        if not np.all(np.isfinite(x)):
            return None
        res2 = max(float(N - 2.0 * x @ r[idx] + x @ Gi @ x), 0.0) / N
        full = np.zeros(4)
        full[idx] = x
        return full, res2
    cands = []
    for idx in (list(range(dim)), [3], list(range(dim)) + [3]):
        f = fit(idx)
        if f is None:
            continue
        x, res2 = f
        scale = max(np.abs(x).max(), 1e-300)
        if np.any(x < -1e-9 * scale) or not np.any(x > 0):
            continue
        if len(idx) == dim and np.any(x[:dim] <= 0):       # stiffness-only needs all directions
            continue
        cands.append((np.maximum(x, 0.0), res2))
    if not cands:
        return [1.0] * dim, 0.0
    best = min(c[1] for c in cands)
    for x, res2 in cands:                                    # simplest first
        if res2 <= 4.0 * best + 1e-20:
            return [float(v) for v in x[:dim]], float(x[3])
    return [1.0] * dim, 0.0


def diag_scale(eig, dinv, mask, c, sigma, nd, dim, offset_last=0):
    """S = diag sqrt(B_ii / C_ii) (``tg_fd_diag_scale``) for the DoFs of ``dinv`` (all of them,
    or a slab of the last direction starting at plane ``offset_last``); None when S = I to
    round-off (affine geometry: the surrogate is exact)."""
    d_md = [dev.from_np(eig[d][2]) for d in range(dim)]
    d_kd = [dev.from_np(eig[d][3]) for d in range(dim)]
    L = dim - 1

    def P(Lst, d):
        if d >= dim:
            return None
        return dev.ptr(Lst[d]) + (8 * offset_last if d == L else 0)
    dC = dinv.reciprocal()
    out = dev.empty(dC.numel())
    cc = list(c) + [0.0] * (3 - dim)
    check(lib.tg_fd_diag_scale(dev.ptr(dC), dev.ptr(mask) if mask is not None else None,
                               P(d_kd, 0), P(d_kd, 1), P(d_kd, 2), P(d_md, 0), P(d_md, 1),
                               P(d_md, 2), cc[0], cc[1], cc[2], float(sigma), nd[0], nd[1],
                               nd[2], dev.ptr(out), dev.stream()))
    # (the .item() below also keeps d_md / d_kd alive until the kernel has run)
    if float((out - 1.0).abs().max().item()) < 1e-9:
        return None
    return out


class FastDiag(object):
    """z = B^-1 r with B the tensor-product surrogate of the extracted operator."""

    def __init__(self, patch, mask=None, diag=1.0, dinv=None, weights=None):
        """mask: device 0/1 over the IGA DoFs (constrained = 1) or None; ``diag``: the value
        zeroRowsColumns put on constrained rows; dinv: device 1/diag(C) (for the fit of the
        direction weights) or None -> unit weights."""
        self.patch = patch
        self.dim = patch.dim
        self.n = patch.n_iga
        self.nd = list(patch.ncp) + [1] * (3 - patch.dim)
        self.mask = mask
        self.cinv = 1.0 / float(diag) if diag else 1.0
        free = self._free_planes(mask)
        self.free = free
        eig = [_gen_eig(D, free[d]) for d, D in enumerate(patch.dirs)]
        c, sigma = (weights if weights is not None else self._fit(eig, dinv))
        self.weights, self.sigma = c, sigma
        self.S = None
        if dinv is not None and weights is None and \
                os.environ.get("TIGAR_B200_FD_SCALE", "0") == "1":
            self.S = diag_scale(eig, dinv, self.mask, c, sigma, self.nd, dim=self.dim)
        self.lam = [dev.from_np(c[d] * eig[d][0]) for d in range(self.dim)]
        self.U = [dev.from_np(np.asfortranarray(eig[d][1]).T.copy()) for d in range(self.dim)]
        # from_np stores C-order; U^T in C order == U in column-major order
        self.t1 = dev.empty(self.n)
        self.t2 = dev.empty(self.n)

    def _free_planes(self, mask):
        """free[d][i] = False iff the whole hyperplane i_d = i is constrained."""
        out = []
        if mask is None:
            return [np.ones(n, dtype=bool) for n in self.patch.ncp]
        shape = tuple(reversed(self.patch.ncp))           # first direction fastest
        m = mask.view(shape) != 0
        for d in range(self.dim):
            ax = self.dim - 1 - d
            other = tuple(a for a in range(self.dim) if a != ax)
            full = m.all(dim=other) if other else m
            out.append(~dev.to_np(full).astype(bool))
        return out

    def _fit(self, eig, dinv):
        dim = self.dim
        if dinv is None:
            return [1.0] * dim, 0.0
        md = [eig[d][2] for d in range(dim)]
        kd = [eig[d][3] for d in range(dim)]
        d_md = [dev.from_np(a) for a in md]
        d_kd = [dev.from_np(a) for a in kd]
        P = lambda L, d: dev.ptr(L[d]) if d < dim else None
        # diag(C) = 1/dinv
        d = dinv.reciprocal()
        if os.environ.get("TIGAR_B200_FD_FIT", "rel") != "abs":
            scratch = dev.empty(15 * 64)
            out15 = dev.zeros(15)
            check(lib.tg_fd_fit_rel(dev.ptr(d), dev.ptr(self.mask) if self.mask is not None
                                    else None, P(d_kd, 0), P(d_kd, 1), P(d_kd, 2), P(d_md, 0),
                                    P(d_md, 1), P(d_md, 2), self.nd[0], self.nd[1], self.nd[2],
                                    dev.ptr(scratch), dev.ptr(out15), dev.stream()))
            return choose_fd_weights(dev.to_np(out15), dim)
        scratch = dev.empty(256)
        out4 = dev.zeros(4)
        check(lib.tg_fd_fit(dev.ptr(d), dev.ptr(self.mask) if self.mask is not None else None,
                            P(d_kd, 0), P(d_kd, 1), P(d_kd, 2), P(d_md, 0), P(d_md, 1),
                            P(d_md, 2), self.nd[0], self.nd[1], self.nd[2], dev.ptr(scratch),
                            dev.ptr(out4), dev.stream()))
        rhs = dev.to_np(out4)
        # Gram matrix of the Kronecker diagonals (separable sums over the free index sets)
        nb = dim + 1

        def f(a, dd):
            v = kd[dd] if a == dd else md[dd]
            return np.where(self.free[dd], v, 0.0)
        G = np.ones((nb, nb))
        for a in range(nb):
            for b in range(nb):
                for dd in range(dim):
                    G[a, b] *= float(np.dot(f(a, dd), f(b, dd)))
        r = np.concatenate([rhs[:dim], rhs[3:4]])
        try:
            sol = np.linalg.solve(G, r)
        except np.linalg.LinAlgError:
            return [1.0] * dim, 0.0
        c, sigma = list(sol[:dim]), float(sol[dim])
        if not all(np.isfinite(sol)) or min(c) <= 0.0:
            # not a second-order operator (mass matrix, ...): fit sigma alone or give up
            if all(abs(x) < 1e-8 * abs(sigma) for x in c) and sigma > 0:
                return [0.0] * dim, sigma
            return [1.0] * dim, 0.0
        cmax = max(c)
        return [float(x) for x in c], max(sigma, 0.0) if sigma > -1e-10 * cmax else 0.0

    def _gemm(self, ta, tb, M, N, K, A, lda, sA, B, ldb, sB, Cc, ldc, sC, batch):
        check(lib.tg_dgemm_batched(ta, tb, M, N, K, 1.0, A, lda, sA, B, ldb, sB, 0.0, Cc, ldc, sC,
                                   batch, dev.stream()))

    def apply(self, r, z):
        n0, n1, n2 = self.nd
        dim = self.dim
        st = dev.stream()
        t1, t2 = dev.ptr(self.t1), dev.ptr(self.t2)
        U = [dev.ptr(u) for u in self.U]
        check(lib.tg_masked_copy(t1, dev.ptr(r), dev.ptr(self.mask) if self.mask is not None
                                 else None, self.n, st))
        if self.S is not None:
            check(lib.tg_vmul(t1, t1, dev.ptr(self.S), self.n, st))
        a, b = t1, t2
        # forward: U_d^T along every direction
        self._gemm(1, 0, n0, n1 * n2, n0, U[0], n0, 0, a, n0, 0, b, n0, 0, 1)
        a, b = b, a
        if dim > 1:
            self._gemm(0, 0, n0, n1, n1, a, n0, n0 * n1, U[1], n1, 0, b, n0, n0 * n1, n2)
            a, b = b, a
        if dim > 2:
            self._gemm(0, 0, n0 * n1, n2, n2, a, n0 * n1, 0, U[2], n2, 0, b, n0 * n1, 0, 1)
            a, b = b, a
        L = [dev.ptr(l) for l in self.lam] + [None] * (3 - dim)
        check(lib.tg_fd_scale(a, L[0], L[1], L[2], n0, n1, n2, 0, n0 * n1, self.sigma, 1, st))
        # backward: U_d
        if dim > 2:
            self._gemm(0, 1, n0 * n1, n2, n2, a, n0 * n1, 0, U[2], n2, 0, b, n0 * n1, 0, 1)
            a, b = b, a
        if dim > 1:
            self._gemm(0, 1, n0, n1, n1, a, n0, n0 * n1, U[1], n1, 0, b, n0, n0 * n1, n2)
            a, b = b, a
        self._gemm(0, 0, n0, n1 * n2, n0, U[0], n0, 0, a, n0, 0, dev.ptr(z), n0, 0, 1)
        if self.S is not None:
            check(lib.tg_vmul(dev.ptr(z), dev.ptr(z), dev.ptr(self.S), self.n, st))
        if self.mask is not None:
            check(lib.tg_masked_fix(dev.ptr(z), dev.ptr(r), dev.ptr(self.mask), self.cinv,
                                    self.n, st))
        return z

    @property
    def flops_per_apply(self):
        n0, n1, n2 = self.nd
        tot = n0 * n1 * n2
        return 4.0 * tot * sum(self.nd[:self.dim])


# ------------------------------------------------------------------------------------------
class _Vec(object):
    """Scalar helpers of the CG driver on one GPU (host-read dot products)."""

    def __init__(self, n):
        self.n = n
        self.scratch = dev.empty(lib.tg_cg_scratch_len())
        self.s = dev.zeros(4)
        self.reduce_dev = None

    def dot(self, a, b):
        check(lib.tg_dot(dev.ptr(a), dev.ptr(b), self.n, dev.ptr(self.scratch), dev.ptr(self.s),
                         dev.stream()))
        if self.reduce_dev is not None:          # all-reduce on the device, one host read
            self.reduce_dev(self.s[0:1])
            return float(self.s[0].item())
        return self.reduce(float(self.s[0].item()))

    def reduce(self, v):
        return v


def pcg(spmv_dot, precond, b, x0=None, rtol=1e-12, atol=0.0, maxit=10000, reduce=None,
        p_buf=None, reduce_dev=None):
    """Preconditioned CG.  ``spmv_dot(p, q) -> p.q`` computes q = A p and returns the (global)
    dot product; ``precond(r, z)`` writes z = B^-1 r; ``reduce`` sums a host scalar over ranks;
    ``p_buf``: where the search direction lives (a view into the halo-extended vector of a
    row-distributed matrix).  Returns (x, iterations, relative residual).  Raises
    SolverBreakdown when p.Ap <= 0."""
    n = b.numel()
    V = _Vec(n)
    if reduce is not None:
        V.reduce = reduce
    V.reduce_dev = reduce_dev
    st = dev.stream
    x = dev.zeros(n) if x0 is None else x0
    r = b.clone()
    q = dev.empty(n)
    z = dev.empty(n)
    bb = V.dot(b, b)
    if x0 is not None:
        spmv_dot(x, q)
        check(lib.tg_axpy(dev.ptr(r), -1.0, dev.ptr(q), n, st()))
    rr = V.dot(r, r)
    tol2 = max(rtol * rtol * bb, atol * atol)
    it = 0
    if rr <= tol2:
        return x, 0, (math.sqrt(rr / bb) if bb > 0 else 0.0)
    precond(r, z)
    if p_buf is None:
        p = z.clone()
    else:
        p = p_buf
        p.copy_(z)
    rz = V.dot(r, z)
    out1 = V.s[1:2]
    while rr > tol2 and it < maxit:
        pAp = spmv_dot(p, q)
        if not (pAp > 0.0):
            if pAp == 0.0 and rz == 0.0:
                break
            raise SolverBreakdown("CG breakdown at iteration %d: p.Ap = %g (matrix not symmetric "
                                  "positive definite)" % (it, pAp))
        alpha = rz / pAp
        check(lib.tg_pcg_update(dev.ptr(x), dev.ptr(r), dev.ptr(p), dev.ptr(q), alpha, n,
                                dev.ptr(V.scratch), dev.ptr(out1), st()))
        if reduce_dev is not None:
            reduce_dev(out1)
            rr = float(out1[0].item())
        else:
            rr = V.reduce(float(out1[0].item()))
        it += 1
        if not (rr == rr):
            raise FloatingPointError("CG produced NaN at iteration %d" % it)
        if rr <= tol2:
            break
        precond(r, z)
        rz_new = V.dot(r, z)
        check(lib.tg_xpby(dev.ptr(p), rz_new / rz, dev.ptr(z), n, st()))
        rz = rz_new
    return x, it, (math.sqrt(rr / bb) if bb > 0 else 0.0)


def solve_fd_pcg(patch, Cm, b, mask=None, diag=1.0, x0=None, rtol=1e-12, atol=0.0, maxit=10000):
    """FD-preconditioned CG on a windowed system matrix (one GPU)."""
    n = Cm.window.nrows
    dinv = dev.empty(n)
    check(lib.tg_win_diag_inv(Cm.window.ref(), dev.ptr(Cm.vals), 0, dev.ptr(dinv), dev.stream()))
    fd = FastDiag(patch, mask, diag, dinv)
    del dinv
    scratch = dev.empty(lib.tg_cg_scratch_len())
    s = dev.zeros(2)

    W = Cm.window
    spmv_bytes = 8 * W.nnz + 24 * W.nrows

    def spmv_dot(p, q):
        with dev.PROF.range("k_win_spmv (CG matvec + p.Ap)", spmv_bytes, 2.0 * W.nnz):
            check(lib.tg_win_spmv_dot(W.ref(), dev.ptr(Cm.vals), dev.ptr(p), 0, dev.ptr(q),
                                      dev.ptr(scratch), dev.ptr(s), dev.stream()))
        return float(s[0].item())

    def precond(r, z):
        with dev.PROF.range("k_dgemm x6 + k_fd_scale (FD preconditioner)", 16 * 8 * n,
                            fd.flops_per_apply):
            return fd.apply(r, z)
    x, its, rel = pcg(spmv_dot, precond, b, x0, rtol, atol, maxit)
    return x, its, rel, fd


# ------------------------------------------------------------------------------------------
class BandCholesky(object):
    """Direct solve of a windowed SPD system: band conversion, factorisation, solves."""

    NB = 32

    def __init__(self, Cm):
        w = Cm.window
        if w.layout != 0:
            raise ValueError("band solver needs the row-major window layout")
        self.n = w.nrows
        self.bw = self.bandwidth(w)
        self.ldab = self.bw + self.NB + ((self.bw + self.NB) & 1)
        self.Cm = Cm
        self.AB = None

    @staticmethod
    def bandwidth(w):
        stride, bw = 1, 0
        for d in range(w.dim):
            r = np.arange(w.nr[d], dtype=np.int64)
            reach = int(max((r - w.lo[d].astype(np.int64)).max(),
                            (w.hi[d].astype(np.int64) - r).max()))
            bw += reach * stride
            stride *= w.nc[d]
        return bw

    @classmethod
    def cost(cls, w):
        """(bytes of band storage, flops of the factorisation)."""
        n, bw = w.nrows, cls.bandwidth(w)
        return 8.0 * n * (bw + cls.NB + 1), float(n) * (bw + cls.NB) ** 2

    def factor(self, sym_tol=1e-9):
        Cm, w = self.Cm, self.Cm.window
        o2 = dev.zeros(2)
        check(lib.tg_win_asym(w.ref(), dev.ptr(Cm.vals), dev.ptr(o2), dev.stream()))
        asym, amax = o2.tolist()
        if asym > sym_tol * max(amax, 1e-300):
            raise SolverBreakdown("matrix is not symmetric (max |A - A^T| = %.3e, max |A| = %.3e): "
                                  "the Cholesky/CG solvers need a symmetric form" % (asym, amax))
        self.AB = dev.zeros(self.ldab * self.n)
        info = dev.zeros(1, dev.I32)
        check(lib.tg_band_from_win(w.ref(), dev.ptr(Cm.vals), self.bw, self.ldab, dev.ptr(self.AB),
                                   dev.ptr(info), 1, 0, 0, dev.stream()))
        check(lib.tg_band_cholesky(self.n, self.bw, self.ldab, dev.ptr(self.AB), dev.ptr(info),
                                   dev.stream()))
        code = int(info.item())
        if code != 0:
            self.AB = None
            raise SolverBreakdown("band Cholesky: %s" % (
                "non-zero outside the computed band" if code < 0 else
                "pivot block at row %d is not positive definite" % (code - 1)))
        return self

    def solve(self, b):
        if self.AB is None:
            self.factor()
        work = dev.empty(self.n)
        rhs = b.clone()
        x = dev.empty(self.n)
        check(lib.tg_band_solve(self.n, self.bw, self.ldab, dev.ptr(self.AB), dev.ptr(rhs),
                                dev.ptr(work), dev.ptr(x), dev.stream()))
        return x


class BlockBandCholesky(object):
    """Direct solve of an equal-order multi-field system (BlockMatrix of windowed blocks on one
    pattern): the blocks are interleaved node-major (row' = nf*row + field), which keeps the
    band at nf*bw + nf - 1, and factored by the same band Cholesky.  This is what makes the
    Newton steps of the Kirchhoff-Love shell (4th order, cond ~ h^-4) solvable: the reference
    uses a direct LU there as well (common.py:1255-1256)."""

    def __init__(self, Bm):
        self.Bm = Bm
        self.nf, self.nb = Bm.nf, Bm.n
        any_block = next(iter(Bm.blocks.values()))
        self.w = any_block.window
        self.n = self.nf * self.nb
        self.bw = self.nf * BandCholesky.bandwidth(self.w) + self.nf - 1
        self.ldab = self.bw + BandCholesky.NB + ((self.bw + BandCholesky.NB) & 1)
        self.AB = None

    @classmethod
    def cost(cls, Bm):
        w = next(iter(Bm.blocks.values())).window
        n = Bm.nf * Bm.n
        bw = Bm.nf * BandCholesky.bandwidth(w) + Bm.nf - 1
        return 8.0 * n * (bw + BandCholesky.NB + 1), float(n) * (bw + BandCholesky.NB) ** 2

    def factor(self, matvec=None, sym_tol=1e-9):
        import torch
        if matvec is not None:                  # symmetry test: x.Ay = y.Ax for random x, y
            g = torch.Generator(device=dev.device()).manual_seed(1)
            x = torch.rand(self.n, dtype=torch.float64, device=dev.device(), generator=g)
            y = torch.rand(self.n, dtype=torch.float64, device=dev.device(), generator=g)
            ax, ay = dev.empty(self.n), dev.empty(self.n)
            matvec(x, ax)
            matvec(y, ay)
            s1, s2 = float(torch.dot(y, ax)), float(torch.dot(x, ay))
            if abs(s1 - s2) > sym_tol * max(abs(s1), abs(s2), 1e-300):
                raise SolverBreakdown("block matrix is not symmetric (y.Ax = %.12e, x.Ay = %.12e)"
                                      % (s1, s2))
        self.AB = dev.zeros(self.ldab * self.n)
        info = dev.zeros(1, dev.I32)
        for (f, g_), B in sorted(self.Bm.blocks.items()):
            check(lib.tg_band_from_win(B.window.ref(), dev.ptr(B.vals), self.bw, self.ldab,
                                       dev.ptr(self.AB), dev.ptr(info), self.nf, f, g_,
                                       dev.stream()))
        check(lib.tg_band_cholesky(self.n, self.bw, self.ldab, dev.ptr(self.AB), dev.ptr(info),
                                   dev.stream()))
        code = int(info.item())
        if code != 0:
            self.AB = None
            raise SolverBreakdown("band Cholesky (block system): %s" % (
                "non-zero outside the computed band" if code < 0 else
                "pivot block at row %d is not positive definite" % (code - 1)))
        return self

    def solve(self, b):
        """b, result: field-major (globalDof numbering, common.py:254-262)."""
        rhs = b.view(self.nf, self.nb).t().contiguous().view(-1)        # -> node-major
        work, x = dev.empty(self.n), dev.empty(self.n)
        check(lib.tg_band_solve(self.n, self.bw, self.ldab, dev.ptr(self.AB), dev.ptr(rhs),
                                dev.ptr(work), dev.ptr(x), dev.stream()))
        return x.view(self.nb, self.nf).t().contiguous().view(-1)


def _affordable(nbytes, flops, free_bytes):
    lim_gb = float(os.environ.get("TIGAR_B200_DIRECT_GB", "12"))
    lim_flops = float(os.environ.get("TIGAR_B200_DIRECT_FLOPS", "3e13"))
    if free_bytes is not None and nbytes > 0.5 * free_bytes:
        return False
    return nbytes <= lim_gb * 2 ** 30 and flops <= lim_flops


def direct_affordable(w, free_bytes=None):
    """Policy: use the band solver when its storage and work are small next to the device."""
    nbytes, flops = BandCholesky.cost(w)
    return _affordable(nbytes, flops, free_bytes)
