"""
Equal-order multi-field systems.

The reference builds a ``MixedElement`` FE space with one sub-element per field
(common.py:337-351) and fills ``M`` field by field, the IGA column of field ``f``
being offset by the control-point counts of the earlier fields
(``globalDof``, common.py:254-262; ``generateM``, common.py:1546-1573).  With every
field on the control mesh's scalar spline (``EqualOrderSpline``, common.py:1891-1945)
that ``M`` is block diagonal with the scalar extraction operator in every block, so

    M^T A M = [ M_s^T A_fg M_s ]_{f,g}          M^T b = [ M_s^T b_f ]_f

and the IGA system is an nf x nf grid of blocks that all live on the scalar
problem's window pattern.  Every block is produced by the scalar hot path
(Gauss-point pass -> sum-factorised assembly -> M^T A M, or the element-fused
variant); this module holds the host-side bookkeeping around it:

  * splitting a form's terms by (test field, trial field),
  * ``BlockMatrix`` -- the grid of windowed blocks, homogeneous BCs per block
    (zeroRowsColumns with a row mask of the test field and a column mask of the
    trial field, common.py:1199-1200),
  * ``BlockOps`` -- the building blocks the Jacobi-CG driver ``multigpu.dist_cg``
    calls, bound to the C-ABI (one windowed SpMV per block, ``tg_axpy`` to sum a
    block row); ``tests/test_multifield_cpu.py`` runs the same driver on numpy.

IGA DoF numbering is the reference's: field-major, ``globalDof(f, i) = f*ncp + i``.
Host logic only; no numerical fallback lives here.
"""
import numpy as np


def part_field(part):
    """Field of one half of a term key: ``None`` (no basis function), a 3-multi-index
    (single-field space: field 0) or ``(a0, a1, a2, field)``."""
    if part is None:
        return None
    return part[3] if len(part) == 4 else 0


def part_alpha(part):
    return None if part is None else tuple(part[:3])


def split_matrix_terms(terms, nf):
    """{(test, trial): node} -> {(f, g): {(alphaTest, alphaTrial): node}}."""
    out = {}
    for (t, u), node in terms.items():
        f, g = part_field(t), part_field(u)
        if f is None or g is None:
            raise ValueError("bilinear form expected")
        if not (0 <= f < nf and 0 <= g < nf):
            raise IndexError("field index outside the function space")
        blk = out.setdefault((f, g), {})
        k = (part_alpha(t), part_alpha(u))
        blk[k] = blk[k] + node if k in blk else node
    return out


def split_vector_terms(terms, nf):
    """{(test, None): node} -> {f: {alphaTest: node}}."""
    out = {}
    for (t, u), node in terms.items():
        if u is not None or t is None:
            raise ValueError("linear form must be linear in the TEST function")
        f = part_field(t)
        if not 0 <= f < nf:
            raise IndexError("field index outside the function space")
        blk = out.setdefault(f, {})
        k = part_alpha(t)
        blk[k] = blk[k] + node if k in blk else node
    return out


def split_zero_dofs(zeroDofs, nf, n):
    """Global zero DoFs (field-major) -> per-field local lists."""
    z = np.asarray(zeroDofs, dtype=np.int64).ravel()
    if z.size and (z.min() < 0 or z.max() >= nf * n):
        raise IndexError("zero DoF outside [0, %d)" % (nf * n))
    return [z[(z >= f * n) & (z < (f + 1) * n)] - f * n for f in range(nf)]


class BlockMatrix(object):
    """nf x nf grid of windowed blocks on one window pattern; ``blocks[(f, g)]`` is a
    ``WinMatrix`` or absent (structurally zero)."""

    def __init__(self, nf, blocks, n):
        self.nf = nf
        self.blocks = dict(blocks)
        self.n = n                         # rows (= columns) per field

    @property
    def shape(self):
        return (self.nf * self.n, self.nf * self.n)

    def block(self, f, g):
        return self.blocks.get((f, g))

    def to_scipy(self, drop_eps=None):
        import scipy.sparse as sp
        grid = [[None] * self.nf for _ in range(self.nf)]
        for (f, g), B in self.blocks.items():
            grid[f][g] = B.to_scipy(drop_eps)
        for f in range(self.nf):
            if all(b is None for b in grid[f]):
                grid[f][f] = sp.csr_matrix((self.n, self.n))
        for g in range(self.nf):
            if all(grid[f][g] is None for f in range(self.nf)):
                grid[g][g] = sp.csr_matrix((self.n, self.n))
        return sp.bmat(grid, format="csr")


class BlockOps(object):
    """``ops`` interface of ``multigpu.dist_cg`` for a BlockMatrix on one GPU.
    Vectors are field-major concatenations; the scalar reductions and vector
    updates run once over the whole vector, the matvec block by block."""

    def __init__(self, Cm):
        from . import dev
        from ._lib import lib, check
        self.dev, self.lib, self.check = dev, lib, check
        self.Cm = Cm
        self.nf, self.nb = Cm.nf, Cm.n
        self.n = Cm.nf * Cm.n

    def _sp(self, i):
        return self.dev.ptr(self.s) + 8 * i

    def _fld(self, t, f):
        return self.dev.ptr(t) + 8 * f * self.nb

    def begin(self, b):
        dev, lib, check = self.dev, self.lib, self.check
        n = self.n
        self.b = b
        self.x = dev.zeros(n)
        self.r = dev.empty(n)
        self.q = dev.zeros(n)
        self.p = dev.zeros(n)
        self.tmp = dev.empty(self.nb)
        self.dinv = dev.empty(n)
        self.scratch = dev.empty(lib.tg_cg_scratch_len())
        self.s = dev.zeros(8)          # 0 rz_a, 1 rr_a, 2 pAp, 3 rz_b, 4 rr_b, 5 bb
        self.flip = 0
        for f in range(self.nf):
            D = self.Cm.block(f, f)
            if D is None:
                raise ValueError("Jacobi-CG needs the diagonal block of field %d" % f)
            check(lib.tg_win_diag_inv(D.window.ref(), dev.ptr(D.vals), 0,
                                      self._fld(self.dinv, f), dev.stream()))

    def dot_bb(self):
        dev, lib = self.dev, self.lib
        self.check(lib.tg_dot(dev.ptr(self.b), dev.ptr(self.b), self.n, dev.ptr(self.scratch),
                              self._sp(5), dev.stream()))
        return self.s[5:6]

    def allreduce_host(self, t):
        return float(t[0].item())

    def init_residual(self):
        dev, lib = self.dev, self.lib
        # x0 = 0 -> y = A x0 = 0 (self.q is zero-initialised)
        self.check(lib.tg_cg_init(dev.ptr(self.b), dev.ptr(self.q), dev.ptr(self.dinv),
                                  dev.ptr(self.r), dev.ptr(self.p), self.n,
                                  dev.ptr(self.scratch), self._sp(0), dev.stream()))
        h = self.s[0:2].tolist()
        return h[0], h[1]

    def exchange_halo(self):
        pass

    def matvec(self, x, y):
        """y = C x, block row by block row."""
        dev, lib, check = self.dev, self.lib, self.check
        st = dev.stream()
        for f in range(self.nf):
            first = True
            for g in range(self.nf):
                B = self.Cm.block(f, g)
                if B is None:
                    continue
                if first:
                    check(lib.tg_win_spmv(B.window.ref(), dev.ptr(B.vals), self._fld(x, g),
                                          self._fld(y, f), st))
                    first = False
                else:
                    check(lib.tg_win_spmv(B.window.ref(), dev.ptr(B.vals), self._fld(x, g),
                                          dev.ptr(self.tmp), st))
                    check(lib.tg_axpy(self._fld(y, f), 1.0, dev.ptr(self.tmp), self.nb, st))
            if first:
                raise ValueError("empty block row %d" % f)

    def spmv_dot(self):
        dev, lib = self.dev, self.lib
        self.matvec(self.p, self.q)
        self.check(lib.tg_dot(dev.ptr(self.p), dev.ptr(self.q), self.n, dev.ptr(self.scratch),
                              self._sp(2), dev.stream()))

    def axpy_dot(self):
        dev, lib = self.dev, self.lib
        cur, nxt = (0, 3) if self.flip == 0 else (3, 0)
        self.check(lib.tg_cg_axpy_dot(dev.ptr(self.x), dev.ptr(self.r), dev.ptr(self.p),
                                      dev.ptr(self.q), dev.ptr(self.dinv), self.n,
                                      self._sp(cur), self._sp(2), dev.ptr(self.scratch),
                                      self._sp(nxt), dev.stream()))

    def update_p(self):
        dev, lib = self.dev, self.lib
        cur, nxt = (0, 3) if self.flip == 0 else (3, 0)
        self.check(lib.tg_cg_xpby(dev.ptr(self.p), dev.ptr(self.r), dev.ptr(self.dinv), self.n,
                                  self._sp(nxt), self._sp(cur), dev.stream()))
        self.flip ^= 1

    def read_rz_rr(self):
        cur = 0 if self.flip == 0 else 3
        h = self.s[cur:cur + 2].tolist()
        return h[0], h[1]

    def solution(self):
        return self.x


def solve_block_cg(Cm, b, rtol=1e-12, atol=0.0, maxit=100000, check_every=5):
    """Jacobi-CG on a BlockMatrix (same driver as the row-distributed solver)."""
    from .multigpu import dist_cg
    return dist_cg(BlockOps(Cm), b, rtol, atol, maxit, check_every)


def solve_block(Cm, b, rtol=1e-12, atol=0.0, maxit=100000, method="auto"):
    """solve() of common.py:1255-1258 for an equal-order multi-field system.  "auto": the
    reference's default is a direct LU -> band Cholesky on the node-major interleaved system
    while it is affordable (with iterative refinement), Jacobi-CG beyond.
    Returns (x, iterations, relative residual, method)."""
    import os
    import torch
    from . import dev, solvers
    method = os.environ.get("TIGAR_B200_SOLVER", method)
    if method in ("auto", "fd"):
        free, _ = torch.cuda.mem_get_info()
        nbytes, flops = solvers.BlockBandCholesky.cost(Cm)
        method = "direct" if solvers._affordable(nbytes, flops, free) else "jacobi"
    if method == "jacobi":
        x, its, rel = solve_block_cg(Cm, b, rtol, atol, maxit)
        return x, its, rel, "jacobi"
    ops = BlockOps(Cm)
    ops.tmp = dev.empty(Cm.n)
    bc = solvers.BlockBandCholesky(Cm).factor(matvec=ops.matvec)
    x = bc.solve(b)
    bb = float(b.norm())
    r = dev.empty(b.numel())
    ops.matvec(x, r)
    r.neg_().add_(b)
    rel = float(r.norm()) / bb if bb > 0 else 0.0
    for _ in range(3):                       # iterative refinement (see TensorPatch.solve)
        if rel < 1e-15:
            break
        x2 = bc.solve(r).add_(x)
        r2 = dev.empty(b.numel())
        ops.matvec(x2, r2)
        r2.neg_().add_(b)
        rel2 = float(r2.norm()) / bb if bb > 0 else 0.0
        if not rel2 < rel:
            break
        x, r, rel = x2, r2, rel2
    return x, 1, rel, "direct"
