"""
Matrix-free IGA operator (SURVEY 7.2 hard part 1): ``C = M^T A_FE M`` of a 3-D cubic
patch needs 4 kB per DoF (557 GB at 512^3), so beyond ~300^3 cells on one GPU the system
matrix cannot exist.  The action of the operator, however, is a linear form,

    (C x)_i = a(x_h, N_i) ,   x_h = sum_j x_j N_j ,

i.e. exactly what ``assembleVector`` computes for a residual that contains a Function
(the Newton path, common.py:1304-1348): one Gauss-point pass that interpolates the jets
of ``x_h`` and evaluates the flux coefficients, followed by the sum-factorised vector
assembly.  ``FormOperator`` therefore needs NO new kernel: it turns the bilinear form's
terms ``c * D^a v * D^b u`` into the linear-form terms ``(c * D^b x_h) * D^a v`` once,
keeps the compiled program, and runs the two verified kernels per application.  It is
~18x slower per CG iteration than the windowed SpMV where the matrix fits (measured
shares at 256^3: Gauss-point pass 0.10 s + vector assembly 0.08 s against 10 ms), and the
only option where it does not.

Homogeneous BCs (common.py:1199-1200, 1154-1158) are applied in operator form,
``P C P + diag (I - P)`` with ``P`` the projector that zeroes the constrained DoFs; inside
CG every vector has zeros there (b is masked), so the application reduces to
``y = C p; y[zeroDofs] = 0``.  The Jacobi diagonal is taken from the assembled matrix one
row slab at a time (the slab partition of multigpu.py, run sequentially on this GPU), so
that only ``1/nparts`` of the matrix exists at any moment.
"""
import math
import os

from . import symbolic as S
from .multifield import BlockOps


class FormOperator(object):
    """``assembleMatrix(form)`` in matrix-free mode: applies M^T A M without forming it."""

    def __init__(self, spline, mterms, applyBCs=True, diag=1.0):
        from . import dev
        from .api import Function
        self.spline = spline
        self.patch = spline._patch
        self.n = self.patch.n_iga
        self.mterms = dict(mterms)             # {(alphaTest, alphaTrial): node}, weighted
        self.applyBCs = applyBCs
        self.diag = float(diag)
        self.xvec = dev.zeros(self.n)          # the operand lives here (aliased by CG's p)
        self.x = Function(spline.V)
        self.x.set_iga(self.xvec)
        vt = {}
        for (aT, aU), node in self.mterms.items():
            k = tuple(aT[:3])
            term = S.mul(node, S.jet(self.x.fid, 0, tuple(aU[:3])))
            vt[k] = S.add(vt[k], term) if k in vt else term
        self.vterms = vt
        self._cache = {}
        self._funcs = spline._funcs("iga")
        self._dinv = None

    @property
    def shape(self):
        return (self.n, self.n)

    def set_diagonal(self, d):
        """diag(C) computed elsewhere (engine.assemble_vector_and_diag: same Gauss-point pass
        as the load vector); a zero diagonal acts as 1 (as tg_win_diag_inv)."""
        d[d == 0.0] = 1.0
        self._dinv = d.reciprocal_()

    def apply(self, y):
        """y = C * self.xvec (no BCs)."""
        y.zero_()
        if os.environ.get("TIGAR_B200_MF_FUSED", "1") == "1":
            return self._apply_fused(y)
        self.patch.assemble_vector(self.vterms, self._funcs, "iga", out=y, cache=self._cache)
        return y

    def _apply_fused(self, y):
        """One generated kernel per colour instead of Gauss-point pass + vector assembly:
        the flux coefficients never leave the SM (jit.generate(..., op=...); numerics
        checked on host threads in tests/test_jit_emulated_cpu.py and, through this very
        glue, in tests/test_scalar_glue_cpu.py).  Default of the matrix-free mode;
        ``TIGAR_B200_MF_FUSED=0`` selects the two-kernel path above."""
        from . import dev, jit
        K = self._cache.get("fused")
        if K is None:
            p = self.patch
            keys = sorted(self.vterms)
            prog = S.compile_program([self.vterms[k] for k in keys], p.dim)
            fids = sorted(set(j[0] for j in prog.jets))
            if len(fids) > jit.MAXFUN:
                raise ValueError("too many coefficient functions for the generated kernel")
            jets = [(fids.index(f), c, tuple(al) + (0,) * (3 - len(al))) for (f, c, al) in prog.jets]
            nder = max([max(al) for (_, _, al) in prog.jets] + [max(k) for k in keys])
            B = p.basis("iga", nder)
            nloc = list(B.nloc) + [1] * (3 - p.dim)
            nq = [int(B.c.nq[d]) for d in range(3)]
            kern = jit.get_op_kernel(prog, p.dim, nloc, nq, B.nder + 1, jets, len(fids), keys)
            K = dict(kern=kern, B=B, fids=fids, stride=list(B.nloc))
            self._cache["fused"] = K
        ptrs = [dev.ptr(self._funcs[f]) for f in K["fids"]]
        jit.launch_op(K["kern"], K["B"], ptrs, y, K["stride"])
        return y

    def matvec(self, x, y=None):
        """y = (P C P + diag (I - P)) x for an arbitrary x (copies x into the operand)."""
        from . import dev
        from ._lib import lib, check
        y = dev.empty(self.n) if y is None else y
        self.xvec.copy_(x)
        mask = self.spline._bc_mask() if self.applyBCs else None
        if mask is not None:
            check(lib.tg_zero_entries(dev.ptr(self.xvec), dev.ptr(mask), self.n, dev.stream()))
        self.apply(y)
        if mask is not None:
            check(lib.tg_zero_entries(dev.ptr(y), dev.ptr(mask), self.n, dev.stream()))
            # + diag (I - P) x = diag * (x - P x)
            check(lib.tg_axpy(dev.ptr(y), self.diag, dev.ptr(x), self.n, dev.stream()))
            check(lib.tg_axpy(dev.ptr(y), -self.diag, dev.ptr(self.xvec), self.n, dev.stream()))
        return y

    # -- Jacobi diagonal -----------------------------------------------------------
    def slab_count(self, budget_bytes=24 << 30):
        nnz = self.patch.window("C").nnz
        return max(1, int(math.ceil(8.0 * nnz / budget_bytes)))

    def jacobi_dinv(self, nparts=None):
        """1 / diag(C), from the matrix assembled one row slab at a time."""
        if self._dinv is not None:
            return self._dinv
        from . import dev
        from ._lib import lib, check
        from .engine import TensorPatch
        p = self.patch
        if os.environ.get("TIGAR_B200_MF_FUSED", "1") == "1" and nparts is None:
            self._dinv = self._dinv_generated()
            return self._dinv
        nparts = self.slab_count() if nparts is None else int(nparts)
        if nparts > 1:                       # slabs at least p + 1 cell layers thick
            nparts = max(1, min(nparts, p.nel[-1] // (p.degrees[-1] + 1)))
        dinv = dev.empty(self.n)
        if nparts == 1:
            Cm = p.assemble_matrix(self.mterms, self._funcs, "iga")
            check(lib.tg_win_diag_inv(Cm.window.ref(), dev.ptr(Cm.vals), 0, dev.ptr(dinv),
                                      dev.stream()))
            del Cm
        else:
            for r in range(nparts):
                sub = TensorPatch(p.degrees, None, quadDeg=p.quadDeg, eps=p.eps,
                                  splines=p.splines, part=(r, nparts))
                Cm = sub.assemble_matrix(self.mterms, self._funcs, "iga")
                pp = sub.pp
                check(lib.tg_win_diag_inv(Cm.window.ref(), dev.ptr(Cm.vals),
                                          pp["k0"] - pp["c0"],
                                          dev.ptr(dinv) + 8 * sub.plane * pp["k0"], dev.stream()))
                dev.sync()
                del Cm, sub
        # constrained rows: their dinv is never used (r and p are zero there inside CG)
        self._dinv = dinv
        return dinv


    def _dinv_generated(self):
        """1 / diag(C) from a generated kernel that sums the diagonal of the element matrices
        (jit.generate(..., diag=True)): no matrix, not even slab-wise.  Same switch as
        ``_apply_fused``; numerics checked on host threads (tests/test_jit_emulated_cpu.py)."""
        from . import dev, jit
        p = self.patch
        keys = sorted(self.mterms)
        # (s, t) and (t, s) with the same coefficient contribute the same diagonal: one slot
        groups = {}
        for k in keys:
            a, b = tuple(k[0][:3]), tuple(k[1][:3])
            groups.setdefault((min(a, b), max(a, b), self.mterms[k].uid), []).append(k)
        outs, pairs = [], []
        for (a, b, _), ks in sorted(groups.items()):
            node = self.mterms[ks[0]]
            outs.append(node if len(ks) == 1 else S.mul(S.const(float(len(ks))), node))
            pairs.append((a, b))
        prog = S.compile_program(outs, p.dim)
        fids = sorted(set(j[0] for j in prog.jets))
        jets = [(fids.index(f), c, tuple(al) + (0,) * (3 - len(al))) for (f, c, al) in prog.jets]
        nder = max([max(al) for (_, _, al) in prog.jets]
                   + [max(tuple(k[0]) + tuple(k[1])) for k in keys])
        B = p.basis("iga", nder)
        nloc = list(B.nloc) + [1] * (3 - p.dim)
        nq = [int(B.c.nq[d]) for d in range(3)]
        kern = jit.get_op_kernel(prog, p.dim, nloc, nq, B.nder + 1, jets, len(fids), pairs,
                                 diag=True)
        d = dev.zeros(self.n)
        with dev.PROF.range("tigar_op diag=1 (generated Jacobi-diagonal kernel, per colour)", 0, 0):
            jit.launch_op(kern, B, [dev.ptr(self._funcs[f]) for f in fids], d, list(B.nloc))
        d[d == 0.0] = 1.0                      # as tg_win_diag_inv: a zero diagonal acts as 1
        return d.reciprocal_()                 # one-time element-wise reciprocal


class MatFreeOps(BlockOps):
    """``ops`` of multigpu.dist_cg for a FormOperator (one GPU)."""

    def __init__(self, op, jacobi=True):
        from . import dev
        from ._lib import lib, check
        self.dev, self.lib, self.check = dev, lib, check
        self.op = op
        self.n = op.n
        self.jacobi = jacobi

    def begin(self, b):
        dev, lib = self.dev, self.lib
        n = self.n
        self.b = b
        self.x = dev.zeros(n)
        self.r = dev.empty(n)
        self.q = dev.zeros(n)
        self.p = self.op.xvec                  # the operator reads its operand from p
        self.p.zero_()
        self.scratch = dev.empty(lib.tg_cg_scratch_len())
        self.s = dev.zeros(8)
        self.flip = 0
        self.dinv = self.op.jacobi_dinv() if self.jacobi else dev.zeros(n) + 1.0
        self.mask = self.op.spline._bc_mask() if self.op.applyBCs else None

    def matvec(self, x, y):
        assert x is self.p
        self.op.apply(y)
        if self.mask is not None:
            self.check(self.lib.tg_zero_entries(self.dev.ptr(y), self.dev.ptr(self.mask), self.n,
                                                self.dev.stream()))


def solve_matfree_fd(op, b, rtol=1e-12, atol=0.0, maxit=10000):
    """CG on the matrix-free operator, preconditioned by fast diagonalisation
    (tigar_b200/solvers.py): exact for an affine geometry, a few dozen operator applications
    otherwise -- what makes the 512^3 patch on one GPU practical (Jacobi-CG needs 175)."""
    from . import dev, solvers
    from ._lib import lib, check
    mask = op.spline._bc_mask() if op.applyBCs else None
    fd = solvers.FastDiag(op.patch, mask, op.diag, op.jacobi_dinv())
    n = op.n
    scratch = dev.empty(lib.tg_cg_scratch_len())
    s = dev.zeros(2)

    def spmv_dot(p, q):
        # y = P C P p + diag (I - P) p ; inside CG every iterate is zero on the constrained
        # DoFs, so this is C p with those entries zeroed
        with dev.PROF.range("tigar_op (matrix-free operator, one launch per colour)", 0, 0):
            if p.data_ptr() != op.xvec.data_ptr():
                op.xvec.copy_(p)
            if mask is not None:
                check(lib.tg_zero_entries(dev.ptr(op.xvec), dev.ptr(mask), n, dev.stream()))
            op.apply(q)
            if mask is not None:
                check(lib.tg_zero_entries(dev.ptr(q), dev.ptr(mask), n, dev.stream()))
        check(lib.tg_dot(dev.ptr(op.xvec), dev.ptr(q), n, dev.ptr(scratch), dev.ptr(s),
                         dev.stream()))
        return float(s[0].item())
    def precond(r, z):
        with dev.PROF.range("k_dgemm x6 + k_fd_scale (FD preconditioner)", 16 * 8 * n,
                            fd.flops_per_apply):
            return fd.apply(r, z)
    x, its, rel = solvers.pcg(spmv_dot, precond, b, None, rtol, atol, maxit, p_buf=op.xvec)
    return x, its, rel


def solve_matfree_cg(op, b, rtol=1e-12, atol=0.0, maxit=100000, check_every=5, jacobi=True):
    from .multigpu import dist_cg
    return dist_cg(MatFreeOps(op, jacobi), b, rtol, atol, maxit, check_every)
