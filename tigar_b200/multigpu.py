"""
Multi-GPU execution of the element-fused hot path: one process per GPU,
``torch.distributed`` (NCCL over NVLink/NVSwitch) for the exchange steps.

The reference distributes by MPI mesh partitioning plus an IGA DoF permutation
that aligns DoF ownership with FE-node ownership
(``generatePermutation``/``applyPermutation``, common.py:407-433, 1583-1669) and
leaves all communication to PETSc (MatPtAP, MatMult, KSP).  Here the
tensor-product structure gives that alignment for free:

* IGA DoF planes of the LAST parametric direction are split into contiguous
  slabs, one per rank; the rank owns those rows of ``C = M^T A M`` and of
  ``M^T b`` (contiguous in the reference's own DoF numbering, BSplines.py:354-358).
* Assembly: each rank runs the Gauss-point pass over the cell layers that touch
  its rows (its own layers plus p halo layers, recomputed instead of
  communicated) and the kernels skip rows it does not own -- no assembly
  communication, no atomics.
* Solve: row-distributed Jacobi-CG.  Per iteration one neighbour exchange of
  the p boundary planes of the search direction (send/recv) and two scalar
  all-reduces (p.Ap, then {r.z, r.r} fused in one 16-byte message).

``dist_cg`` is written against a small ``ops`` interface so that the exchange
logic is exercised on CPU with the gloo backend (tests/test_multigpu_cpu.py)
while the product always runs it with ``DeviceOps`` (the C-ABI kernels).
"""
import math
import os

import numpy as np


def plane_partition(ncp, cell_first, cell_last, win_last, rank, size):
    """Split ``ncp`` DoF planes into ``size`` contiguous slabs.

    cell_first/cell_last: first/last plane touched by each cell layer;
    win_last = (lo, hi): column window of every row plane in the last direction.
    Returns dict(k0, k1: owned planes; c0, c1: planes of the extended vector
    (owned + halo); cells: (lo, hi) cell layers to assemble; up, dn: number of
    halo planes received from rank+1 / rank-1)."""
    if size > ncp:
        raise ValueError("more ranks (%d) than DoF planes (%d)" % (size, ncp))
    k0 = (ncp * rank) // size
    k1 = (ncp * (rank + 1)) // size
    lo, hi = win_last
    c0 = int(np.min(lo[k0:k1]))
    c1 = int(np.max(hi[k0:k1])) + 1
    touch = np.flatnonzero((np.asarray(cell_last) >= k0) & (np.asarray(cell_first) < k1))
    cells = (int(touch[0]), int(touch[-1]) + 1)
    return dict(k0=k0, k1=k1, c0=c0, c1=c1, cells=cells, dn=k0 - c0, up=c1 - k1,
                bounds=[(ncp * r) // size for r in range(size + 1)])


# ----------------------------------------------------------------------------
def halo_plan(bounds, rank, c0, c1):
    """Which plane ranges to receive from / send to which rank so that the
    extended vector [c0, c1) is complete.  Returns (recvs, sends): lists of
    (peer, plane_lo, plane_hi) in GLOBAL plane indices.  Halos may span more
    than one neighbour when slabs are thinner than p."""
    size = len(bounds) - 1
    k0, k1 = bounds[rank], bounds[rank + 1]
    recvs, sends = [], []
    for r in range(size):
        if r == rank:
            continue
        a, b = bounds[r], bounds[r + 1]
        lo, hi = max(a, c0), min(b, c1)
        if lo < hi and (hi <= k0 or lo >= k1):
            recvs.append((r, lo, hi))
    return recvs, sends


def all_halo_plans(bounds, ext):
    """ext[r] = (c0, c1) of every rank -> per-rank (recvs, sends), consistent."""
    size = len(bounds) - 1
    plans = [([], []) for _ in range(size)]
    for r in range(size):
        recvs, _ = halo_plan(bounds, r, ext[r][0], ext[r][1])
        for (peer, lo, hi) in recvs:
            plans[r][0].append((peer, lo, hi))
            plans[peer][1].append((r, lo, hi))
    return plans


def dist_cg(ops, b, rtol, atol, maxit, check_every):
    """Row-distributed Jacobi-CG.  ``ops`` provides local kernels and the
    collectives; all vectors are rank-local.  Mirrors tg_cg_driver
    (tigar_b200/csrc/tg_solve.cu) step for step, with global reductions."""
    ops.begin(b)
    bb = ops.allreduce_host(ops.dot_bb())
    # x0 = 0: r = b, p = dinv r
    rz, rr = ops.init_residual()
    tol2 = max(rtol * rtol * bb, atol * atol)
    it = 0
    while rr > tol2 and it < maxit:
        nstep = min(check_every, maxit - it)
        for _ in range(nstep):
            ops.exchange_halo()
            ops.spmv_dot()             # q = C p_ext ; pAp (all-reduced, on device)
            ops.axpy_dot()             # x += a p ; r -= a q ; {rz', rr'} all-reduced
            ops.update_p()             # p = dinv r + (rz'/rz) p
            it += 1
        rz, rr = ops.read_rz_rr()
        if not (rr == rr):
            raise FloatingPointError("distributed CG produced NaN at iteration %d" % it)
    return ops.solution(), it, (math.sqrt(rr / bb) if bb > 0 else 0.0)


class DeviceOps(object):
    """CUDA implementation of the ``ops`` interface (C-ABI building blocks of
    include/tigar_b200.h + torch.distributed/NCCL)."""

    def __init__(self, patch, Cm):
        import torch
        import torch.distributed as dist
        from . import dev
        from ._lib import lib, check
        self.torch, self.dist, self.dev, self.lib, self.check = torch, dist, dev, lib, check
        self.patch, self.Cm = patch, Cm
        pp = patch.pp
        self.pp = pp
        self.plane = patch.plane
        self.n = patch.n_loc
        self.xoff = patch.xoff
        self.rank, self.size = patch.part
        # every rank's extended range: a pure function of the partition bounds and the global
        # window of the last direction, so each rank computes all of them (no all-gather)
        lo, hi = patch.window_global_C_last()
        b = pp["bounds"]
        ext = [(int(np.min(lo[b[r]:b[r + 1]])), int(np.max(hi[b[r]:b[r + 1]])) + 1)
               for r in range(self.size)]
        assert ext[self.rank] == (pp["c0"], pp["c1"])
        self.recvs, self.sends = all_halo_plans(pp["bounds"], ext)[self.rank]

    def begin(self, b):
        dev, lib, check = self.dev, self.lib, self.check
        n = self.n
        self.b = b
        self.x = dev.zeros(n)
        self.r = dev.empty(n)
        self.q = dev.empty(n)
        self.dinv = dev.empty(n)
        self.p_ext = dev.zeros(self.patch.n_ext)
        self.p = self.p_ext[self.xoff:self.xoff + n]
        self.scratch = dev.empty(lib.tg_cg_scratch_len())
        self.s = dev.zeros(8)          # 0 rz_a, 1 rr_a, 2 pAp, 3 rz_b, 4 rr_b, 5 bb
        self.flip = 0
        check(lib.tg_win_diag_inv(self.Cm.window.ref(), dev.ptr(self.Cm.vals),
                                  self.pp["k0"] - self.pp["c0"], dev.ptr(self.dinv),
                                  dev.stream()))

    def _sp(self, i):
        return self.dev.ptr(self.s) + 8 * i

    def dot_bb(self):
        dev, lib = self.dev, self.lib
        self.check(lib.tg_dot(dev.ptr(self.b), dev.ptr(self.b), self.n, dev.ptr(self.scratch),
                              self._sp(5), dev.stream()))
        return self.s[5:6]

    def allreduce_host(self, t):
        self.dist.all_reduce(t)
        return float(t[0].item())

    def init_residual(self):
        dev, lib = self.dev, self.lib
        self.q.zero_()
        self.check(lib.tg_cg_init(dev.ptr(self.b), dev.ptr(self.q), dev.ptr(self.dinv),
                                  dev.ptr(self.r), dev.ptr(self.p), self.n,
                                  dev.ptr(self.scratch), self._sp(0), dev.stream()))
        self.dist.all_reduce(self.s[0:2])
        h = self.s[0:2].tolist()
        return h[0], h[1]

    def exchange_halo(self):
        if not self.recvs and not self.sends:
            return
        dist, pl, c0 = self.dist, self.plane, self.pp["c0"]
        ops = []
        for (peer, lo, hi) in self.sends:
            ops.append(dist.P2POp(dist.isend, self.p_ext[(lo - c0) * pl:(hi - c0) * pl], peer))
        for (peer, lo, hi) in self.recvs:
            ops.append(dist.P2POp(dist.irecv, self.p_ext[(lo - c0) * pl:(hi - c0) * pl], peer))
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def spmv_dot(self):
        dev, lib = self.dev, self.lib
        self.check(lib.tg_win_spmv_dot(self.Cm.window.ref(), dev.ptr(self.Cm.vals),
                                       dev.ptr(self.p_ext), self.xoff, dev.ptr(self.q),
                                       dev.ptr(self.scratch), self._sp(2), dev.stream()))
        self.dist.all_reduce(self.s[2:3])

    def axpy_dot(self):
        dev, lib = self.dev, self.lib
        cur, nxt = (0, 3) if self.flip == 0 else (3, 0)
        self.check(lib.tg_cg_axpy_dot(dev.ptr(self.x), dev.ptr(self.r), dev.ptr(self.p),
                                      dev.ptr(self.q), dev.ptr(self.dinv), self.n,
                                      self._sp(cur), self._sp(2), dev.ptr(self.scratch),
                                      self._sp(nxt), dev.stream()))
        self.dist.all_reduce(self.s[nxt:nxt + 2])

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



# ----------------------------------------------------------------------------
class SlabTranspose(object):
    """Re-partition of a slab-distributed tensor-product vector (slabs of planes of the LAST
    direction, layout [i_L local][plane]) into chunks of the plane index holding whole fibres
    of the last direction (layout [i_L (all)][plane chunk] = column-major (chunk x n_L)), and
    back: one ``all_to_all_single`` each way.  Pure torch.distributed (NCCL on the device, gloo
    in tests/test_multigpu_cpu.py)."""

    def __init__(self, bounds, plane, rank, size):
        self.rank, self.size, self.plane = rank, size, plane
        self.bounds = list(bounds)
        self.nl = bounds[rank + 1] - bounds[rank]
        self.nL = bounds[-1]
        self.q = [(plane * r) // size for r in range(size + 1)]
        self.mq = self.q[rank + 1] - self.q[rank]
        self.nloc = self.nl * plane
        self.send_splits = [(self.q[s + 1] - self.q[s]) * self.nl for s in range(size)]
        self.recv_splits = [self.mq * (bounds[s + 1] - bounds[s]) for s in range(size)]

    def to_fibres(self, src, dst):
        import torch
        import torch.distributed as dist
        v = src[:self.nloc].view(self.nl, self.plane)
        pack = torch.cat([v[:, self.q[s]:self.q[s + 1]].reshape(-1) for s in range(self.size)])
        dist.all_to_all_single(dst[:self.mq * self.nL], pack, self.recv_splits, self.send_splits)

    def to_slabs(self, src, dst):
        import torch
        import torch.distributed as dist
        recv = torch.empty(self.nloc, dtype=src.dtype, device=src.device)
        dist.all_to_all_single(recv, src[:self.mq * self.nL].contiguous(), self.send_splits,
                               self.recv_splits)
        out = dst[:self.nloc].view(self.nl, self.plane)
        off = 0
        for s in range(self.size):
            w = self.q[s + 1] - self.q[s]
            out[:, self.q[s]:self.q[s + 1]].copy_(recv[off:off + w * self.nl].view(self.nl, w))
            off += w * self.nl


class FastDiagDist(object):
    """Fast-diagonalisation preconditioner (tigar_b200/solvers.py) on a slab-distributed
    vector.  The mode products of the first directions are local to a slab; the one along the
    LAST (partitioned) direction needs whole fibres, so the vector is re-partitioned by an
    all-to-all over NVLink into chunks of the (i0, i1) plane index, transformed, scaled in
    the eigenbasis, transformed back and returned by the reverse all-to-all: two exchanges
    of the local vector (17 MB per rank at 256^3 on 8 GPUs) per application -- the path's
    one genuine exchange step besides the halo of the matvec."""

    def __init__(self, patch, mask, diag, dinv_local):
        import torch
        import torch.distributed as dist
        from . import dev, solvers
        from ._lib import lib, check
        self.torch, self.dist, self.dev, self.lib, self.check = torch, dist, dev, lib, check
        self.patch = patch
        self.rank, self.size = patch.part
        pp = patch.pp
        self.dim = patch.dim
        L = self.dim - 1
        self.k0, self.k1 = pp["k0"], pp["k1"]
        self.nl = self.k1 - self.k0
        self.plane = patch.plane
        self.nloc = patch.n_loc
        self.nd = list(patch.ncp) + [1] * (3 - self.dim)
        self.nL = patch.ncp[L]
        bounds = pp["bounds"]
        self.bounds = bounds
        # plane chunks of the transposed partition
        self.tr = SlabTranspose(bounds, self.plane, self.rank, self.size)
        self.q, self.mq = self.tr.q, self.tr.mq
        self.gmask = mask
        self.lmask = None if mask is None else mask[self.k0 * self.plane:self.k1 * self.plane]
        self.cinv = 1.0 / float(diag) if diag else 1.0
        # 1-D eigen data: every rank computes the (tiny) decompositions itself
        helper = solvers.FastDiag.__new__(solvers.FastDiag)
        helper.patch, helper.dim = patch, self.dim
        free = solvers.FastDiag._free_planes(helper, mask)
        eig = [solvers._gen_eig(D, free[d]) for d, D in enumerate(patch.dirs)]
        c, sigma = self._fit(eig, free, dinv_local)
        self.weights, self.sigma = c, sigma
        self.S = None
        if os.environ.get("TIGAR_B200_FD_SCALE", "0") == "1":
            nloc3 = list(self.nd)
            nloc3[L] = self.nl
            self.S = solvers.diag_scale(eig, dinv_local, self.lmask, c, sigma, nloc3, self.dim,
                                        offset_last=self.k0)
        self.lam = [dev.from_np(c[d] * eig[d][0]) for d in range(self.dim)]
        self.U = [dev.from_np(np.ascontiguousarray(eig[d][1].T)) for d in range(self.dim)]
        self.t1 = dev.empty(max(self.nloc, self.mq * self.nL))
        self.t2 = dev.empty(max(self.nloc, self.mq * self.nL))

    def _fit(self, eig, free, dinv_local):
        """Least-squares fit of the direction weights to diag(C) (solvers.FastDiag._fit) with
        the right-hand side summed over the ranks."""
        torch, dist, dev, lib, check = self.torch, self.dist, self.dev, self.lib, self.check
        dim, L = self.dim, self.dim - 1
        md = [eig[d][2] for d in range(dim)]
        kd = [eig[d][3] for d in range(dim)]
        d_md = [dev.from_np(a) for a in md]
        d_kd = [dev.from_np(a) for a in kd]
        scratch = dev.empty(256)
        out4 = dev.zeros(4)

        def P(Lst, d):
            if d >= dim:
                return None
            return dev.ptr(Lst[d]) + (8 * self.k0 if d == L else 0)
        nloc3 = list(self.nd)
        nloc3[L] = self.nl
        dC = dinv_local.reciprocal()
        if os.environ.get("TIGAR_B200_FD_FIT", "rel") != "abs":
            from .solvers import choose_fd_weights
            scr = dev.empty(15 * 64)
            out15 = dev.zeros(15)
            check(lib.tg_fd_fit_rel(dev.ptr(dC), dev.ptr(self.lmask) if self.lmask is not None
                                    else None, P(d_kd, 0), P(d_kd, 1), P(d_kd, 2), P(d_md, 0),
                                    P(d_md, 1), P(d_md, 2), nloc3[0], nloc3[1], nloc3[2],
                                    dev.ptr(scr), dev.ptr(out15), dev.stream()))
            dist.all_reduce(out15)
            return choose_fd_weights(dev.to_np(out15), dim)
        check(lib.tg_fd_fit(dev.ptr(dC), dev.ptr(self.lmask) if self.lmask is not None else None,
                            P(d_kd, 0), P(d_kd, 1), P(d_kd, 2), P(d_md, 0), P(d_md, 1),
                            P(d_md, 2), nloc3[0], nloc3[1], nloc3[2], dev.ptr(scratch),
                            dev.ptr(out4), dev.stream()))
        dist.all_reduce(out4)
        rhs = dev.to_np(out4)
        nb = dim + 1

        def f(a, dd):
            v = kd[dd] if a == dd else md[dd]
            return np.where(free[dd], v, 0.0)
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
            return [1.0] * dim, 0.0
        return [float(x) for x in c], max(sigma, 0.0) if sigma > -1e-10 * max(c) else 0.0

    def _gemm(self, ta, tb, M, N, K, A, lda, sA, B, ldb, sB, Cc, ldc, sC, batch):
        self.check(self.lib.tg_dgemm_batched(ta, tb, M, N, K, 1.0, A, lda, sA, B, ldb, sB, 0.0,
                                             Cc, ldc, sC, batch, self.dev.stream()))

    def _to_fibres(self, src, dst):
        self.tr.to_fibres(src, dst)

    def _to_slabs(self, src, dst):
        self.tr.to_slabs(src, dst)

    def apply(self, r, z):
        dev, lib, check = self.dev, self.lib, self.check
        dim, L = self.dim, self.dim - 1
        n0, n1, _ = self.nd
        nl, nL, mq = self.nl, self.nL, self.mq
        st = dev.stream()
        U = [dev.ptr(u) for u in self.U]
        a, b = self.t1, self.t2
        check(lib.tg_masked_copy(dev.ptr(a), dev.ptr(r), dev.ptr(self.lmask) if self.lmask
                                 is not None else None, self.nloc, st))
        if self.S is not None:
            check(lib.tg_vmul(dev.ptr(a), dev.ptr(a), dev.ptr(self.S), self.nloc, st))
        # local mode products (directions before the partitioned one)
        if dim == 3:
            self._gemm(1, 0, n0, n1 * nl, n0, U[0], n0, 0, dev.ptr(a), n0, 0, dev.ptr(b), n0, 0, 1)
            a, b = b, a
            self._gemm(0, 0, n0, n1, n1, dev.ptr(a), n0, n0 * n1, U[1], n1, 0, dev.ptr(b), n0,
                       n0 * n1, nl)
            a, b = b, a
        else:
            self._gemm(1, 0, n0, nl, n0, U[0], n0, 0, dev.ptr(a), n0, 0, dev.ptr(b), n0, 0, 1)
            a, b = b, a
        # whole fibres of the last direction on every rank
        self._to_fibres(a, b)
        a, b = b, a
        self._gemm(0, 0, mq, nL, nL, dev.ptr(a), mq, 0, U[L], nL, 0, dev.ptr(b), mq, 0, 1)
        a, b = b, a
        lam = [dev.ptr(l) for l in self.lam] + [None] * (3 - dim)
        if dim == 3:
            check(lib.tg_fd_scale(dev.ptr(a), lam[0], lam[1], lam[2], n0, n1, nL,
                                  self.q[self.rank], mq, self.sigma, 1, st))
        else:
            # 2-D: the "plane" is the first direction, the fibre direction is the second
            check(lib.tg_fd_scale(dev.ptr(a), lam[0], None, lam[1], n0, 1, nL,
                                  self.q[self.rank], mq, self.sigma, 1, st))
        self._gemm(0, 1, mq, nL, nL, dev.ptr(a), mq, 0, U[L], nL, 0, dev.ptr(b), mq, 0, 1)
        a, b = b, a
        self._to_slabs(a, b)
        a, b = b, a
        if dim == 3:
            self._gemm(0, 1, n0, n1, n1, dev.ptr(a), n0, n0 * n1, U[1], n1, 0, dev.ptr(b), n0,
                       n0 * n1, nl)
            a, b = b, a
            self._gemm(0, 0, n0, n1 * nl, n0, U[0], n0, 0, dev.ptr(a), n0, 0, dev.ptr(z), n0, 0, 1)
        else:
            self._gemm(0, 0, n0, nl, n0, U[0], n0, 0, dev.ptr(a), n0, 0, dev.ptr(z), n0, 0, 1)
        if self.S is not None:
            check(lib.tg_vmul(dev.ptr(z), dev.ptr(z), dev.ptr(self.S), self.nloc, st))
        if self.lmask is not None:
            check(lib.tg_masked_fix(dev.ptr(z), dev.ptr(r), dev.ptr(self.lmask), self.cinv,
                                    self.nloc, st))
        return z


def solve_fd_pcg_dist(patch, Cm, b, mask, diag, rtol, atol, maxit):
    """CG preconditioned by fast diagonalisation on the slab-distributed system: the halo
    exchange + local SpMV of ``DeviceOps``, the preconditioner above, scalars all-reduced."""
    import torch
    import torch.distributed as dist
    from . import dev, solvers
    from ._lib import lib, check
    ops = DeviceOps(patch, Cm)
    ops.begin(b)                    # allocates p_ext / dinv (Jacobi diagonal: used for the fit)
    fd = FastDiagDist(patch, mask, diag, ops.dinv)
    s = dev.zeros(2)
    red = dev.zeros(1)

    def allsum(v):
        red.fill_(v)
        dist.all_reduce(red)
        return float(red.item())

    def spmv_dot(p, q):
        assert p.data_ptr() == ops.p.data_ptr()
        ops.exchange_halo()
        check(lib.tg_win_spmv_dot(Cm.window.ref(), dev.ptr(Cm.vals), dev.ptr(ops.p_ext), ops.xoff,
                                  dev.ptr(q), dev.ptr(ops.scratch), dev.ptr(s), dev.stream()))
        dist.all_reduce(s[0:1])              # reduced on the device: one host read, not two
        return float(s[0].item())
    x, its, rel = solvers.pcg(spmv_dot, fd.apply, b, None, rtol, atol, maxit, reduce=allsum,
                              p_buf=ops.p, reduce_dev=dist.all_reduce)
    return x, its, rel


def gather_planes(local, patch):
    """All-gather a slab-distributed IGA vector into the full vector (every
    rank gets it): the FE/IGA functions the forms evaluate are replicated."""
    import torch
    import torch.distributed as dist
    from . import dev
    size = patch.part[1]
    bounds = patch.pp["bounds"]
    pl = patch.plane
    mx = max(bounds[r + 1] - bounds[r] for r in range(size)) * pl
    pad = dev.zeros(mx)
    pad[:local.numel()].copy_(local)
    out = dev.empty(mx * size)
    dist.all_gather_into_tensor(out, pad)
    full = dev.empty(patch.n_iga)
    for r in range(size):
        n = (bounds[r + 1] - bounds[r]) * pl
        full[bounds[r] * pl:bounds[r] * pl + n].copy_(out[r * mx:r * mx + n])
    return full
