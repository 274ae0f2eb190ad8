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
import ctypes as C
import json
import math
import os
import time

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
        # every rank's extended range (tiny all-gather of two ints)
        mine = torch.tensor([pp["c0"], pp["c1"]], dtype=torch.int64, device=dev.device())
        allr = [torch.zeros_like(mine) for _ in range(self.size)]
        dist.all_gather(allr, mine)
        ext = [(int(t[0]), int(t[1])) for t in allr]
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


# ----------------------------------------------------------------------------
def bench(args, METRIC, UNIT, CG_RTOL):
    """bench.py leg for N > 1 (launched by torchrun, one rank per GPU): the same
    256^3 workload, strong scaling; max-over-ranks device time."""
    import torch
    import torch.distributed as dist
    import bench as B
    from ._lib import lib
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nel = args.nel
    kv, cm, pinned = B.build_inputs(nel)
    dev_cols = [pinned[:, i].contiguous().cuda() for i in range(4)]

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        B.one_step(kv, cm, dev_cols, "fused", CG_RTOL, False)
    barrier()
    sampler = B.ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.tg_prof_enable(0)
    l0 = lib.tg_launch_count()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    stages = []
    for _ in range(args.steps):
        n_dofs, iters, ev, _, MTAM = B.one_step(kv, cm, dev_cols, "fused", CG_RTOL, False)
        stages.append(ev)
    t1.record()
    barrier()
    ms = torch.tensor([t0.elapsed_time(t1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = lib.tg_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    W = MTAM.window
    stage_ms = [0.0, 0.0, 0.0]
    for ev in stages:
        for i in range(3):
            stage_ms[i] += ev[i].elapsed_time(ev[i + 1]) / max(args.steps, 1)
    st = torch.tensor(stage_ms, device="cuda")
    dist.all_reduce(st, op=dist.ReduceOp.MAX)
    local_nnz, local_rows = W.nnz, W.nrows
    del MTAM, stages
    # live timing of the local SpMV (device events around the kernel only)
    # end to end from host buffers: rank 0 holds the host control net and
    # broadcasts it; every rank returns its slab, gathered to the host of rank 0
    B.one_step(kv, cm, pinned, "fused", CG_RTOL, True)
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        _, _, _, res, MT = B.one_step(kv, cm, pinned, "fused", CG_RTOL, True)
        del MT
    e1.record()
    barrier()
    e2e_ms = torch.tensor([max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - wall0))],
                          device="cuda")
    dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())
    if rank == 0:
        peak, which = B.measured_peaks()
        out = {
            "metric": METRIC, "value": n_dofs * args.steps / (ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "3D cubic B-spline Poisson %d^3 cells, %d GPUs" % (nel, world),
                       "degree": B.P, "cells": nel ** 3, "iga_dofs": n_dofs, "path": "fused",
                       "partition": "slabs of IGA planes (last direction), p halo cell layers "
                                    "recomputed, row-distributed Jacobi-CG over NCCL",
                       "quad_degree": 2 * B.P, "cg_rtol": CG_RTOL, "cg_iterations": iters,
                       "l2": "inputs larger than L2 (local matrix %.1f GB per rank)"
                             % (8e-9 * local_nnz)},
            "stage_ms": {"extract": float(st[0]), "assemble_ptap_bcs": float(st[1]),
                         "solve": float(st[2])},
            "clocks": clocks,
            "e2e": {"value": n_dofs * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(pinned.numel() * 8 * world),
                    "d2h_bytes_per_step": int(n_dofs * 8)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_win_spmv<true> (rank-local rows)",
                         "achieved": None, "peak": peak, "peak_source": which, "unit": "GB/s",
                         "frac": None, "traffic": None,
                         "note": "per-kernel roofline is reported by the N=1 run"}}
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
