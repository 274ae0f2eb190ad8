"""CPU (gloo, world_size 2 and 3) tests of the multi-GPU host logic: slab
partition, halo plan, and the row-distributed CG driver ``dist_cg`` run with a
numpy ``ops`` test double in place of the CUDA kernels, against the oracle."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import bsplines as OB
from oracle import pipeline as OP
from tigar_b200.multigpu import plane_partition, all_halo_plans, dist_cg


def test_plane_partition_covers_everything():
    ncp, p = 19, 3
    first = np.arange(16)
    lo = np.maximum(np.arange(ncp) - p, 0)
    hi = np.minimum(np.arange(ncp) + p, ncp - 1)
    for size in (1, 2, 3, 5, 8):
        parts = [plane_partition(ncp, first, first + p, (lo, hi), r, size) for r in range(size)]
        assert parts[0]["k0"] == 0 and parts[-1]["k1"] == ncp
        for a, b in zip(parts[:-1], parts[1:]):
            assert a["k1"] == b["k0"]
        for pp in parts:
            assert pp["c0"] == max(pp["k0"] - p, 0) and pp["c1"] == min(pp["k1"] + p, ncp)
            e0, e1 = pp["cells"]
            # every cell that touches an owned plane is assembled, and no other
            touch = [e for e in range(16) if e + p >= pp["k0"] and e < pp["k1"]]
            assert list(range(e0, e1)) == touch
        ext = [(pp["c0"], pp["c1"]) for pp in parts]
        plans = all_halo_plans(parts[0]["bounds"], ext)
        for r, (recvs, sends) in enumerate(plans):
            got = set()
            for (peer, a, b) in recvs:
                assert peer != r and parts[peer]["k0"] <= a < b <= parts[peer]["k1"]
                got |= set(range(a, b))
            need = set(range(ext[r][0], ext[r][1])) - set(range(parts[r]["k0"], parts[r]["k1"]))
            assert got == need
            for (peer, a, b) in sends:
                assert (r, a, b) in plans[peer][0]
    with pytest.raises(ValueError):
        plane_partition(3, first[:1], first[:1] + 2, (lo[:3], hi[:3]), 0, 4)


class NumpyOps(object):
    """Test double for multigpu.DeviceOps: same interface, numpy + gloo."""

    def __init__(self, Cfull, pp, plane, dist, torch):
        self.pp, self.plane, self.dist, self.torch = pp, plane, dist, torch
        r0, r1 = pp["k0"] * plane, pp["k1"] * plane
        c0, c1 = pp["c0"] * plane, pp["c1"] * plane
        self.C = Cfull[r0:r1, c0:c1].tocsr()
        assert Cfull[r0:r1].nnz == self.C.nnz           # the extended range holds every column
        self.dinv = 1.0 / Cfull.diagonal()[r0:r1]
        self.n, self.xoff = r1 - r0, r0 - c0
        rank = dist.get_rank()
        ext = [None] * dist.get_world_size()
        dist.all_gather_object(ext, (pp["c0"], pp["c1"]))
        self.recvs, self.sends = all_halo_plans(pp["bounds"], ext)[rank]

    def _ar(self, vals):
        t = self.torch.tensor(vals, dtype=self.torch.float64)
        self.dist.all_reduce(t)
        return t.tolist()

    def begin(self, b):
        self.b = b
        self.x = np.zeros(self.n)
        self.p_ext = np.zeros(self.C.shape[1])

    def dot_bb(self):
        return self.torch.tensor([float(self.b @ self.b)], dtype=self.torch.float64)

    def allreduce_host(self, t):
        self.dist.all_reduce(t)
        return float(t[0])

    def init_residual(self):
        self.r = self.b.copy()
        z = self.dinv * self.r
        self.p_ext[self.xoff:self.xoff + self.n] = z
        self.rz, self.rr = self._ar([float(self.r @ z), float(self.r @ self.r)])
        return self.rz, self.rr

    def exchange_halo(self):
        dist, torch, pl, c0 = self.dist, self.torch, self.plane, self.pp["c0"]
        reqs, bufs = [], []
        for (peer, lo, hi) in self.sends:
            t = torch.from_numpy(self.p_ext[(lo - c0) * pl:(hi - c0) * pl].copy())
            reqs.append(dist.isend(t, peer))
        for (peer, lo, hi) in self.recvs:
            t = torch.zeros((hi - lo) * pl, dtype=torch.float64)
            bufs.append((t, lo, hi))
            reqs.append(dist.irecv(t, peer))
        for q in reqs:
            q.wait()
        for (t, lo, hi) in bufs:
            self.p_ext[(lo - c0) * pl:(hi - c0) * pl] = t.numpy()

    def spmv_dot(self):
        self.q = self.C @ self.p_ext
        p = self.p_ext[self.xoff:self.xoff + self.n]
        self.pAp = self._ar([float(p @ self.q)])[0]

    def axpy_dot(self):
        a = self.rz / self.pAp
        p = self.p_ext[self.xoff:self.xoff + self.n]
        self.x += a * p
        self.r -= a * self.q
        self.rz_new, self.rr = self._ar([float(self.r @ (self.dinv * self.r)),
                                         float(self.r @ self.r)])

    def update_p(self):
        p = self.p_ext[self.xoff:self.xoff + self.n]
        p[:] = self.dinv * self.r + (self.rz_new / self.rz) * p
        self.rz = self.rz_new

    def read_rz_rr(self):
        return self.rz, self.rr

    def solution(self):
        return self.x


def _worker(rank, size, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        deg, nel = [2, 2, 2], [4, 3, 5]
        kv = [OB.uniform_knots(p, 0.0, 1.0, n) for p, n in zip(deg, nel)]
        pr = OP.Problem(deg, kv)
        Uo = pr.run(lambda X: 1.0 + X[..., 0] * X[..., 2])
        s = pr.ts.splines[-1]
        plane = pr.ts.ncp // s.ncp
        first = s.element_spans() - s.p
        # last-direction window of C: supports overlap
        lo = np.maximum(np.arange(s.ncp) - s.p, 0)
        hi = np.minimum(np.arange(s.ncp) + s.p, s.ncp - 1)
        pp = plane_partition(s.ncp, first, first + s.p, (lo, hi), rank, size)
        ops = NumpyOps(pr.C.tocsr(), pp, plane, dist, torch)
        b = pr.b[pp["k0"] * plane:pp["k1"] * plane].copy()
        x, its, rel = dist_cg(ops, b, 1e-13, 0.0, 5000, 7)
        ref = Uo[pp["k0"] * plane:pp["k1"] * plane]
        err = np.linalg.norm(x - ref) / np.linalg.norm(Uo)
        q.put((rank, err, its))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("size", [2, 3])
def test_distributed_cg_driver_gloo(size):
    import torch.multiprocessing as mp
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, size, port, q)) for r in range(size)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(size)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    its = set(r[2] for r in res)
    assert len(its) == 1                      # all ranks took the same number of iterations
    for rank, err, _ in res:
        assert err < 1e-10, (rank, err)


def _fd_worker(rank, size, port, q):
    """Row-distributed fast-diagonalisation apply on CPU tensors: the slab <-> fibre
    re-partition of multigpu.SlabTranspose (gloo all-to-all) around numpy mode products, against
    the single-process Kronecker formula."""
    import torch
    import torch.distributed as dist
    from tigar_b200.multigpu import SlabTranspose
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        n0, n1, n2 = 5, 4, 7                       # 7 planes over 2 or 3 ranks: uneven slabs
        rng = np.random.RandomState(11)
        U = [np.linalg.qr(rng.rand(n, n))[0] for n in (n0, n1, n2)]
        lam = [1.0 + rng.rand(n) for n in (n0, n1, n2)]
        r_full = rng.rand(n2, n1, n0)              # [i2][i1][i0], i0 fastest
        # reference: z = (U2 x U1 x U0) D^-1 (U2 x U1 x U0)^T r
        t = np.einsum("kji,ia,jb,kc->cba", r_full, U[0], U[1], U[2])
        t = t / (lam[2][:, None, None] + lam[1][None, :, None] + lam[0][None, None, :])
        z_ref = np.einsum("cba,ia,jb,kc->kji", t, U[0], U[1], U[2])
        bounds = [(n2 * r) // size for r in range(size + 1)]
        k0, k1 = bounds[rank], bounds[rank + 1]
        plane = n0 * n1
        tr = SlabTranspose(bounds, plane, rank, size)
        assert tr.nl == k1 - k0 and sum(tr.send_splits) == tr.nloc
        loc = r_full[k0:k1]                         # this rank's slab
        # local mode products (directions 0 and 1)
        a = np.einsum("kji,ia,jb->kba", loc, U[0], U[1])
        src = torch.from_numpy(np.ascontiguousarray(a).ravel())
        fib = torch.zeros(max(tr.nloc, tr.mq * n2), dtype=torch.float64)
        tr.to_fibres(src, fib)
        X = fib[:tr.mq * n2].numpy().reshape(n2, tr.mq)          # [i2][plane chunk]
        # every rank holds whole fibres of the last direction for its plane chunk
        full = np.einsum("kji,ia,jb->kba", r_full, U[0], U[1]).reshape(n2, plane)
        assert np.array_equal(X, full[:, tr.q[rank]:tr.q[rank + 1]])
        Y = U[2].T @ X                                            # forward along direction 2
        pl = np.arange(tr.q[rank], tr.q[rank + 1])
        D = lam[2][:, None] + lam[1][pl // n0][None, :] + lam[0][pl % n0][None, :]
        Y = U[2] @ (Y / D)
        back = torch.zeros(max(tr.nloc, tr.mq * n2), dtype=torch.float64)
        tr.to_slabs(torch.from_numpy(np.ascontiguousarray(Y).ravel()), back)
        b_ = back[:tr.nloc].numpy().reshape(k1 - k0, n1, n0)
        z = np.einsum("kba,ia,jb->kji", b_, U[0], U[1])
        q.put((rank, float(np.abs(z - z_ref[k0:k1]).max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("size", [2, 3])
def test_distributed_fast_diagonalisation_gloo(size):
    import torch.multiprocessing as mp
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_fd_worker, args=(r, size, port, q)) for r in range(size)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(size)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err in res:
        assert err < 1e-12, (rank, err)
