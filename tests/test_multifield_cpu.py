"""CPU tests of the equal-order multi-field host logic (SURVEY 8f n1;
tigar_b200/multifield.py, the vector-valued arguments / Functions of tigar_b200/api.py)
against the oracle's block assembly (oracle.assembly.assemble_elasticity).

The product's form language turns a vector-valued bilinear form into per-(test field,
trial field) term lists with symbolic coefficients; the device kernels then integrate
each list.  Here the same term lists are integrated on the host (a vectorised
interpreter of the coefficient programs + the oracle's basis tables), so the splitting,
the symbolic geometry and the Gateaux derivative are checked without a GPU.
"""
import math

import numpy as np
import pytest
import scipy.sparse as sp

from tigar_b200 import symbolic as S
from tigar_b200 import ufl_lite as U
from tigar_b200 import multifield as MF
from oracle import assembly as OA
from oracle import bsplines as OB
from oracle import pipeline as OP


# ------------------------------------------------------------ host interpreter
def vec_run_program(prog, xi, wq, jetvals):
    """Coefficient program on arrays of points: xi [dim, n], wq [n], jetvals {jet: [n]}."""
    n = wq.shape[0]
    R = np.zeros((prog.nreg, n))
    R[:prog.dim] = xi
    R[prog.dim] = wq
    for k, j in enumerate(prog.jets):
        R[prog.dim + 1 + k] = jetvals[j]
    f1 = dict(neg=lambda a: -a, sin=np.sin, cos=np.cos, exp=np.exp, log=np.log, sqrt=np.sqrt,
              abs=np.abs, tan=np.tan, tanh=np.tanh, sinh=np.sinh, cosh=np.cosh,
              atan=np.arctan, mov=lambda a: a)
    f2 = dict(add=np.add, sub=np.subtract, mul=np.multiply, div=np.divide, pow=np.power,
              max=np.maximum, min=np.minimum, gt=lambda a, b: (a > b).astype(float),
              selz=lambda a, b: np.where(np.asarray(a) != 0, b, 0.0))
    names = {v: k for k, v in S.OPCODES.items()}
    for op, dst, a, b in prog.prog:
        nm = names[op]
        if nm == "const":
            R[dst] = prog.consts[a]
        elif nm in f1:
            R[dst] = f1[nm](R[a])
        else:
            R[dst] = f2[nm](R[a], R[b])
    return [R[r].copy() for r in prog.outregs]


class _Patch(object):
    part = None


def symbolic_spline(dim, nsd, nfields, n_iga):
    """An ExtractedSpline with its symbolic geometry (common.py:896-945) but no device
    state: enough for the form language."""
    from tigar_b200 import api as A
    sp_ = object.__new__(A.ExtractedSpline)
    sp_._patch = _Patch()
    sp_._patch.dim, sp_._patch.n_iga = dim, n_iga
    sp_.nsd, sp_.nFields, sp_.generator = nsd, nfields, None
    sp_.V = A.FunctionSpace(sp_, nfields)
    sp_.V_control = A.FunctionSpace(sp_, 1, control=True)
    sp_.cpFuncs = [A.Function(sp_.V_control) for _ in range(nsd + 1)]
    sp_.genericSetup()
    return sp_


class HostIntegrator(object):
    """Integrates term lists {(alphaTest, alphaTrial): node} / {alphaTest: node} over the
    patch in the B-spline basis with the oracle's tables."""

    def __init__(self, ts, P, nq, funcs, order=2):
        self.dim = ts.nvar
        self.tabs = [OA.tab_iga(s, nq, order) for s in ts.splines]
        ncell = int(np.prod([tb.T.shape[0] for tb in self.tabs]))
        self.blk = OA.CellBlock(self.tabs, np.arange(ncell), order)
        self.n = int(np.prod([tb.n for tb in self.tabs]))
        self.funcs = funcs                         # {fid: coefficient vector [n]}
        self.order = order

    def _eval(self, nodes):
        blk, dim = self.blk, self.dim
        prog = S.compile_program(nodes, dim)
        nc, nqp = blk.wq.shape
        jv = {}
        for (fid, comp, al) in prog.jets:
            assert comp == 0
            cc = self.funcs[fid][blk.gidx]                                   # [c, a]
            jv[(fid, comp, al)] = np.einsum("cqa,ca->cq", blk.jets[tuple(al[:dim])], cc).ravel()
        xi = np.stack([blk.xi[..., d].ravel() for d in range(dim)])
        out = vec_run_program(prog, xi, blk.wq.ravel(), jv)
        return [o.reshape(nc, nqp) for o in out]

    def matrix(self, terms):
        keys = sorted(terms)
        vals = self._eval([terms[k] for k in keys])
        blk, dim = self.blk, self.dim
        Ke = 0.0
        for (aT, aU), c in zip(keys, vals):
            Ke = Ke + np.einsum("cq,cqa,cqb->cab", c, blk.jets[tuple(aT[:dim])],
                                blk.jets[tuple(aU[:dim])], optimize=True)
        nen = blk.gidx.shape[1]
        r = np.repeat(blk.gidx, nen, axis=1).ravel()
        c_ = np.tile(blk.gidx, (1, nen)).ravel()
        return sp.coo_matrix((Ke.ravel(), (r, c_)), shape=(self.n, self.n)).tocsr()

    def vector(self, terms):
        keys = sorted(terms)
        vals = self._eval([terms[k] for k in keys])
        b = np.zeros(self.n)
        for aT, c in zip(keys, vals):
            fe = np.einsum("cq,cqa->ca", c, self.blk.jets[tuple(aT[:self.dim])])
            np.add.at(b, self.blk.gidx.ravel(), fe.ravel())
        return b


def curved_patch(p, nel, amp=0.08):
    """2-D explicit B-spline patch with a smoothly perturbed control net and
    non-constant weights (a genuinely rational, non-affine map)."""
    kv = [OB.uniform_knots(p, 0.0, 1.0, n) for n in nel]
    ts = OB.TensorSpline([p, p], kv)
    P = OB.explicit_control_net(ts, 0).copy()
    x, y = P[:, 0].copy(), P[:, 1].copy()
    P[:, 0] = x + amp * np.sin(math.pi * x) * np.sin(math.pi * y)
    P[:, 1] = y + amp * x * (1 - x) * np.cos(1.3 * y)
    w = 1.0 + 0.2 * x * y
    P[:, :2] *= w[:, None]
    P[:, 2] = w
    return ts, kv, P


MU, LAM = 0.7, 1.9


def body_force(X):
    return np.stack([np.sin(2.0 * X[..., 0]) * X[..., 1], 0.5 + X[..., 0] * X[..., 1] ** 2], -1)


def elasticity_forms(spline, u, v):
    from tigar_b200 import api as A   # noqa: F401

    def eps(w):
        g = spline.grad(w)
        return 0.5 * (g + g.T)
    x = spline.spatialCoordinates()
    f = U.as_vector([U.sin(2.0 * x[0]) * x[1], 0.5 + x[0] * x[1] ** 2])
    a = (2.0 * MU * U.inner(eps(u), eps(v)) + LAM * spline.div(u) * spline.div(v)) * spline.dx
    L = U.inner(f, v) * spline.dx
    return a, L


def test_arguments_and_term_splitting():
    from tigar_b200 import api as A
    spl = symbolic_spline(2, 2, 2, 10)
    u, v = A.TrialFunction(spl.V), A.TestFunction(spl.V)
    assert u.ufl_shape == (2,) and v.ufl_shape == (2,)
    assert list(u[1].a[()].terms) == [(None, (0, 0, 0, 1))]
    assert list(v[0].a[()].terms) == [((0, 0, 0, 0), None)]
    # d/dxi keeps the field, raises the multi-index
    assert list(u[1].dx(0).a[()].terms) == [(None, (1, 0, 0, 1))]
    m = (u[0].dx(1) * v[1]).a[()]
    blocks = MF.split_matrix_terms(m.terms, 2)
    assert list(blocks) == [(1, 0)] and list(blocks[(1, 0)]) == [((0, 0, 0), (0, 1, 0))]
    with pytest.raises(IndexError):
        MF.split_matrix_terms(m.terms, 1)
    with pytest.raises(ValueError):
        MF.split_vector_terms(m.terms, 2)
    z = MF.split_zero_dofs([0, 3, 10, 19, 12], 2, 10)
    assert [list(a) for a in z] == [[0, 3], [0, 9, 2]]
    with pytest.raises(IndexError):
        MF.split_zero_dofs([20], 2, 10)
    # scalar spaces keep the 3-index keys of the single-field hot path
    s1 = symbolic_spline(2, 2, 1, 10)
    assert list(A.TrialFunction(s1.V).a[()].terms) == [(None, (0, 0, 0))]
    assert MF.part_field((0, 1, 0)) == 0 and MF.part_field(None) is None


@pytest.mark.parametrize("p,nel", [(2, (5, 4)), (3, (4, 4))])
def test_block_terms_integrate_to_the_oracle_elasticity_system(p, nel):
    """Every (test field, trial field) term list of the vector-valued form, integrated on
    the host, equals the oracle's independently written block (<= 1e-12), and so does the
    load vector; the Gateaux derivative of the residual gives the same blocks."""
    from tigar_b200 import api as A
    ts, kv, P = curved_patch(p, nel)
    n = ts.ncp
    spl = symbolic_spline(2, 2, 2, n)
    funcs = {f.fid: P[:, i].copy() for i, f in enumerate(spl.cpFuncs)}
    H = HostIntegrator(ts, P, p + 1, funcs, order=1)
    u, v = A.TrialFunction(spl.V), A.TestFunction(spl.V)
    a, L = elasticity_forms(spl, u, v)
    prob = OP.ElasticityProblem([p, p], kv, P, MU, LAM, [])
    Ao, bo = prob.direct_iga(body_force)
    scale = abs(Ao).max()

    mterms = spl._weighted(a.scalar())
    blocks = MF.split_matrix_terms(mterms, 2)
    assert sorted(blocks) == [(0, 0), (0, 1), (1, 0), (1, 1)]
    K = {}
    for (f, g), bt in blocks.items():
        K[(f, g)] = H.matrix(bt)
        ref = Ao[f * n:(f + 1) * n, g * n:(g + 1) * n]
        assert abs(K[(f, g)] - ref).max() < 1e-12 * scale, (f, g)
    vterms = MF.split_vector_terms(spl._weighted(L.scalar()), 2)
    for f, vt in vterms.items():
        bf = H.vector(vt)
        assert np.abs(bf - bo[f * n:(f + 1) * n]).max() < 1e-13 * max(1.0, np.abs(bo).max())

    # residual at a given displacement state; its Gateaux derivative is the bilinear form
    uh = A.Function(spl.V)
    assert isinstance(uh, A.VectorFunction) and uh.ufl_shape == (2,)
    R = elasticity_forms(spl, uh, v)[0]
    assert R.arity() == 1
    J = A.derivative(R, uh)
    jb = MF.split_matrix_terms(spl._weighted(J.scalar()), 2)
    rng = np.random.RandomState(1)
    Uh = rng.rand(2 * n)
    H.funcs.update({c.fid: Uh[i * n:(i + 1) * n] for i, c in enumerate(uh.comps)})
    for fg, bt in jb.items():
        assert abs(H.matrix(bt) - K[fg]).max() < 1e-12 * scale
    # ... and the residual vector is A_o U
    rv = MF.split_vector_terms(spl._weighted(R.scalar()), 2)
    Rh = np.concatenate([H.vector(rv[f]) for f in range(2)])
    assert np.abs(Rh - Ao @ Uh).max() < 1e-11 * scale


def test_oracle_multifield_fe_path_equals_direct_iga():
    """SURVEY 8c KAT 3 for the multi-field system: (I (x) M)^T A_FE (I (x) M) and
    (I (x) M)^T b_FE equal the directly assembled IGA Galerkin system."""
    ts, kv, P = curved_patch(2, (4, 3))
    prob = OP.ElasticityProblem([2, 2], kv, P, MU, LAM, [])
    C1, b1 = prob.fe_path(body_force)
    C2, b2 = prob.direct_iga(body_force)
    assert abs(C1 - C2).max() < 1e-12 * abs(C2).max()
    assert np.abs(b1 - b2).max() < 1e-13 * max(1.0, np.abs(b2).max())
    assert abs(C2 - C2.T).max() < 1e-13 * abs(C2).max()
    # rigid translations are in the kernel of the un-constrained operator (the B-spline
    # fields are not rationalised here, and sum_a N_a = 1 whatever the geometry map is)
    n = ts.ncp
    for f in range(2):
        t = np.zeros(2 * n)
        t[f * n:(f + 1) * n] = 1.0
        assert np.abs(C2 @ t).max() < 1e-11 * abs(C2).max()


class NumpyBlockOps(object):
    """numpy twin of multifield.BlockOps (same call sequence, same scalar slots)."""

    def __init__(self, blocks, nf, nb):
        self.blocks, self.nf, self.nb, self.n = blocks, nf, nb, nf * nb

    def begin(self, b):
        self.b = b
        self.x, self.r, self.q, self.p = (np.zeros(self.n) for _ in range(4))
        self.s = np.zeros(8)
        self.flip = 0
        self.dinv = np.concatenate([1.0 / self.blocks[(f, f)].diagonal() for f in range(self.nf)])

    def dot_bb(self):
        self.s[5] = self.b @ self.b
        return self.s[5:6]

    def allreduce_host(self, t):
        return float(t[0])

    def init_residual(self):
        self.r[:] = self.b - self.q
        self.p[:] = self.dinv * self.r
        self.s[0], self.s[1] = self.r @ self.p, self.r @ self.r
        return self.s[0], self.s[1]

    def exchange_halo(self):
        pass

    def spmv_dot(self):
        nb = self.nb
        for f in range(self.nf):
            first = True
            for g in range(self.nf):
                B = self.blocks.get((f, g))
                if B is None:
                    continue
                y = B @ self.p[g * nb:(g + 1) * nb]
                if first:
                    self.q[f * nb:(f + 1) * nb] = y
                    first = False
                else:
                    self.q[f * nb:(f + 1) * nb] += y
        self.s[2] = self.p @ self.q

    def axpy_dot(self):
        cur, nxt = (0, 3) if self.flip == 0 else (3, 0)
        a = self.s[cur] / self.s[2]
        self.x += a * self.p
        self.r -= a * self.q
        self.s[nxt], self.s[nxt + 1] = self.r @ (self.dinv * self.r), self.r @ self.r

    def update_p(self):
        cur, nxt = (0, 3) if self.flip == 0 else (3, 0)
        self.p[:] = self.dinv * self.r + (self.s[nxt] / self.s[cur]) * self.p
        self.flip ^= 1

    def read_rz_rr(self):
        cur = 0 if self.flip == 0 else 3
        return self.s[cur], self.s[cur + 1]

    def solution(self):
        return self.x


def test_block_cg_driver_solves_the_constrained_block_system():
    """The Jacobi-CG driver (multigpu.dist_cg) on a block operator with per-block BCs
    reproduces the oracle's LU solution of the constrained system."""
    from tigar_b200.multigpu import dist_cg
    ts, kv, P = curved_patch(2, (6, 5))
    n = ts.ncp
    z = []
    for f in range(2):                      # clamp side 0 of direction 0, both fields
        z += [f * n + d for d in ts.getSideDofs(0, 0, 1)]
    z += [n + d for d in ts.getSideDofs(1, 1, 1)]        # and field 1 on another side
    prob = OP.ElasticityProblem([2, 2], kv, P, MU, LAM, z)
    Uo = prob.solve(body_force)
    C0, b0 = prob.direct_iga(body_force)
    per = MF.split_zero_dofs(np.unique(z), 2, n)
    masks = [np.zeros(n, bool) for _ in range(2)]
    for f in range(2):
        masks[f][per[f]] = True
    blocks = {}
    for f in range(2):
        for g in range(2):
            B = C0[f * n:(f + 1) * n, g * n:(g + 1) * n].tolil()
            B[masks[f], :] = 0.0                          # rows: test field's zero DoFs
            B[:, masks[g]] = 0.0                          # columns: trial field's
            if f == g:
                B[per[f], per[f]] = 1.0
            blocks[(f, g)] = B.tocsr()
    assert abs(sp.bmat([[blocks[(0, 0)], blocks[(0, 1)]], [blocks[(1, 0)], blocks[(1, 1)]]])
               - prob.C).max() == 0.0
    x, its, rel = dist_cg(NumpyBlockOps(blocks, 2, n), prob.b, 1e-13, 0.0, 5000, 5)
    assert rel <= 1e-13 and its < 5000
    assert np.linalg.norm(x - Uo) < 1e-10 * np.linalg.norm(Uo)


def test_vector_function_plumbing():
    import torch
    from tigar_b200 import api as A
    spl = symbolic_spline(2, 2, 3, 5)
    w = A.Function(spl.V)
    assert isinstance(w, A.VectorFunction) and len(w.comps) == 3 and w.iga is None
    assert A.Function(spl.V.sub(1)).ufl_shape == ()
    t = torch.arange(15, dtype=torch.float64)
    w.set_iga(t)
    assert torch.equal(w.iga, t) and torch.equal(w.comps[2].iga, t[10:])
    assert [c.V.nfields for c in w.comps] == [1, 1, 1]
    assert A.split(w)[1] is w.comps[1] and w.sub(0) is w.comps[0]
    w2 = A.Function(spl.V)
    w2.assign(w)
    assert torch.equal(w2.iga, t)
    w3 = A.Function(spl.V)
    w3.assign(2.0 * w - 0.5 * w2)
    assert torch.allclose(w3.iga, 1.5 * t)
    with pytest.raises(ValueError):
        w.set_iga(torch.zeros(16, dtype=torch.float64))
    # the component jets are scalar functions (comp 0): nothing new for the device kernels
    jets = S.jets_of([w[1].a[()].node()])
    assert [j.args for j in jets] == [(w.comps[1].fid, 0, (0, 0, 0))]
    assert w.fid_fields() == {c.fid: i for i, c in enumerate(w.comps)}
