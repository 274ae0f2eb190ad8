"""CPU check of the form language on the Kirchhoff-Love St. Venant-Kirchhoff shell of the
reference's kl-shell-svk demo (dynamic-tspline.py:135-247, static part) -- BASELINE
configs[4]'s integrand -- on the Scordelis-Lo roof:

  * the elastic energy W(y) is written exactly as in the demo (shellGeometry, cartesian,
    voigt, material matrix), three displacement fields in homogeneous representation on a
    cubic NURBS surface in 3-D (2 parametric / 3 physical dimensions);
  * residual  = derivative(W, y_hom, z_hom)  (first variation, TestFunction direction),
    tangent   = derivative(residual, y_hom)   (multi-field Gateaux derivative);
  * both are split into (test field, trial field) term lists and integrated on the host
    (test_multifield_cpu.HostIntegrator) -- the same lists the device kernels consume.

Checked: residual = dW/dU and tangent = dR/dU by central differences, symmetry of the
tangent, and the textbook mid-side displacement of the roof (0.3006 for Kirchhoff-Love
theory; load scaled into the linear regime).
"""

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from tigar_b200 import multifield as MF
from oracle import bsplines as OB
from test_multifield_cpu import HostIntegrator, symbolic_spline

E_MOD, NU, H_TH = 4.32e8, 0.0, 0.25


def shell_forms(spline, y_hom, z_hom, load):
    """dynamic-tspline.py:135-247 without inertia and contact."""
    from tigar_b200 import api as A
    from tigar_b200.ufl_lite import (as_vector, as_matrix, inner, cross, sqrt, inv, Constant)
    y = spline.rationalize(y_hom)
    X = spline.F
    x = X + y

    def unit(v):
        return v / sqrt(inner(v, v))

    def shellGeometry(x):
        dxdxi = spline.parametricGrad(x)
        a0 = as_vector([dxdxi[0, 0], dxdxi[1, 0], dxdxi[2, 0]])
        a1 = as_vector([dxdxi[0, 1], dxdxi[1, 1], dxdxi[2, 1]])
        a2 = unit(cross(a0, a1))
        a = as_matrix(((inner(a0, a0), inner(a0, a1)), (inner(a1, a0), inner(a1, a1))))
        deriva2 = spline.parametricGrad(a2)
        b = -as_matrix(((inner(a0, deriva2[:, 0]), inner(a0, deriva2[:, 1])),
                        (inner(a1, deriva2[:, 0]), inner(a1, deriva2[:, 1]))))
        return (a0, a1, a2, a, b)

    A0, A1, A2, Am, B = shellGeometry(X)
    a0, a1, a2, a, b = shellGeometry(x)
    epsilon = 0.5 * (a - Am)
    kappa = B - b

    def cartesian(T, a, a0, a1):
        ac = inv(a)
        a0c = ac[0, 0] * a0 + ac[0, 1] * a1
        a1c = ac[1, 0] * a0 + ac[1, 1] * a1
        e0 = unit(a0)
        e1 = unit(a1 - e0 * inner(a1, e0))
        ea = as_matrix(((inner(e0, a0c), inner(e0, a1c)), (inner(e1, a0c), inner(e1, a1c))))
        ae = ea.T
        return ea * T * ae

    epsilonBar = cartesian(epsilon, Am, A0, A1)
    kappaBar = cartesian(kappa, Am, A0, A1)

    def voigt(T):
        return as_vector([T[0, 0], T[1, 1], 2.0 * T[0, 1]])

    E, nu = Constant(E_MOD), Constant(NU)
    D = (E / (1.0 - nu * nu)) * as_matrix([[1.0, nu, 0.0], [nu, 1.0, 0.0],
                                           [0.0, 0.0, 0.5 * (1.0 - nu)]])
    nBar = H_TH * D * voigt(epsilonBar)
    mBar = (H_TH ** 3) * D * voigt(kappaBar) / 12.0
    Wint = 0.5 * (inner(voigt(epsilonBar), nBar) + inner(voigt(kappaBar), mBar)) * spline.dx
    z = spline.rationalize(z_hom)
    dWint = Constant(1.0) * A.derivative(Wint, y_hom, z_hom)
    dWext = inner(-as_vector([0.0, 0.0, load]), z) * spline.dx
    res = dWint + dWext
    dRes = A.derivative(res, y_hom)
    return Wint, res, dRes


class Roof(object):
    def __init__(self, nel, load):
        from tigar_b200 import api as A
        from tigar_b200.nurbs import cylindrical_roof, NURBSControlMesh
        nrb = cylindrical_roof(3, nel)
        cm = NURBSControlMesh(nrb)
        self.P = cm.controlNet()
        kv = [np.asarray(k, dtype=float) for k in nrb.knots]
        self.ts = OB.TensorSpline([3, 3], kv)
        self.n = self.ts.ncp
        assert self.P.shape == (self.n, 4)
        self.spl = symbolic_spline(2, 3, 3, self.n)
        self.y = A.Function(self.spl.V)
        self.z = A.TestFunction(self.spl.V)
        self.W, self.res, self.dres = shell_forms(self.spl, self.y, self.z, load)
        funcs = {f.fid: self.P[:, i].copy() for i, f in enumerate(self.spl.cpFuncs)}
        self.H = HostIntegrator(self.ts, self.P, 4, funcs, order=2)
        self.rterms = MF.split_vector_terms(self.spl._weighted(self.res.scalar()), 3)
        self.jterms = MF.split_matrix_terms(self.spl._weighted(self.dres.scalar()), 3)
        self.wnode = self.spl._weighted(self.W.scalar())[(None, None)]

    def set_state(self, Uv):
        n = self.n
        self.H.funcs.update({c.fid: Uv[i * n:(i + 1) * n] for i, c in enumerate(self.y.comps)})

    def energy(self, Uv, load_vec=None):
        self.set_state(Uv)
        return float(self.H._eval([self.wnode])[0].sum())

    def residual(self, Uv):
        self.set_state(Uv)
        return np.concatenate([self.H.vector(self.rterms[f]) for f in range(3)])

    def tangent(self, Uv):
        self.set_state(Uv)
        grid = [[self.H.matrix(self.jterms[(f, g)]) for g in range(3)] for f in range(3)]
        return sp.bmat(grid, format="csr")


@pytest.fixture(scope="module")
def roof():
    return Roof([4, 4], -90.0e-3)


def test_shell_residual_and_tangent_are_consistent_variations(roof):
    n = roof.n
    rng = np.random.RandomState(3)
    U0 = 1e-3 * rng.rand(3 * n)
    assert sorted(roof.jterms) == [(f, g) for f in range(3) for g in range(3)]
    # second derivatives of both arguments appear (bending): a genuinely 4th-order form
    assert any(sum(k[0]) == 2 and sum(k[1]) == 2 for k in roof.jterms[(2, 2)])
    R0 = roof.residual(U0)
    K0 = roof.tangent(U0)
    assert abs(K0 - K0.T).max() < 1e-9 * abs(K0).max()
    # external work is linear in U: W_ext = f_ext . U with f_ext = R_int - R
    d = rng.rand(3 * n) - 0.5
    eps = 1e-6
    # tangent = dR/dU (central difference along a random direction)
    fd = (roof.residual(U0 + eps * d) - roof.residual(U0 - eps * d)) / (2 * eps)
    assert np.linalg.norm(K0 @ d - fd) < 1e-6 * np.linalg.norm(fd)
    # internal residual = dW/dU: R(U).d - R_ext.d, with R_ext = R(0) (W has no load term)
    Rext = roof.residual(np.zeros(3 * n))
    dW = (roof.energy(U0 + eps * d) - roof.energy(U0 - eps * d)) / (2 * eps)
    assert abs((R0 - Rext) @ d - dW) < 1e-6 * abs(dW)
    # the undeformed roof is stress free
    assert abs(roof.energy(np.zeros(3 * n))) < 1e-9 * abs(roof.energy(U0))


def test_scordelis_lo_midside_displacement():
    """Newton on the host with the product's residual / tangent term lists: rigid
    diaphragms at the curved ends (u_x = u_z = 0), free straight edges, gravity load
    90 per unit area (scaled by 1e-3 into the linear regime).  Kirchhoff-Love reference
    value of the vertical mid-side displacement: 0.3006."""
    scale = 1e-3
    r = Roof([6, 6], -90.0 * scale)
    n, ts = r.n, r.ts
    z = []
    for side in (0, 1):
        dofs = ts.getSideDofs(1, side, 1)
        z += [0 * n + d for d in dofs] + [2 * n + d for d in dofs]
    n0 = ts.splines[0].ncp
    z += [1 * n + (n0 // 2)]                      # pin the axial rigid-body translation
    z = np.unique(z)
    free = np.setdiff1d(np.arange(3 * n), z)
    Uv = np.zeros(3 * n)
    for it in range(6):
        R = r.residual(Uv)
        if it == 0:
            r0 = np.linalg.norm(R[free])
        if np.linalg.norm(R[free]) < 1e-7 * r0:        # round-off floor ~ 5e-9 * r0
            break
        K = r.tangent(Uv).tocsr()
        dU = spla.spsolve(K[free][:, free].tocsc(), R[free])
        Uv[free] -= dU
    assert it <= 3
    # rationalised z-displacement at the mid-side of a free straight edge (angular end
    # i = 0, axial parameter 0.5); 6 x 6 cubic elements give 0.2978, 10 x 10 give 0.3004
    from oracle import assembly as OA
    s1 = ts.splines[1]
    span = int(s1.getKnotSpan(0.5))
    N = OA.bspline_ders(s1.ghostKnots, s1.p, span + s1.nGhost, 0.5, 0)[0]
    idx = (span - s1.p + np.arange(s1.p + 1)) * n0 + 0        # i = 0 (edge), j over the span
    w = r.P[idx, 3]
    uz = (N * Uv[2 * n + idx]).sum() / (N * w).sum()
    assert abs(abs(uz) / scale - 0.3006) < 0.02 * 0.3006, uz / scale


def test_shell_tangent_block_kernel_compiles_for_sm100a(roof):
    """The Gauss-point program of one tangent block (36 coefficient outputs, ~4 300
    operations, 335 registers, second-derivative jets of 7 functions) goes through the form
    compiler's CUDA generator and NVRTC for sm_100a (no GPU needed to compile)."""
    from tigar_b200 import symbolic as S
    from tigar_b200 import jit
    bt = roof.jterms[(2, 0)]
    alS = sorted(set(k[0] for k in bt))
    alT = sorted(set(k[1] for k in bt))
    grid = [[S.ZERO] * len(alT) for _ in alS]
    for (a, b), node in bt.items():
        grid[alS.index(a)][alT.index(b)] = node
    prog = S.compile_program([grid[s][t] for s in range(len(alS)) for t in range(len(alT))], 2)
    assert len(prog.outregs) == 36 and prog.nreg <= 2048       # interpreter limit (tg_qp.cu)
    fids = sorted(set(j[0] for j in prog.jets))
    assert len(fids) == 7 <= jit.MAXFUN
    assert max(max(al) for (_, _, al) in prog.jets) == 2
    jets = [(fids.index(f), c, al) for (f, c, al) in prog.jets]
    src, nth = jit.generate(prog, 2, [4, 4, 1], [4, 4, 1], 3, jets, len(fids))
    assert nth == 32 and jit.check_source(src) > 1000


def test_generated_cuda_of_a_shell_tangent_block_runs_correctly_on_host_threads(roof):
    """The CUDA source the form compiler generates for one tangent block of the shell
    (jit.generate: sum-factorised second-derivative jets of 7 functions through shared
    memory, ~4 300-operation straight-line program, 36 outputs per Gauss point) is compiled
    UNCHANGED for the host (tests/cuda_emu.py: CUDA threads -> std::threads, __syncthreads
    -> std::barrier) and must reproduce the host interpreter's coefficients at every Gauss
    point of every cell."""
    import cuda_emu
    from tigar_b200 import symbolic as S
    from tigar_b200 import jit
    from oracle import assembly as OA
    bt = roof.jterms[(0, 2)]
    alS = sorted(set(k[0] for k in bt))
    alT = sorted(set(k[1] for k in bt))
    grid = [[S.ZERO] * len(alT) for _ in alS]
    for (a, b), node in bt.items():
        grid[alS.index(a)][alT.index(b)] = node
    outputs = [grid[s][t] for s in range(len(alS)) for t in range(len(alT))]
    prog = S.compile_program(outputs, 2)
    fids = sorted(set(j[0] for j in prog.jets))
    jets = [(fids.index(f), c, al) for (f, c, al) in prog.jets]
    nd = 3                                                     # derivative orders 0..2
    src, nth = jit.generate(prog, 2, [4, 4, 1], [4, 4, 1], nd, jets, len(fids))
    rng = np.random.RandomState(11)
    Uv = 1e-2 * rng.rand(3 * roof.n)
    roof.set_state(Uv)
    tabs = [OA.tab_iga(s, 4, nd - 1) for s in roof.ts.splines]
    ncell = int(np.prod([tb.T.shape[0] for tb in tabs]))
    got = cuda_emu.run_qp_kernel(src, nth, tabs, [roof.H.funcs[f] for f in fids],
                                 len(outputs), ncell)
    ref = np.stack(roof.H._eval(outputs), axis=1)              # [cell, output, qp]
    scale = np.abs(ref).max()
    assert scale > 0 and np.abs(got - ref).max() < 1e-11 * scale
