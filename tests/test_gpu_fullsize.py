"""Parity at BASELINE.json's full size (3-D cubic Poisson, 256^3 cells, 17.4 M
DoFs, element-fused path) through size-independent properties, since the
oracle cannot hold that problem: constants are in the kernel of the stiffness
matrix (partition of unity of the extracted basis), symmetry, the load vector
sums to the integral of f, the CG residual recomputed independently, and the
manufactured-solution error is at the O(h^4) level."""
import math

import numpy as np
import pytest

from gpu_util import make_pair, uk

pytestmark = pytest.mark.gpu
PI = math.pi


@pytest.mark.parametrize("nel", [96, 256])
def test_full_size_properties(nel):
    import torch
    from tIGAr import TrialFunction, TestFunction, Function, inner, sin, assemble
    from tigar_b200 import dev
    free, _ = torch.cuda.mem_get_info()
    if nel == 256 and free < 90e9:
        pytest.skip("needs ~80 GB of free HBM")
    from tIGAr import EqualOrderSpline, ExtractedSpline
    from tIGAr.BSplines import ExplicitBSplineControlMesh
    p = 3
    kv = [uk(p, nel)] * 3
    gen = EqualOrderSpline(1, ExplicitBSplineControlMesh([p] * 3, kv))
    sp = gen.getScalarSpline(0)
    for d in range(3):
        for side in (0, 1):
            gen.addZeroDofs(0, sp.getSideDofs(d, side))
    spline = ExtractedSpline(gen, 2 * p, mode="fused")
    u, v = TrialFunction(spline.V), TestFunction(spline.V)
    x = spline.spatialCoordinates()
    soln = sin(PI * x[0]) * sin(PI * x[1]) * sin(PI * x[2])
    f = 3 * PI ** 2 * soln
    a = inner(spline.grad(u), spline.grad(v)) * spline.dx
    L = inner(f, v) * spline.dx
    n = spline.patch().n_iga
    assert n == (nel + p) ** 3
    # --- without BCs
    C0 = spline.assembleMatrix(a, applyBCs=False)
    b0 = spline.assembleVector(L, applyBCs=False).t
    ones = torch.ones(n, dtype=torch.float64, device=b0.device)
    y = C0.matvec(ones)
    scale = float(C0.vals.abs().max())
    assert float(y.abs().max()) < 1e-11 * scale * 343          # C 1 = 0
    g = torch.Generator(device=b0.device).manual_seed(1)
    r1 = torch.rand(n, dtype=torch.float64, device=b0.device, generator=g)
    r2 = torch.rand(n, dtype=torch.float64, device=b0.device, generator=g)
    s12 = float(torch.dot(r1, C0.matvec(r2)))
    s21 = float(torch.dot(r2, C0.matvec(r1)))
    assert abs(s12 - s21) < 1e-11 * abs(s12)                    # symmetry
    # sum of the load vector = int f = 3 pi^2 (2/pi)^3 (Gauss rule is near-exact for sin)
    assert abs(float(b0.sum()) - 24.0 / PI) < 1e-6
    del C0, y, r1, r2
    # --- with BCs: solve and verify the residual with an independent product
    C, b = spline.assembleLinearSystem(a, L)
    uh = Function(spline.V)
    from tIGAr import KrylovSolver
    ks = KrylovSolver("cg", "jacobi")
    ks.parameters["relative_tolerance"] = 1e-12
    spline.setSolverOptions(linearSolver=ks)
    U = spline.solveLinearSystem(C, b, uh)
    res = b.t - C.matvec(U.t)
    relres = float(res.norm() / b.t.norm())
    assert relres < 5e-12, relres
    z = torch.from_numpy(np.unique(spline.zeroDofs)).to(b.t.device)
    assert float(U.t[z].abs().max()) == 0.0                     # homogeneous Dirichlet rows
    del C, res
    err = math.sqrt(assemble(((uh - soln) ** 2) * spline.dx))
    h = 1.0 / nel
    # optimal order for p=3 (oracle: err ~ 0.055 h^4), or the floor set by solving an
    # ill-conditioned system (cond ~ h^-2) to 1e-12 in double precision
    assert err < max(0.2 * h ** 4, 2e-9), (err, h ** 4)


@pytest.mark.parametrize("nel", [40, 64])
def test_full_size_global_ptap_properties(nel):
    """Global-CSR M^T A M (MatPtAP, common.py:1194-1195) at sizes scipy cannot
    check: a checksum of checksums through INDEPENDENT kernels,
        C x = M^T (A (M x))   with tg_win_spmv (M x, A .) and tg_mt_vec (M^T .),
    partition of unity (M 1 = 1, so C 1 = M^T (A 1)), symmetry of C for a
    symmetric A, and agreement of the march passes with the row-wise Kronecker
    kernels.  64^3: A = 7.1 GB (0.89 G non-zeros), C = 0.76 GB."""
    import torch
    from tigar_b200.engine import TensorPatch, WinMatrix
    from tigar_b200 import dev
    free, _ = torch.cuda.mem_get_info()
    if nel == 64 and free < 60e9:
        pytest.skip("needs ~50 GB of free HBM")
    p = 3
    patch = TensorPatch([p] * 3, [uk(p, nel)] * 3)
    assert patch._march_setup() is not None
    # the real stiffness-like operand: A_FE of the mass form is symmetric and cheap to
    # get at this size only through the assembly kernels; use a synthetic symmetric
    # windowed A instead: A = B + B^T is not expressible without a transpose kernel, so
    # take A = diag-scaled constant pattern D S D with S = all-ones window (symmetric)
    A = WinMatrix(patch.window("A"))
    g = torch.Generator(device=A.vals.device).manual_seed(3)
    dsc = 0.5 + torch.rand(patch.n_fe, dtype=torch.float64, device=A.vals.device, generator=g)
    A.vals.fill_(1.0)
    # rows scaled by d_I and columns by d_J: apply through two SpMV-free passes on values
    rl = dev.empty(patch.n_fe, dev.I64)
    from tigar_b200._lib import lib, check
    check(lib.tg_win_rowlen(A.window.ref(), dev.ptr(rl), dev.stream()))
    A.vals.mul_(torch.repeat_interleave(dsc, rl))
    cols = A.window.columns()
    A.vals.mul_(dsc[cols.long()])
    del cols, rl
    M = patch.build_M()
    C = patch.ptap_march(A)
    n = patch.n_iga
    ones = torch.ones(n, dtype=torch.float64, device=A.vals.device)
    x = torch.rand(n, dtype=torch.float64, device=A.vals.device, generator=g)
    y = torch.rand(n, dtype=torch.float64, device=A.vals.device, generator=g)
    # partition of unity of the extraction operator
    assert float((M.matvec(ones) - 1.0).abs().max()) < 1e-13
    for v in (ones, x):
        ref = patch.mt_vec(M, A.matvec(M.matvec(v)))
        got = C.matvec(v)
        assert float((got - ref).norm() / ref.norm()) < 1e-13
    sxy = float(torch.dot(x, C.matvec(y)))
    syx = float(torch.dot(y, C.matvec(x)))
    assert abs(sxy - syx) < 1e-12 * abs(sxy)
    if nel <= 48:
        Ck = patch.ptap_kron(A)
        d = (C.csr_values() - Ck.csr_values()).abs().max() / Ck.csr_values().abs().max()
        assert float(d) < 1e-13
