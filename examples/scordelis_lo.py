"""
Scordelis-Lo roof with the St. Venant-Kirchhoff Kirchhoff-Love shell formulation of the
reference's demos/kl-shell-svk (dynamic-tspline.py:135-247, static part): cubic NURBS
cylinder segment, three displacement fields in homogeneous representation, energy ->
residual -> tangent by derivative(), Newton (BASELINE configs[4] in small).

The forms are CPU-verified (tests/test_kl_shell_cpu.py: 0.2978 at 6x6, 0.3004 at 10x10
elements against the Kirchhoff-Love value 0.3006) and device-verified
(tests/test_zz_gpu_multifield.py).  The Newton steps use the default solver of the
reference -- a direct solve (common.py:1255-1256) -- which here is the band Cholesky on the
node-major interleaved 3-field system; ``jacobi`` as third argument selects Jacobi-CG.
Usage: python examples/scordelis_lo.py [nel] [load_scale] [jacobi]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tIGAr import *                 # noqa: E402,F401,F403
from tIGAr.NURBS import *           # noqa: E402,F401,F403

nel = int(sys.argv[1]) if len(sys.argv) > 1 else 16
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
E, nu, h_th, load = 4.32e8, 0.0, 0.25, -90.0 * scale

t0 = time.perf_counter()
gen = EqualOrderSpline(3, NURBSControlMesh(cylindrical_roof(3, [nel, nel])))
scalar = gen.getScalarSpline(0)
for side in (0, 1):                                   # rigid diaphragms: u_x = u_z = 0
    gen.addZeroDofs(0, scalar.getSideDofs(1, side))
    gen.addZeroDofs(2, scalar.getSideDofs(1, side))
n0 = scalar.splines[0].getNcp()
gen.addZeroDofs(1, [n0 // 2])                         # axial rigid-body translation
spline = ExtractedSpline(gen, 6, mode="fused")

y_hom = Function(spline.V)
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


A0, A1, A2, A, B = shellGeometry(X)
a0, a1, a2, a, b = shellGeometry(x)
epsilon = 0.5 * (a - A)
kappa = B - b


def cartesian(T, a, a0, a1):
    ac = inv(a)
    a0c = ac[0, 0] * a0 + ac[0, 1] * a1
    a1c = ac[1, 0] * a0 + ac[1, 1] * a1
    e0 = unit(a0)
    e1 = unit(a1 - e0 * inner(a1, e0))
    ea = as_matrix(((inner(e0, a0c), inner(e0, a1c)), (inner(e1, a0c), inner(e1, a1c))))
    return ea * T * ea.T


def voigt(T):
    return as_vector([T[0, 0], T[1, 1], 2.0 * T[0, 1]])


epsilonBar = cartesian(epsilon, A, A0, A1)
kappaBar = cartesian(kappa, A, A0, A1)
D = (E / (1.0 - nu * nu)) * as_matrix([[1.0, nu, 0.0], [nu, 1.0, 0.0],
                                       [0.0, 0.0, 0.5 * (1.0 - nu)]])
nBar = h_th * D * voigt(epsilonBar)
mBar = (h_th ** 3) * D * voigt(kappaBar) / 12.0
Wint = 0.5 * (inner(voigt(epsilonBar), nBar) + inner(voigt(kappaBar), mBar)) * spline.dx
z_hom = TestFunction(spline.V)
z = spline.rationalize(z_hom)
res = derivative(Wint, y_hom, z_hom) - inner(as_vector([0.0, 0.0, load]), z) * spline.dx
dRes = derivative(res, y_hom)

ks = None
if len(sys.argv) > 3 and sys.argv[3] == "jacobi":
    ks = PETScKrylovSolver("cg", "jacobi")
    ks.parameters["relative_tolerance"] = 1e-11
spline.setSolverOptions(maxIters=20, relativeTolerance=1e-6, linearSolver=ks)
t1 = time.perf_counter()
spline.solveNonlinearVariationalProblem(res, dRes, y_hom)
t2 = time.perf_counter()

d0, d1, d2 = y_hom.split()
File("results/disp-z.pvd") << d2
# vertical displacement at the middle of the free edge (xi_0 = 0 side, xi_1 = 0.5)
import numpy as np                  # noqa: E402
n1 = scalar.splines[1].getNcp()
Uz = d2.iga.cpu().numpy().reshape(n1, n0)[:, 0]
Wt = spline.cpFuncs[3].iga.cpu().numpy().reshape(n1, n0)[:, 0]
from scipy.interpolate import BSpline as _BS   # noqa: E402  (1-D post-processing only)
s1 = scalar.splines[1]
uz = float(_BS(s1.knots, Uz, s1.p)(0.5) / _BS(s1.knots, Wt, s1.p)(0.5))
if mpirank == 0:
    print("%dx%d cubic elements, %d DoFs, setup %.1f s, Newton %.1f s (solver: %s); mid-side "
          "z-displacement / load scale = %.5f (Kirchhoff-Love reference 0.3006)"
          % (nel, nel, spline.n_total(), t1 - t0, t2 - t1, spline.lastSolve["method"],
             abs(uz) / scale))
