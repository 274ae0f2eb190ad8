"""Golden vector for BASELINE configs[2] at FULL size (demos/biharmonic, C^3 quartic B-spline,
512^2 cells, 266 256 IGA DoFs): the oracle's direct IGA Galerkin system (oracle.pipeline.
Problem.direct_iga, independent of the FE path) solved by a sparse direct LU (scipy SuperLU --
the role of dolfin.solve's default LU, common.py:1255-1256) and, as a cross-check of what the
conditioning (~h^-4) allows, by LAPACK's band Cholesky.  SuperLU needs ~12 min on one core, so
the result is committed as a fixture (float64, 2 MB) instead of being recomputed by the GPU test.

    python tests/golden/gen_cfg3_golden.py [nel=512]
"""
import math
import os
import sys
import time

import numpy as np
import scipy.linalg as sla
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pipeline as OP, bsplines as OB   # noqa: E402

PI = math.pi


def rhs(X):
    """lap(lap(soln)), soln = (cos(pi x)+1)(cos(pi y)+1)  (biharmonic.py:105-107)"""
    cx, cy = np.cos(PI * X[..., 0]), np.cos(PI * X[..., 1])
    return PI ** 4 * (cx * (cy + 1.0) + (cx + 1.0) * cy + 2.0 * cx * cy)


def main(nel):
    kv = [OB.uniform_knots(4, -1.0, 1.0, nel)] * 2
    pr = OP.Problem([4, 4], kv, form="biharmonic", nLayers=2)
    t = time.time()
    C, b = pr.direct_iga(rhs)
    C = OP.apply_bcs_matrix_fast(C.tocsr(), pr.zeroDofs, 1.0)
    b = OP.apply_bcs_vector(b, pr.zeroDofs)
    print("assembled in %.1f s, n = %d, nnz = %d" % (time.time() - t, C.shape[0], C.nnz))
    # band Cholesky (LAPACK dpbsv)
    t = time.time()
    n = C.shape[0]
    Cc = C.tocoo()
    bw = int((Cc.row - Cc.col).max())
    ab = np.zeros((bw + 1, n))
    lower = Cc.row >= Cc.col
    ab[(Cc.row - Cc.col)[lower], Cc.col[lower]] = Cc.data[lower]
    Ub = sla.solveh_banded(ab, b, lower=True)
    del ab
    print("band Cholesky %.1f s (bw = %d), residual %.3e" % (
        time.time() - t, bw, np.linalg.norm(C @ Ub - b) / np.linalg.norm(b)))
    t = time.time()
    U = spla.spsolve(C.tocsc(), b)
    print("sparse LU %.1f s, residual %.3e" % (time.time() - t,
                                               np.linalg.norm(C @ U - b) / np.linalg.norm(b)))
    print("|U_LU - U_chol| / |U_LU| = %.3e" % (np.linalg.norm(U - Ub) / np.linalg.norm(U)))
    out = os.path.join(HERE, "cfg3_biharmonic_%d.npz" % nel)
    np.savez_compressed(out, U_lu=U, U_chol=Ub, nel=nel, bw=bw)
    print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 512)
