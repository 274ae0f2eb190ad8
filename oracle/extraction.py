"""
Oracle: background FE mesh and extraction operator M (numpy/scipy, CPU).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows BSplines.py:505-588 (one rectangular cell per non-degenerate knot
span, FE degree = max p) and common.py:321-383, 1460-1578 (row-wise M with the
1e-15 filter; cpFuncs = M_control * P).  What the reference delegates to
DOLFIN and we must choose: continuous Q_pf Lagrange elements
on equispaced nodes, lexicographic node numbering (first direction fastest).
"""
import numpy as np
import scipy.sparse as sp

from .bsplines import IGNORE_EPS


def fe_nodes_1d(spline1, pf):
    """1-D FE node coordinates for CG Q_pf on the one-cell-per-span mesh.
    Node g = e*pf + a;  a=0 / a=pf sit exactly on the unique knots, interior
    nodes are uk[e] + (a*h)/pf (this exact operation order is part of the
    parity contract with the CUDA path)."""
    uk = spline1.uniqueKnots
    nel = spline1.nel
    x = np.empty(nel * pf + 1)
    x[0::pf] = uk
    for a in range(1, pf):
        h = uk[1:] - uk[:-1]
        x[a::pf] = uk[:-1] + (a * h) / pf
    return x


def m1d(spline1, pf, eps=IGNORE_EPS):
    """1-D extraction matrix [n_fe x ncp] (CSR) evaluated exactly as the
    reference does per node: left span via searchsorted, Cox-de Boor, drop
    |v|<=eps (BSplines.py:285-351, common.py:1507-1509)."""
    x = fe_nodes_1d(spline1, pf)
    span = spline1.spans_vec(x)
    vals = spline1.basis_vec(span, x)                       # [n, p+1]
    p = spline1.p
    cols = (span[:, None] - p + np.arange(p + 1)[None, :]) % spline1.ncp
    rows = np.repeat(np.arange(len(x)), p + 1)
    keep = np.abs(vals.ravel()) > eps
    M = sp.csr_matrix((vals.ravel()[keep], (rows[keep], cols.ravel()[keep])),
                      shape=(len(x), spline1.ncp))
    M.sort_indices()
    return M


def build_M_kron(tspline, eps=IGNORE_EPS):
    """M = M_w (x) M_v (x) M_u.  NB: the reference filters the
    *product* against eps, not the factors; with factors in (1e-15,1] a product
    could in principle drop below eps.  ``build_M_loop`` is the literal
    statement; tests check both agree on every case used."""
    pf = tspline.getDegree()
    Ms = [m1d(s, pf, eps) for s in tspline.splines]
    M = Ms[0]
    for d in range(1, tspline.nvar):
        M = sp.kron(Ms[d], M, format="csr")
    M = M.tocsr()
    M.data[np.abs(M.data) <= eps] = 0.0
    M.eliminate_zeros()
    M.sort_indices()
    return M


def fe_node_coords(tspline):
    """[n_nodes, nvar] lexicographic FE node coordinates."""
    pf = tspline.getDegree()
    xs = [fe_nodes_1d(s, pf) for s in tspline.splines]
    grids = np.meshgrid(*xs, indexing="ij")
    # first direction fastest -> Fortran-order ravel
    return np.stack([g.ravel(order="F") for g in grids], axis=1)


def build_M_loop(tspline, eps=IGNORE_EPS):
    """Literal restatement of common.py:1497-1509 / 1554-1571: Python loop
    over FE nodes -> getNodesAndEvals -> scalar insertion with eps filter."""
    X = fe_node_coords(tspline)
    M = sp.lil_matrix((X.shape[0], tspline.getNcp()))
    for I in range(X.shape[0]):
        for col, val in tspline.getNodesAndEvals(X[I]):
            if abs(val) > eps:
                M[I, col] = val
    M = M.tocsr()
    M.sort_indices()
    return M


def control_funcs(M_control, P):
    """cpFuncs[i] = M_control * P[:, i]  (common.py:367-380)."""
    return M_control @ P


def n_fe_nodes(tspline):
    pf = tspline.getDegree()
    return [s.nel * pf + 1 for s in tspline.splines]
