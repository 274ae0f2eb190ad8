"""
Oracle: Gauss-point assembly of A_FE / b_FE on the Lagrange background mesh,
and (independently) of the IGA Galerkin matrix directly in the B-spline basis.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

What the kernels must reproduce per Gauss point is defined by
common.py:896-945 and calculusUtils.py:18-24, 56-69, 255-276, 351-410:
  F = cp[0:nsd]/cp[nsd]; DF = grad F; g = DF^T DF; J = sqrt(det g);
  pinvDF = g^-1 DF^T; grad_x f = (df/dxi) . pinvDF; div_x = trace;
  integral = sum_q w_q J_q (.)       (Gauss-Legendre, floor(quadDeg/2)+1 / dir)
The assembly itself is DOLFIN's in the reference (common.py:1206-1220,
1162-1173); this file is the restatement, pinned by the KATs of SURVEY 8c.
"""
import itertools
import numpy as np
import scipy.sparse as sp


# ---------------------------------------------------------------- 1-D tables
def gauss_rule(nq):
    """Gauss-Legendre on [0,1]."""
    x, w = np.polynomial.legendre.leggauss(nq)
    return 0.5 * (x + 1.0), 0.5 * w


def lagrange_jets(pf, t, nder):
    """Equispaced Lagrange basis of degree pf on [0,1]: [len(t), pf+1, nder+1]."""
    nodes = np.arange(pf + 1) / float(pf)
    out = np.zeros((len(t), pf + 1, nder + 1))
    for a in range(pf + 1):
        others = np.delete(nodes, a)
        poly = np.poly1d(others, r=True) / np.prod(nodes[a] - others)
        for k in range(nder + 1):
            out[:, a, k] = np.polyder(poly, k)(t) if k > 0 else poly(t)
    return out


def bspline_ders(knots, p, span, u, nder):
    """Piegl-Tiller A2.3 (derivatives of the p+1 non-zero functions); used
    only by the independent direct-IGA cross-check."""
    ndu = np.zeros((p + 1, p + 1))
    left = np.zeros(p + 1)
    right = np.zeros(p + 1)
    ndu[0, 0] = 1.0
    for j in range(1, p + 1):
        left[j] = u - knots[span + 1 - j]
        right[j] = knots[span + j] - u
        saved = 0.0
        for r in range(j):
            ndu[j, r] = right[r + 1] + left[j - r]
            temp = ndu[r, j - 1] / ndu[j, r]
            ndu[r, j] = saved + right[r + 1] * temp
            saved = left[j - r] * temp
        ndu[j, j] = saved
    ders = np.zeros((nder + 1, p + 1))
    ders[0] = ndu[:, p]
    a = np.zeros((2, p + 1))
    for r in range(p + 1):
        s1, s2 = 0, 1
        a[0, 0] = 1.0
        for k in range(1, min(nder, p) + 1):
            d = 0.0
            rk, pk = r - k, p - k
            if r >= k:
                a[s2, 0] = a[s1, 0] / ndu[pk + 1, rk]
                d = a[s2, 0] * ndu[rk, pk]
            j1 = 1 if rk >= -1 else -rk
            j2 = k - 1 if r - 1 <= pk else p - r
            for j in range(j1, j2 + 1):
                a[s2, j] = (a[s1, j] - a[s1, j - 1]) / ndu[pk + 1, rk + j]
                d += a[s2, j] * ndu[rk + j, pk]
            if r <= pk:
                a[s2, k] = -a[s1, k - 1] / ndu[pk + 1, r]
                d += a[s2, k] * ndu[r, pk]
            ders[k, r] = d
            s1, s2 = s2, s1
    r = p
    for k in range(1, min(nder, p) + 1):
        ders[k] *= r
        r *= (p - k)
    return ders                                   # [nder+1, p+1]


class Tab1D(object):
    """Per-direction tables: T[e,q,a,k] = d^k basis_a / dxi^k at Gauss point q
    of element e; idx[e,a] global 1-D index; w[e,q] = weight * h_e; x[e,q]."""
    pass


def tab_fe(spline1, pf, nq, nder):
    uk = spline1.uniqueKnots
    h = uk[1:] - uk[:-1]
    t, w = gauss_rule(nq)
    L = lagrange_jets(pf, t, nder)                                # [q,a,k]
    tb = Tab1D()
    scale = (1.0 / h)[:, None] ** np.arange(nder + 1)[None, :]     # [e,k]
    tb.T = L[None, :, :, :] * scale[:, None, None, :]
    tb.idx = np.arange(spline1.nel)[:, None] * pf + np.arange(pf + 1)[None, :]
    tb.w = w[None, :] * h[:, None]
    tb.x = uk[:-1, None] + t[None, :] * h[:, None]
    tb.n = spline1.nel * pf + 1
    return tb


def tab_iga(spline1, nq, nder):
    uk = spline1.uniqueKnots
    h = uk[1:] - uk[:-1]
    t, w = gauss_rule(nq)
    p = spline1.p
    spans = spline1.element_spans()
    tb = Tab1D()
    tb.T = np.zeros((spline1.nel, nq, p + 1, nder + 1))
    tb.x = uk[:-1, None] + t[None, :] * h[:, None]
    # ghost knots make periodic / out-of-range indices valid
    gk = spline1.ghostKnots
    nG = spline1.nGhost
    for e in range(spline1.nel):
        for q in range(nq):
            d = bspline_ders(gk, p, int(spans[e]) + nG, tb.x[e, q], nder)
            tb.T[e, q] = d.T
    tb.idx = (spans[:, None] - p + np.arange(p + 1)[None, :]) % spline1.ncp
    tb.w = w[None, :] * h[:, None]
    tb.n = spline1.ncp
    return tb


# ------------------------------------------------------ multi-index helpers
def multi_indices(dim, order):
    out = []
    for tot in range(order + 1):
        for al in itertools.product(range(tot + 1), repeat=dim):
            if sum(al) == tot:
                out.append(al)
    return out


def unit(dim, j):
    return tuple(1 if d == j else 0 for d in range(dim))


def add(a, b):
    return tuple(x + y for x, y in zip(a, b))


class CellBlock(object):
    """Basis jets for a chunk of cells.  jets[alpha] -> [nc, nq, nen];
    gidx [nc, nen]; wq [nc, nq]; xi [nc, nq, dim]."""

    def __init__(self, tabs, cells, order):
        dim = len(tabs)
        self.dim = dim
        nels = [tb.T.shape[0] for tb in tabs]
        e = np.unravel_index(cells, nels, order="F")           # first dir fastest
        self.jets = {}
        for al in multi_indices(dim, order):
            J = tabs[0].T[e[0], :, :, al[0]]                    # [nc,q0,a0]
            for d in range(1, dim):
                Td = tabs[d].T[e[d], :, :, al[d]]               # [nc,qd,ad]
                J = np.einsum("cqa,crb->crqba", J, Td).reshape(
                    len(cells), Td.shape[1] * J.shape[1], Td.shape[2] * J.shape[2])
            self.jets[al] = J
        g = tabs[0].idx[e[0]]
        w = tabs[0].w[e[0]]
        stride = tabs[0].n
        xs = [tabs[0].x[e[0]]]
        for d in range(1, dim):
            gd = tabs[d].idx[e[d]]
            g = (gd[:, :, None] * stride + g[:, None, :]).reshape(len(cells), -1)
            w = (tabs[d].w[e[d]][:, :, None] * w[:, None, :]).reshape(len(cells), -1)
            stride *= tabs[d].n
            xs = [np.broadcast_to(x[:, None, :], (len(cells), tabs[d].x.shape[1], x.shape[1]))
                  .reshape(len(cells), -1) for x in xs]
            xs.append(np.broadcast_to(tabs[d].x[e[d]][:, :, None],
                                      (len(cells), tabs[d].x.shape[1], w.shape[1] // tabs[d].x.shape[1]))
                      .reshape(len(cells), -1))
        self.gidx = g
        self.wq = w
        self.xi = np.stack(xs, axis=-1)


class Geometry(object):
    """Per-Gauss-point geometry from homogeneous control functions
    (common.py:917-945, calculusUtils.py:18-24,56-69)."""

    def __init__(self, blk, coef, order):
        # coef: [n_global, nsd+1] coefficients of cp functions in blk's basis
        dim = blk.dim
        nsd = coef.shape[1] - 1
        cc = coef[blk.gidx]                                      # [nc,nen,nsd+1]
        X = {al: np.einsum("cqa,cai->cqi", blk.jets[al], cc)
             for al in multi_indices(dim, order)}
        z = (0,) * dim
        w = X[z][..., nsd]
        self.w = w
        self.F = X[z][..., :nsd] / w[..., None]
        self.dw = np.stack([X[unit(dim, j)][..., nsd] for j in range(dim)], -1)   # [c,q,j]
        # DF[c,q,i,j] = dF_i/dxi_j
        self.DF = np.stack([(X[unit(dim, j)][..., :nsd] - self.F * self.dw[..., j, None]) / w[..., None]
                            for j in range(dim)], -1)
        self.g = np.einsum("cqij,cqik->cqjk", self.DF, self.DF)
        self.ginv = np.linalg.inv(self.g)
        self.J = np.sqrt(np.linalg.det(self.g))
        self.P = np.einsum("cqjk,cqik->cqji", self.ginv, self.DF)                 # [c,q,j,i]
        if order >= 2:
            self.d2w = np.empty(w.shape + (dim, dim))
            D2F = np.empty(w.shape + (nsd, dim, dim))
            for j in range(dim):
                for k in range(dim):
                    al = add(unit(dim, j), unit(dim, k))
                    self.d2w[..., j, k] = X[al][..., nsd]
                    D2F[..., :, j, k] = (X[al][..., :nsd]
                                         - self.DF[..., :, j] * self.dw[..., k, None]
                                         - self.DF[..., :, k] * self.dw[..., j, None]
                                         - self.F * X[al][..., nsd, None]) / w[..., None]
            self.D2F = D2F
            # dg[c,q,j,k,l] = d g_jk / dxi_l
            dg = (np.einsum("cqijl,cqik->cqjkl", D2F, self.DF)
                  + np.einsum("cqij,cqikl->cqjkl", self.DF, D2F))
            dginv = -np.einsum("cqja,cqabl,cqbk->cqjkl", self.ginv, dg, self.ginv)
            # dP[c,q,j,i,l] = d P_ji / dxi_l
            self.dP = (np.einsum("cqjkl,cqik->cqjil", dginv, self.DF)
                       + np.einsum("cqjk,cqikl->cqjil", self.ginv, D2F))


def rationalized_jets(blk, geo, order):
    """jets of psi = phi / w   (common.py:1134-1139 applied to trial/test)."""
    dim = blk.dim
    z = (0,) * dim
    out = {}
    w = geo.w[..., None]
    out[z] = blk.jets[z] / w
    if order >= 1:
        for j in range(dim):
            out[unit(dim, j)] = (blk.jets[unit(dim, j)] - out[z] * geo.dw[..., j, None]) / w
    if order >= 2:
        for j in range(dim):
            for k in range(dim):
                al = add(unit(dim, j), unit(dim, k))
                out[al] = (blk.jets[al]
                           - out[unit(dim, j)] * geo.dw[..., k, None]
                           - out[unit(dim, k)] * geo.dw[..., j, None]
                           - out[z] * geo.d2w[..., j, k, None]) / w
    return out


def wdot(W, La, Lb):
    """K[c,a,b] = sum_q W[c,q] La[c,q,a] Lb[c,q,b] as one batched GEMM per chunk (BLAS;
    the same contraction written as an einsum runs ~6x slower)."""
    return np.matmul((La * W[..., None]).transpose(0, 2, 1), Lb)


def operators(form, blk, geo, jets):
    """Return list of per-basis operator arrays L_s[c,q,a] such that the
    bilinear form is sum_s int L_s(u) L_s(v) J dxi."""
    dim = blk.dim
    z = (0,) * dim
    if form == "mass":
        return [jets[z]]
    if form == "poisson":
        # grad_x psi_i = sum_j d_j psi P_ji      (calculusUtils.py:255-261)
        nsd = geo.P.shape[-1]
        return [sum(jets[unit(dim, j)] * geo.P[..., j, i, None] for j in range(dim))
                for i in range(nsd)]
    if form == "biharmonic":
        # lap psi = sum_{i,k} d_k( sum_j d_j psi P_ji ) P_ki (calculusUtils.py:269-276 on :255-261)
        nsd = geo.P.shape[-1]
        lap = 0.0
        for j in range(dim):
            for k in range(dim):
                al = add(unit(dim, j), unit(dim, k))
                cjk = np.einsum("cqi,cqi->cq", geo.P[..., j, :], geo.P[..., k, :])
                lap = lap + jets[al] * cjk[..., None]
            bj = np.einsum("cqik,cqki->cq", geo.dP[..., j, :, :], geo.P)
            lap = lap + jets[unit(dim, j)] * bj[..., None]
        return [lap]
    raise ValueError(form)


FORM_ORDER = {"mass": 0, "poisson": 1, "biharmonic": 2}


def assemble(tabs, coef, form, f=None, rationalize=False, chunk=2048):
    """Assemble (A, b) in the basis described by ``tabs``.  ``f(x)`` maps
    physical points [...,nsd] -> values, b_a = int f psi_a J."""
    dim = len(tabs)
    order = max(FORM_ORDER[form], 1)
    if rationalize and form == "biharmonic":
        order = 2
    ncell = int(np.prod([tb.T.shape[0] for tb in tabs]))
    ntot = int(np.prod([tb.n for tb in tabs]))
    rows, cols, vals = [], [], []
    b = np.zeros(ntot)
    for c0 in range(0, ncell, chunk):
        cells = np.arange(c0, min(ncell, c0 + chunk))
        blk = CellBlock(tabs, cells, order)
        geo = Geometry(blk, coef, order if form == "biharmonic" else 1)
        jets = rationalized_jets(blk, geo, order) if rationalize else blk.jets
        W = blk.wq * geo.J
        Ke = 0.0
        for Ls in operators(form, blk, geo, jets):
            Ke = Ke + wdot(W, Ls, Ls)
        nen = blk.gidx.shape[1]
        rows.append(np.repeat(blk.gidx, nen, axis=1).ravel())
        cols.append(np.tile(blk.gidx, (1, nen)).ravel())
        vals.append(Ke.ravel())
        if f is not None:
            fe = np.einsum("cq,cqa->ca", W * f(geo.F), jets[(0,) * dim])
            np.add.at(b, blk.gidx.ravel(), fe.ravel())
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(ntot, ntot)).tocsr()
    A.sort_indices()
    return A, b


def assemble_elasticity(tabs, coef, mu, lam, f=None, chunk=2048):
    """Equal-order multi-field system (MixedElement with one sub-element per field,
    common.py:337-351; field-major IGA numbering ``globalDof``, common.py:254-262):
    plane / 3-D linear elasticity on the mapped patch,

        a(u, v) = int [ 2 mu eps(u):eps(v) + lam div(u) div(v) ] J dxi ,
        L(v) = int f . v J dxi ,    eps = sym(grad_x) ,

    with grad_x as in calculusUtils.py:255-261.  Block (i, k) of the matrix couples
    test field i with trial field k; all blocks share the scalar pattern.  Returns
    (A [nf*n, nf*n] CSR, b [nf*n]) in the basis described by ``tabs``."""
    dim = len(tabs)
    nsd = coef.shape[1] - 1
    ncell = int(np.prod([tb.T.shape[0] for tb in tabs]))
    n = int(np.prod([tb.n for tb in tabs]))
    rows, cols, vals = [], [], []
    b = np.zeros(nsd * n)
    for c0 in range(0, ncell, chunk):
        cells = np.arange(c0, min(ncell, c0 + chunk))
        blk = CellBlock(tabs, cells, 1)
        geo = Geometry(blk, coef, 1)
        G = operators("poisson", blk, geo, blk.jets)        # G[i][c,q,a] = d psi_a / d x_i
        W = blk.wq * geo.J
        gg = sum(wdot(W, G[j], G[j]) for j in range(nsd))
        nen = blk.gidx.shape[1]
        r0 = np.repeat(blk.gidx, nen, axis=1).ravel()
        c0_ = np.tile(blk.gidx, (1, nen)).ravel()
        for i in range(nsd):
            for k in range(nsd):
                # 2 mu eps(u):eps(v) = mu (grad u : grad v + grad u : grad v^T)
                Ke = mu * wdot(W, G[k], G[i]) + lam * wdot(W, G[i], G[k])
                if i == k:
                    Ke = Ke + mu * gg
                rows.append(r0 + i * n)
                cols.append(c0_ + k * n)
                vals.append(Ke.ravel())
        if f is not None:
            fv = f(geo.F)                                     # [c,q,nsd]
            z = (0,) * dim
            for i in range(nsd):
                fe = np.einsum("cq,cqa->ca", W * fv[..., i], blk.jets[z])
                np.add.at(b, blk.gidx.ravel() + i * n, fe.ravel())
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(nsd * n, nsd * n)).tocsr()
    A.sort_indices()
    return A, b


def functional(tabs, coef, U, kind, exact, rationalize=False, chunk=2048):
    """int (u_h - exact)^2 J  (kind='l2', poisson.py:132) or
    int (lap(u_h) - exact)^2 J (kind='energy', biharmonic.py:127), with
    u_h = sum_a U_a psi_a and ``exact(x)`` the exact u / lap u."""
    dim = len(tabs)
    order = 2 if kind == "energy" else 1
    ncell = int(np.prod([tb.T.shape[0] for tb in tabs]))
    tot = 0.0
    for c0 in range(0, ncell, chunk):
        cells = np.arange(c0, min(ncell, c0 + chunk))
        blk = CellBlock(tabs, cells, order)
        geo = Geometry(blk, coef, order)
        jets = rationalized_jets(blk, geo, order) if rationalize else blk.jets
        if kind == "l2":
            L = jets[(0,) * dim]
        else:
            L = operators("biharmonic", blk, geo, jets)[0]
        uh = np.einsum("cqa,ca->cq", L, U[blk.gidx])
        tot += np.sum(blk.wq * geo.J * (uh - exact(geo.F)) ** 2)
    return tot
