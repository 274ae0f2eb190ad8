"""CPU tests of the product's host logic (no device work): knot bookkeeping vs
the reference golden, symbolic compiler vs direct evaluation, form language,
C-ABI surface."""
import math
import os
import re

import numpy as np
import pytest

from tigar_b200 import bsplines as PB
from tigar_b200 import symbolic as S
from tigar_b200 import ufl_lite as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_uniform_knots_and_spline1_state(golden):
    for i, a in enumerate(golden["uk_args"]):
        got = PB.uniformKnots(int(a[0]), a[1], a[2], int(a[3]), bool(a[4]), int(a[5]))
        assert np.array_equal(np.array(got, dtype=float), golden["uk_%d" % i])
    for name in golden["s1_names"]:
        pre = "s1_%s_" % name
        s = PB.BSpline1(int(golden[pre + "p"]), golden[pre + "knots"])
        assert s.nel == int(golden[pre + "nel"]) and s.ncp == int(golden[pre + "ncp"])
        assert np.array_equal(s.uniqueKnots, golden[pre + "uniqueKnots"])
        assert np.array_equal(s.multiplicities, golden[pre + "multiplicities"])
        assert np.array_equal(s.ghostKnots, golden[pre + "ghostKnots"])
        assert np.array_equal(s.grevilleAll(), golden[pre + "greville"])
        assert s.isDiscontinuous() == bool(golden[pre + "disc"])


def test_tensor_bookkeeping(golden):
    for name in golden["tp_names"]:
        pre = "tp_%s_" % name
        deg = [int(x) for x in golden[pre + "deg"]]
        kv = [golden[pre + "kv%d" % d] for d in range(len(deg))]
        b = PB.BSpline(deg, kv)
        assert b.getNcp() == int(golden[pre + "ncp"]) and b.nel == int(golden[pre + "nel"])
        assert b.getDegree() == int(golden[pre + "degree"])
        assert b.getPrealloc() == int(golden[pre + "prealloc"])
        for d in range(len(deg)):
            for side in (0, 1):
                for nl in (1, 2):
                    assert np.array_equal(np.array(b.getSideDofs(d, side, nl)),
                                          golden[pre + "side_%d_%d_%d" % (d, side, nl)])
        extra = int(golden[pre + "nsd"]) - len(deg)
        cm = PB.ExplicitBSplineControlMesh(deg, kv, extraDim=extra)
        assert np.array_equal(cm.controlNet(), golden[pre + "P"])
        P1 = np.array([[cm.getHomogeneousCoordinate(n, d) for d in range(cm.getNsd() + 1)]
                       for n in range(0, b.getNcp(), 5)])
        assert np.array_equal(P1, golden[pre + "P"][::5])


# ---- interpreter for compiled programs (test infrastructure) -------------------
def run_program(prog, xi, wq, jetvals):
    R = np.zeros(prog.nreg)
    R[:prog.dim] = xi
    R[prog.dim] = wq
    for k, j in enumerate(prog.jets):
        R[prog.dim + 1 + k] = jetvals[j]
    f1 = dict(neg=lambda a: -a, sin=math.sin, cos=math.cos, exp=math.exp, log=math.log,
              sqrt=math.sqrt, abs=abs, tan=math.tan, tanh=math.tanh, sinh=math.sinh,
              cosh=math.cosh, atan=math.atan, mov=lambda a: a)
    f2 = dict(add=lambda a, b: a + b, sub=lambda a, b: a - b, mul=lambda a, b: a * b,
              div=lambda a, b: a / b, pow=lambda a, b: a ** b, max=max, min=min,
              gt=lambda a, b: float(a > b), selz=lambda a, b: b if a != 0 else 0.0)
    names = {v: k for k, v in S.OPCODES.items()}
    for op, dst, a, b in prog.prog:
        n = names[op]
        if n == "const":
            R[dst] = prog.consts[a]
        elif n in f1:
            R[dst] = f1[n](R[a])
        else:
            R[dst] = f2[n](R[a], R[b])
    return [R[r] for r in prog.outregs]


def test_symbolic_compile_and_diff():
    x, y = S.xi(0), S.xi(1)
    u = S.jet(1, 0, (0, 0, 0))
    e = S.func("sin", x * 3.0) * S.func("exp", y) + u * u / (x + 2.0) - S.power(y + 1.5, S.const(2.5))
    de = S.diff(e, 0)
    prog = S.compile_program([e, de, S.diff(e, 1)], 2)
    jv = {(1, 0, (0, 0, 0)): 0.7, (1, 0, (1, 0, 0)): -0.3, (1, 0, (0, 1, 0)): 1.1}
    X, Y = 0.4, 0.9

    def ev(X, Y, U0):
        return math.sin(3 * X) * math.exp(Y) + U0 * U0 / (X + 2.0) - (Y + 1.5) ** 2.5
    got = run_program(prog, [X, Y], 1.0, jv)
    assert abs(got[0] - ev(X, Y, 0.7)) < 1e-14
    # d/dx with u depending on x through its jet (1,0,0)
    ex = 3 * math.cos(3 * X) * math.exp(Y) + 2 * 0.7 * (-0.3) / (X + 2) - 0.49 / (X + 2) ** 2
    ey = math.sin(3 * X) * math.exp(Y) + 2 * 0.7 * 1.1 / (X + 2) - 2.5 * (Y + 1.5) ** 1.5
    assert abs(got[1] - ex) < 1e-13 and abs(got[2] - ey) < 1e-13
    assert prog.nreg < 40


def test_shared_denominators_are_inverted_once():
    """compile_program: a/w, b/w, c/w share ONE reciprocal (the geometry divides 9-12
    numerators by the weight / determinant); a denominator used once keeps its division, and
    a guarded division by zero still selects the other branch."""
    a, b, c = (S.jet(1, i, (0, 0, 0)) for i in range(3))
    w, d = S.jet(2, 0, (0, 0, 0)), S.xi(0) + 2.0
    outs = [a / w + b / w, c / w, a / d, S.selz(S.binary("gt", w, S.ZERO), a / w)]
    prog = S.compile_program(outs, 1)
    names = {v: k for k, v in S.OPCODES.items()}
    assert sum(names[op] == "div" for op, _, _, _ in prog.prog) == 2        # 1/w and a/d
    jv = {(1, 0, (0, 0, 0)): 0.7, (1, 1, (0, 0, 0)): -1.3, (1, 2, (0, 0, 0)): 2.9,
          (2, 0, (0, 0, 0)): 1.7}
    got = run_program(prog, [0.25], 1.0, jv)
    ref = [0.7 / 1.7 - 1.3 / 1.7, 2.9 / 1.7, 0.7 / 2.25, 0.7 / 1.7]
    assert np.allclose(got, ref, rtol=4e-16, atol=0)
    jv[(2, 0, (0, 0, 0))] = 0.0
    with np.errstate(all="ignore"):
        assert run_program(prog, [0.25], 1.0, jv)[3] == 0.0


def test_hash_consing_and_folding():
    a = S.xi(0) + S.xi(1)
    b = S.xi(1) + S.xi(0)
    assert a is b
    assert S.mul(S.const(2.0), S.const(3.0)) is S.const(6.0)
    assert S.sub(a, a) is S.ZERO and S.mul(a, S.ZERO) is S.ZERO


def test_form_language_poisson_keys():
    # identity geometry written by hand: grad u . grad v
    u = U.Tensor(U.Scalar({(None, U.ZERO3): S.ONE}))
    v = U.Tensor(U.Scalar({(U.ZERO3, None): S.ONE}))
    gu, gv = U.parametric_grad(u, 2), U.parametric_grad(v, 2)
    a = U.inner(gu, gv).a[()]
    assert set(a.terms) == {((1, 0, 0), (1, 0, 0)), ((0, 1, 0), (0, 1, 0))}
    assert a.arity() == 2
    f = U.sin(U.Tensor(U.Scalar.coef(S.xi(0))))
    L = (f * v).a[()]
    assert L.arity() == 1 and list(L.terms) == [(U.ZERO3, None)]
    with pytest.raises(ValueError):
        (u * u)
    res = U.Form([(a, None)]) - U.Form([(L, None)])
    assert U.lhs(res).arity() == 2 and U.rhs(res).arity() == 1
    r = U.rhs(res).scalar().terms[(U.ZERO3, None)]
    assert r is L.terms[(U.ZERO3, None)]


def test_tensor_algebra_det_inv():
    m = U.as_matrix([[2.0, 1.0], [0.5, 3.0]])
    assert abs(U.det(m).a[()].node().args[0] - 5.5) < 1e-15
    mi = U.inv(m)
    ref = np.linalg.inv(np.array([[2.0, 1.0], [0.5, 3.0]]))
    for i in range(2):
        for j in range(2):
            assert abs(mi.a[i, j].node().args[0] - ref[i, j]) < 1e-15
    m3 = np.array([[2.0, 1.0, 0.3], [0.5, 3.0, -1.0], [0.2, 0.1, 1.5]])
    mi3 = U.inv(U.as_matrix(m3.tolist()))
    for i in range(3):
        for j in range(3):
            assert abs(mi3.a[i, j].node().args[0] - np.linalg.inv(m3)[i, j]) < 1e-14


def test_cabi_exports_every_declared_symbol():
    import ctypes
    hdr = open(os.path.join(ROOT, "include", "tigar_b200.h")).read()
    names = set(re.findall(r"\b(tg_[a-z0-9_]+)\s*\(", hdr))
    names -= {"tg_basis", "tg_win"}
    assert len(names) > 30
    lib = ctypes.CDLL(os.path.join(ROOT, "tigar_b200", "libtigar_b200.so"))
    for n in sorted(names):
        assert hasattr(lib, n), "missing export " + n
    from tigar_b200 import _lib
    assert set(_lib.SIGNATURES) == names
    assert lib.tg_version() >= 100


def test_integration_stub_matches_header_and_library():
    """INTEGRATION.md's ctypes structs are generated from the header (tools/gen_ctypes_stub.py);
    executing the documented stub against the built library must give the header's field
    lists and the library's own sizeof (a stale stub passed a short struct)."""
    import ctypes as C
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "gen_ctypes_stub", os.path.join(ROOT, "tools", "gen_ctypes_stub.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    hdr = open(os.path.join(ROOT, "include", "tigar_b200.h")).read()
    structs = gen.parse_structs(hdr)
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = md[md.index(gen.BEGIN) + len(gen.BEGIN):md.index(gen.END)]
    code = re.search(r"```python\n(.*?)```", block, re.S).group(1)
    assert code.strip() == gen.stub(structs).strip(), "run tools/gen_ctypes_stub.py --write"
    lib = C.CDLL(os.path.join(ROOT, "tigar_b200", "libtigar_b200.so"))
    lib.tg_sizeof_win.restype = lib.tg_sizeof_basis.restype = C.c_int64
    ns = {"C": C, "lib": lib}
    exec(code, ns)                                    # includes the sizeof assertions
    from tigar_b200 import _lib
    for name in ("tg_win", "tg_basis"):
        assert [f[0] for f in ns[name]._fields_] == [f[0] for f in structs[name]]
        assert [f[0] for f in getattr(_lib, name)._fields_] == [f[0] for f in structs[name]]
        assert C.sizeof(ns[name]) == C.sizeof(getattr(_lib, name))


def test_product_does_not_import_oracle():
    import subprocess
    import sys
    code = ("import sys; import tigar_b200.api, tigar_b200.engine, tIGAr; "
            "bad=[m for m in sys.modules if m=='oracle' or m.startswith('oracle.')]; "
            "assert not bad, bad")
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tigar_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tigar_b200.engine import TensorPatch
    with pytest.raises(RuntimeError):
        TensorPatch([2, 2], [PB.uniformKnots(2, 0.0, 1.0, 4)] * 2)


def test_generated_qp_kernel_compiles_for_sm100a():
    """The form compiler's CUDA output goes through NVRTC (sm_100a) without a
    GPU; 2-D and 3-D shapes, derivative orders up to 2."""
    from tigar_b200 import jit
    x, y, z = S.xi(0), S.xi(1), S.xi(2)
    u = [S.jet(1, 0, a) for a in [(0, 0, 0), (1, 0, 0), (0, 2, 0), (0, 1, 1)]]
    g = S.jet(2, 0, (0, 0, 1))
    e = S.func("sin", x * y) * u[0] + u[1] * u[2] / (S.func("sqrt", u[3] * u[3] + 1.0)) + g * S.wq() + z
    prog = S.compile_program([e, S.diff(e, 0), S.ZERO], 3)
    fids = sorted(set(j[0] for j in prog.jets))
    jets = [(fids.index(f), c, al) for (f, c, al) in prog.jets]
    src, nth = jit.generate(prog, 3, [4, 4, 4], [4, 4, 4], 4, jets, len(fids))
    assert nth == 64 and jit.check_source(src) > 1000
    prog2 = S.compile_program([S.func("cos", x) * S.jet(7, 0, (1, 1, 0))], 2)
    src2, nth2 = jit.generate(prog2, 2, [3, 5, 1], [3, 3, 1], 3, [(0, 0, (1, 1, 0))], 1)
    assert nth2 == 32 and jit.check_source(src2) > 1000
    with pytest.raises(Exception):
        jit.check_source("this is not CUDA")


def test_gateaux_derivative_matches_finite_difference():
    """derivative(F, u) of a nonlinear scalar form: symbolic d/du against a
    finite difference of the compiled integrand."""
    v = U.Tensor(U.Scalar({(U.ZERO3, None): S.ONE}))
    u0 = S.jet(11, 0, (0, 0, 0))
    ux = S.jet(11, 0, (1, 0, 0))
    g = S.jet(12, 0, (0, 0, 0))
    uT = U.Tensor(U.Scalar.coef(u0))
    uxT = U.Tensor(U.Scalar.coef(ux))
    F = U.Form([(((1.0 + uT * uT) * uxT * v.dx(0) + U.sin(uT) * U.Tensor(U.Scalar.coef(g)) * v).a[()], None)])
    J = U.gateaux(F, 11).scalar()
    keys = set(J.terms)
    assert keys == {((1, 0, 0), (0, 0, 0)), ((1, 0, 0), (1, 0, 0)), ((0, 0, 0), (0, 0, 0))}
    vals = {(11, 0, (0, 0, 0)): 0.3, (11, 0, (1, 0, 0)): -0.8, (12, 0, (0, 0, 0)): 1.7}

    def ev(node, vv):
        prog = S.compile_program([node], 2)
        return run_program(prog, [0.0, 0.0], 1.0, vv)[0]
    # d/du0 of (1+u0^2) ux  = 2 u0 ux ; d/dux = 1+u0^2 ; d/du0 of sin(u0) g = cos(u0) g
    assert abs(ev(J.terms[((1, 0, 0), (0, 0, 0))], vals) - 2 * 0.3 * -0.8) < 1e-14
    assert abs(ev(J.terms[((1, 0, 0), (1, 0, 0))], vals) - (1 + 0.09)) < 1e-14
    assert abs(ev(J.terms[((0, 0, 0), (0, 0, 0))], vals) - math.cos(0.3) * 1.7) < 1e-14
    with pytest.raises(ValueError):
        U.gateaux(U.Form([((U.Tensor(U.Scalar({(None, U.ZERO3): S.ONE})) * v).a[()], None)]), 11)


def test_march_task_packing_covers_every_fibre_once():
    """Warp tasks of tg_ptap_march_w (engine.TensorPatch._march_tasks is pure
    numpy): every fibre (line, column offset) belongs to exactly one piece, no
    task exceeds a warp or MARCH_MAXSUB pieces, pieces wider than a warp are cut
    along the second direction only."""
    import importlib.util
    import os
    import re
    import textwrap
    import numpy as np
    src = open(os.path.join(os.path.dirname(__file__), "..", "tigar_b200", "engine.py")).read()
    i = src.index("    @classmethod\n    def _march_tasks")
    j = src.index("    def ptap_march(self, A, keep=False):")
    ns = {"np": np}
    exec("class X(object):\n    MARCH_MAXSUB = 8\n" + src[i:j], ns)
    X = ns["X"]
    for lena, lenb in [(np.array([4, 4, 4, 7, 4, 4, 7, 4, 4, 4]), np.array([4, 5, 6, 7, 7, 6, 5, 4])),
                       (np.array([5, 5, 5, 9] * 6 + [5]), np.ones(1, np.int64)),
                       (np.array([3, 5, 3, 5, 3]), np.array([9, 5, 5, 5, 9])),
                       (np.array([7] * 5), np.array([7] * 5))]:
        T = X._march_tasks(lena, lenb)
        seen = {}
        for row in T:
            npc = row[0]
            assert 1 <= npc <= 8
            lanes = 0
            for k in range(npc):
                ra, rb, cb0, ncb = row[4 + 4 * k:8 + 4 * k]
                assert ncb >= 1 and cb0 + ncb <= lenb[rb]
                lanes += lena[ra] * ncb
                for cb in range(cb0, cb0 + ncb):
                    assert (ra, rb, cb) not in seen
                    seen[(ra, rb, cb)] = 1
            assert lanes <= 32
        assert len(seen) == len(lena) * int(lenb.sum())


@pytest.mark.parametrize("p,pf,nel,uniform", [(1, 1, 7, True), (2, 2, 9, False), (3, 3, 8, True),
                                              (3, 3, 11, False), (4, 4, 6, False), (2, 3, 7, False)])
def test_march_tables_drive_a_correct_1d_ptap(p, pf, nel, uniform):
    """tigar_b200.march_tables.dir_tables (pure numpy) + a numpy emulation of what one
    lane of k_ptap_march_w does with them (groups of rows sharing first(I), zero-padded
    coefficient vectors addressed through the 2-bit shifts, sliding (p+1)x(2p+1)
    accumulator block, emit through jrec) reproduce the dense M^T A M of a banded
    1-D A on its window -- for every segment split."""
    import numpy as np
    from oracle import bsplines as OB
    from tigar_b200 import march_tables
    from tigar_b200.engine import Window
    rng = np.random.RandomState(11)
    kn = np.array(OB.uniform_knots(p, 0.0, 1.0, nel))
    if not uniform:
        inner = kn[p + 1:-(p + 1)]
        kn[p + 1:-(p + 1)] = inner + 0.35 / nel * (rng.rand(len(inner)) - 0.5)
    s1 = OB.Spline1(p, list(kn))
    uk_ = np.asarray(s1.uniqueKnots)
    x = np.concatenate([uk_[e] + (uk_[e + 1] - uk_[e]) * np.arange(pf) / pf for e in range(nel)]
                       + [uk_[-1:]])
    nfe, ncp = len(x), s1.ncp
    span = np.array([s1.getKnotSpan(u) for u in x])
    vals = np.array([s1.basisFuncs(sp, u) for sp, u in zip(span, x)])
    first = span - p
    keep = np.abs(vals) > 1e-15
    m_lo = first + keep.argmax(axis=1)
    m_hi = first + p - keep[:, ::-1].argmax(axis=1)
    M = np.zeros((nfe, ncp))
    for I in range(nfe):
        for k in range(p + 1):
            if keep[I, k]:
                M[I, first[I] + k] = vals[I, k]
    g = np.arange(nfe)
    loA = np.maximum((g - 1) // pf, 0) * pf
    hiA = (np.minimum(g // pf, nel - 1) + 1) * pf
    wM = Window([nfe], [ncp], [m_lo], [m_hi])
    wA = Window([nfe], [nfe], [loA], [hiA])
    wMT = wM.transpose()
    wC = wMT.compose(wA.compose(wM))
    loC, hiC = wC.lo[0].astype(np.int64), wC.hi[0].astype(np.int64)
    slo, shi = wMT.lo[0], wMT.hi[0]
    T = march_tables.dir_tables(p, pf, first, vals, m_lo, m_hi, loA, hiA, loC, hiC, 4)
    assert T is not None
    irec, jrec, cpad, grp, gidx = T["irec"], T["jrec"], T["cpad"], T["grp"], T["gidx"]
    assert T["GMAX"] >= 2 * p + 1 and np.all(np.diff(grp) <= min(4, pf))
    A = np.zeros((nfe, nfe))
    for I in range(nfe):
        A[I, loA[I]:hiA[I] + 1] = rng.randn(hiA[I] - loA[I] + 1)
    ref = M.T @ A @ M
    inwin = np.zeros_like(ref, dtype=bool)
    for i in range(ncp):
        inwin[i, loC[i]:hiC[i] + 1] = True
    assert np.abs(ref[~inwin]).max() < 1e-14               # the window holds the pattern
    CW, TW = 2 * p + 1, p + 2
    for nseg in (1, 3):
        seg = [(ncp * k) // nseg for k in range(nseg + 1)]
        C = np.full((ncp, ncp), np.nan)
        for sgi in range(nseg):
            i_lo, i_hi = seg[sgi], seg[sgi + 1]
            g0, g1 = gidx[slo[i_lo]], gidx[shi[i_hi - 1]]
            acc = np.zeros((p + 1, CW))
            ib = first[grp[g0]]

            def emit_shift():
                nonlocal acc, ib
                if i_lo <= ib < i_hi:
                    clo, lenC = jrec[ib, 0], jrec[ib, 1]
                    for c in range(CW):
                        jj = c - clo
                        if 0 <= jj < lenC:
                            C[ib, loC[ib] + jj] = acc[0, c]
                acc[:-1] = acc[1:]
                acc[-1] = 0.0
                ib += 1
            for gk in range(g0, g1 + 1):
                n0, n1 = grp[gk], grp[gk + 1]
                while ib < irec[n0, 1]:
                    emit_shift()
                for I in range(n0, n1):
                    lenI, lo, sb = irec[I, 0] & 255, irec[I, 0] >> 8, irec[I, 2]
                    assert irec[I, 1] == irec[n0, 1]              # one first(I) per group
                    tv = np.zeros(TW)
                    for q in range(lenI):
                        e = (sb >> (2 * q)) & 3
                        tv += A[I, lo + q] * cpad[lo + q, 2 - e:2 - e + TW]
                    for k in range(p + 1):
                        for m in range(TW):
                            c = m - k + p
                            if 0 <= c < CW:
                                acc[k, c] += cpad[I, 1 + k] * tv[m]
            for k in range(p + 1):
                emit_shift()
        assert not np.isnan(C[inwin]).any()                   # every window entry written once
        assert np.abs(C[inwin] - ref[inwin]).max() < 1e-13 * max(1.0, np.abs(ref).max())


def _eval_coef(t, X):
    """Value of a coefficient-only ufl_lite scalar at parametric point X."""
    sc = U.as_tensor(t).a[()]
    assert sc.is_coef()
    prog = S.compile_program([sc.node()], len(X))
    return run_program(prog, list(X), 1.0, {})[0]


def test_curl_and_parametric_expression_symbolics():
    """curl_from_grad (cartesianCurl, calculusUtils.py:278-302) and
    expression_from_string (parametricExpression, common.py:1111-1117) on the
    identity geometry, against closed forms."""
    X3 = (0.3, 0.7, 0.2)
    x = [U.Tensor(U.Scalar.coef(S.xi(d))) for d in range(3)]
    f = U.as_vector([x[1] * x[2] * x[2], U.sin(x[0]) * x[2], x[0] * x[0] * x[1]])
    c = U.curl_from_grad(f, U.parametric_grad(f, 3))
    ref = [X3[0] ** 2 - math.sin(X3[0]), 2 * X3[1] * X3[2] - 2 * X3[0] * X3[1],
           math.cos(X3[0]) * X3[2] - X3[2] ** 2]
    for i in range(3):
        assert abs(_eval_coef(c[i], X3) - ref[i]) < 1e-14
    X2 = (0.4, 0.9)
    y = x[:2]
    g = U.as_vector([y[0] * y[1] * y[1], U.exp(y[0]) * y[1]])
    c2 = U.curl_from_grad(g, U.parametric_grad(g, 2))
    assert abs(_eval_coef(c2, X2) - (math.exp(X2[0]) * X2[1] - 2 * X2[0] * X2[1])) < 1e-14
    s = y[0] * y[0] * U.cos(y[1])
    c3 = U.curl_from_grad(s, U.parametric_grad(s, 2))
    assert abs(_eval_coef(c3[0], X2) - X2[0] ** 2 * math.sin(X2[1])) < 1e-14
    assert abs(_eval_coef(c3[1], X2) - 2 * X2[0] * math.cos(X2[1])) < 1e-14
    with pytest.raises(ValueError):
        U.curl_from_grad(U.as_matrix([[1.0, 0.0], [0.0, 1.0]]), U.as_matrix([[1.0, 0.0], [0.0, 1.0]]))
    e = U.expression_from_string("sin(pi*x[0])*pow(x[1],2) + sqrt(x[0]+1.0)/exp(x[1])", y)
    assert abs(_eval_coef(e, X2) - (math.sin(math.pi * X2[0]) * X2[1] ** 2
                                    + math.sqrt(X2[0] + 1.0) / math.exp(X2[1]))) < 1e-14
    v = U.expression_from_string(("x[0]", "2.0*x[1]"), y)
    assert abs(_eval_coef(v[1], X2) - 2 * X2[1]) < 1e-15


def test_calculus_utils_on_a_polar_map():
    """tigar_b200.calculus (the reference's calculusUtils.py:18-24, 56-69, 255-302,
    412-470) on the polar map F(r, t) = (r cos t, r sin t): metric diag(1, r^2), volume
    element r, Cartesian gradient / divergence / curl of fields given in the
    parametric coordinates, and the Gauss rules against the reference's constants."""
    from tigar_b200 import calculus as CU
    import tIGAr.calculusUtils as TCU
    assert TCU.cartesianGrad is CU.cartesianGrad
    old = U.DEFAULT_DIM[0]
    U.DEFAULT_DIM[0] = 2
    try:
        r, t = (U.Tensor(U.Scalar.coef(S.xi(d))) for d in range(2))
        F = U.as_vector([r * U.cos(t), r * U.sin(t)])
        X = (1.7, 0.6)
        x_, y_ = X[0] * math.cos(X[1]), X[0] * math.sin(X[1])
        g = CU.getMetric(F)
        assert abs(_eval_coef(g[0, 0], X) - 1.0) < 1e-14 and abs(_eval_coef(g[1, 1], X) - X[0] ** 2) < 1e-13
        assert abs(_eval_coef(g[0, 1], X)) < 1e-14
        assert abs(_eval_coef(CU.volumeJacobian(g), X) - X[0]) < 1e-14
        P = CU.pinvD(F)
        DF = U.grad(F)
        I2 = U.dot(P, DF)                                   # pinv(DF) DF = I for a square map
        for i in range(2):
            for j in range(2):
                assert abs(_eval_coef(I2[i, j], X) - (1.0 if i == j else 0.0)) < 1e-13
        f = r * r                                            # x^2 + y^2
        gf = CU.cartesianGrad(f, F)
        assert abs(_eval_coef(gf[0], X) - 2 * x_) < 1e-13 and abs(_eval_coef(gf[1], X) - 2 * y_) < 1e-13
        v = U.as_vector([r * U.cos(t) * r * U.sin(t), r * r])   # (x y, x^2 + y^2)
        assert abs(_eval_coef(CU.cartesianDiv(v, F), X) - (y_ + 2 * y_)) < 1e-13
        assert abs(_eval_coef(CU.cartesianCurl(v, F), X) - (2 * x_ - x_)) < 1e-13
        lap = CU.cartesianDiv(CU.cartesianGrad(f, F), F)
        assert abs(_eval_coef(lap, X) - 4.0) < 1e-12
    finally:
        U.DEFAULT_DIM[0] = old
    xs, ws = CU.getQuadRule(3)
    vals = [c.a[()].node().args[0] for c in xs], [c.a[()].node().args[0] for c in ws]
    assert vals[0][1] == 0.0 and abs(vals[0][2] - 0.77459666924148337703585308) < 1e-15
    assert abs(vals[1][0] - 0.55555555555555555555555556) < 1e-15
    x4, w4 = CU.getQuadRuleInterval(4, 0.2)
    assert abs(x4[3].a[()].node().args[0] - 0.1 * 0.86113631159405257524) < 1e-16
    assert abs(sum(c.a[()].node().args[0] for c in w4) - 0.2) < 1e-15


def test_window_algebra_matches_sparse_patterns():
    """engine.Window (pure numpy part): nnz, transpose and compose of tensor-product row
    windows against explicit scipy patterns, and the closed-form row pointer
    rowptr(r) = S0[r0] l1 l2 + T0 (S1[r1] l2 + T1 S2[r2]) the kernels use."""
    import scipy.sparse as sp
    from tigar_b200.engine import Window
    rng = np.random.RandomState(3)

    def rand_win(nr, nc):
        lo, hi = [], []
        for a, b in zip(nr, nc):
            l = np.sort(rng.randint(0, b, size=a))
            h = np.minimum(b - 1, l + rng.randint(0, 3, size=a))
            h = np.maximum.accumulate(h)
            lo.append(l)
            hi.append(h)
        return Window(nr, nc, lo, hi)

    def pattern(w):
        mats = []
        for d in range(w.dim):
            m = np.zeros((w.nr[d], w.nc[d]))
            for r in range(w.nr[d]):
                m[r, w.lo[d][r]:w.hi[d][r] + 1] = 1.0
            mats.append(sp.csr_matrix(m))
        out = mats[0]
        for m in mats[1:]:
            out = sp.kron(m, out, format="csr")          # first direction fastest
        return out

    for nr, nm, nc in [([5, 4], [6, 3], [4, 5]), ([3, 4, 2], [4, 3, 3], [2, 5, 3]), ([7], [5], [6])]:
        A, B = rand_win(nr, nm), rand_win(nm, nc)
        PA, PB = pattern(A), pattern(B)
        assert A.nnz == PA.nnz and A.nrows == PA.shape[0] and A.ncols == PA.shape[1]
        # closed-form row pointer
        S = [np.concatenate([[0], np.cumsum(l)]) for l in A.len] + [np.array([0, 1])] * (3 - A.dim)
        ln = list(A.len) + [np.array([1])] * (3 - A.dim)
        n3 = list(A.nr) + [1] * (3 - A.dim)
        T0, T1 = S[0][-1], S[1][-1]
        row = 0
        for r2 in range(n3[2]):
            for r1 in range(n3[1]):
                for r0 in range(n3[0]):
                    rp = S[0][r0] * ln[1][r1] * ln[2][r2] + T0 * (S[1][r1] * ln[2][r2] + T1 * S[2][r2])
                    assert rp == PA.indptr[row]
                    row += 1
        # transpose: smallest window containing the transposed pattern
        AT = A.transpose() if all((PA.sum(axis=0) > 0).A1) else None
        if AT is not None:
            PT = pattern(AT)
            assert (PT.multiply(PA.T) - PA.T).nnz == 0        # contains A^T
            for d in range(A.dim):
                for c in range(A.nc[d]):
                    rows = [r for r in range(A.nr[d]) if A.lo[d][r] <= c <= A.hi[d][r]]
                    assert AT.lo[d][c] == min(rows) and AT.hi[d][c] == max(rows)
        # compose: window of the product pattern, tight per direction
        AB = A.compose(B)
        PAB = ((PA @ PB) > 0).astype(float)
        PW = pattern(AB)
        assert (PW.multiply(PAB) - PAB).nnz == 0              # contains the product
        for d in range(A.dim):
            for r in range(A.nr[d]):
                cols = np.concatenate([np.arange(B.lo[d][k], B.hi[d][k] + 1)
                                       for k in range(A.lo[d][r], A.hi[d][r] + 1)])
                assert AB.lo[d][r] == cols.min() and AB.hi[d][r] == cols.max()


def _eval_with_functions(expr, values):
    """Value of a coefficient-only expression given {Function.fid: number}."""
    sc = U.as_tensor(expr).a[()]
    prog = S.compile_program([sc.node()], 2)
    jv = {j: values[j[0]] for j in prog.jets}
    return run_program(prog, [0.0, 0.0], 1.0, jv)[0]


class _DummyOwner(object):
    pass


def test_linear_combination_assign_and_time_integrators():
    """api.linear_combination / Function.assign of linear combinations and the
    integrators of tigar_b200.time_integration (timeIntegration.py:13-247) against
    the Newmark / generalized-alpha update formulas solved independently here."""
    import torch
    from tigar_b200 import api as A
    from tigar_b200 import time_integration as TI
    import tIGAr.timeIntegration as TTI
    assert TTI.GeneralizedAlphaIntegrator is TI.GeneralizedAlphaIntegrator
    V = A.FunctionSpace(_DummyOwner())
    u, v, w = A.Function(V), A.Function(V), A.Function(V)
    e = 2.0 * u - U.Constant(0.5) * v + (u + v) * 3.0
    assert sorted(A.linear_combination(e)) == sorted([(u.fid, 5.0), (v.fid, 2.5)])
    for bad in (u * u, u + 1.0, u.dx(0), U.sin(u)):
        with pytest.raises(ValueError):
            A.linear_combination(bad)
    rng = np.random.RandomState(0)
    for f in (u, v, w):
        f.set_iga(torch.from_numpy(rng.rand(7)))
    t = A.Function(V)
    t.assign(e)
    assert torch.allclose(t.iga, 5.0 * u.iga + 2.5 * v.iga, rtol=0, atol=1e-15)

    dt, rho = 0.1, 0.4
    x, x0, v0, a0 = (A.Function(V) for _ in range(4))
    vals = {x.fid: 1.3, x0.fid: 0.9, v0.fid: -0.7, a0.fid: 2.1}
    F_ = {f.fid: f for f in (x, x0, v0, a0)}
    # ---- second order: Newmark relations solved for a1, v1 given x1
    gi = TI.GeneralizedAlphaIntegrator(rho, dt, x, [x0, v0, a0], t=1.0)
    am, af = (2 - rho) / (1 + rho), 1 / (1 + rho)
    g, b = 0.5 + am - af, 0.25 * (1 + am - af) ** 2
    assert abs(gi.ALPHA_M - am) < 1e-15 and abs(gi.GAMMA - g) < 1e-15 and abs(gi.BETA - b) < 1e-15
    a1 = (vals[x.fid] - vals[x0.fid] - dt * vals[v0.fid] - 0.5 * dt * dt * (1 - 2 * b) * vals[a0.fid]) / (b * dt * dt)
    v1 = vals[v0.fid] + dt * ((1 - g) * vals[a0.fid] + g * a1)
    assert abs(_eval_with_functions(gi.xddot(), vals) - a1) < 1e-11
    assert abs(_eval_with_functions(gi.xdot(), vals) - v1) < 1e-12
    assert abs(_eval_with_functions(gi.x_alpha(), vals) - (af * vals[x.fid] + (1 - af) * vals[x0.fid])) < 1e-14
    assert abs(_eval_with_functions(gi.xdot_alpha(), vals) - (af * v1 + (1 - af) * vals[v0.fid])) < 1e-12
    assert abs(_eval_with_functions(gi.xddot_alpha(), vals) - (am * a1 + (1 - am) * vals[a0.fid])) < 1e-11
    # same-velocity predictor: x1 such that v1 == v0
    xp = _eval_with_functions(gi.sameVelocityPredictor(), vals)
    ap = (xp - vals[x0.fid] - dt * vals[v0.fid] - 0.5 * dt * dt * (1 - 2 * b) * vals[a0.fid]) / (b * dt * dt)
    assert abs(vals[v0.fid] + dt * ((1 - g) * vals[a0.fid] + g * ap) - vals[v0.fid]) < 1e-12
    # advance(): data moves with the OLD values on the right-hand sides
    for fid, val in vals.items():
        F_[fid].set_iga(torch.full((3,), val, dtype=torch.float64))
    gi.advance()
    assert np.allclose(x0.iga.numpy(), 1.3) and np.allclose(v0.iga.numpy(), v1, atol=1e-12)
    assert np.allclose(a0.iga.numpy(), a1, atol=1e-10)
    assert abs(gi.t - 1.2) < 1e-15
    # ---- first order
    y, y0, yd0 = (A.Function(V) for _ in range(3))
    vv = {y.fid: 0.3, y0.fid: 0.5, yd0.fid: 1.1}
    g1 = TI.GeneralizedAlphaIntegrator(rho, dt, y, [y0, yd0])
    am1 = 0.5 * (3 - rho) / (1 + rho)
    gm = 0.5 + am1 - af
    yd1 = ((vv[y.fid] - vv[y0.fid]) / dt - (1 - gm) * vv[yd0.fid]) / gm
    assert abs(_eval_with_functions(g1.xdot(), vv) - yd1) < 1e-12
    assert abs(_eval_with_functions(g1.xdot_alpha(), vv) - (am1 * yd1 + (1 - am1) * vv[yd0.fid])) < 1e-12
    assert g1.sameVelocityPredictor() is y0
    with pytest.raises(ValueError):
        g1.xddot()
    # ---- backward Euler, second order
    be = TI.BackwardEulerIntegrator(dt, x, [x0, v0])
    for fid, val in vals.items():
        F_[fid].set_iga(torch.full((2,), val, dtype=torch.float64))
    bv = (vals[x.fid] - vals[x0.fid]) / dt
    assert abs(_eval_with_functions(be.xdot(), vals) - bv) < 1e-12
    assert abs(_eval_with_functions(be.xddot(), vals) - (bv - vals[v0.fid]) / dt) < 1e-11
    be.advance()
    assert np.allclose(x0.iga.numpy(), vals[x.fid]) and np.allclose(v0.iga.numpy(), bv, atol=1e-12)
    ls = TI.LoadStepper(0.25)
    ls.advance()
    assert abs(float(ls.t) - 0.5) < 1e-15
    with pytest.raises(NotImplementedError):
        TI.LinearDGSpaceTimeIntegrator(dt, x, x0)


def test_nurbs_shim_refine_elevate_preserve_geometry_and_annulus_is_exact():
    """tigar_b200.nurbs (the igakit stand-in of SURVEY 8f n2): knot insertion and
    Bezier degree elevation leave the rational map unchanged, the quarter annulus
    of BASELINE configs[3] is an exact circular arc with a linear radius, and the
    NURBSControlMesh net is flattened first-direction-fastest (NURBS.py:43-77)."""
    from oracle import bsplines as OB
    from tigar_b200.nurbs import NURBS, NURBSControlMesh, quarter_annulus

    def evaluate(nrb, pts):
        sp = [OB.Spline1(int(p), list(k)) for p, k in zip(nrb.degree, nrb.knots)]
        out = []
        for xi in pts:
            acc = np.zeros(nrb.control.shape[-1])
            spans = [s.getKnotSpan(u) for s, u in zip(sp, xi)]
            vals = [s.basisFuncs(k, u) for s, k, u in zip(sp, spans, xi)]
            idx = [range(k - s.p, k + 1) for s, k in zip(sp, spans)]
            for loc in np.ndindex(*[len(r) for r in idx]):
                w = np.prod([vals[d][loc[d]] for d in range(len(sp))])
                acc += w * nrb.control[tuple(idx[d][loc[d]] for d in range(len(sp)))]
            out.append(acc[:-1] / acc[-1])
        return np.array(out)

    rng = np.random.RandomState(5)
    s = 1.0 / math.sqrt(2.0)
    arc = np.array([[1.0, 0.0, 1.0], [s, s, s], [0.0, 1.0, 1.0]])
    net = np.zeros((2, 3, 3))                               # (radial, angular) quadratic arc
    for i, r in enumerate((1.0, 2.0)):
        net[i, :, 0], net[i, :, 1], net[i, :, 2] = r * arc[:, 0], r * arc[:, 1], arc[:, 2]
    base = NURBS([[0, 0, 1, 1], [0, 0, 0, 1, 1, 1]], net, homogeneous=True)
    pts = rng.rand(12, 2) * 0.98 + 0.01
    ref = evaluate(base, pts)
    assert np.allclose(np.hypot(ref[:, 0], ref[:, 1]), 1.0 + pts[:, 0], atol=1e-14)
    mod = NURBS([k.copy() for k in base.knots], base.control.copy(), homogeneous=True)
    mod.elevate(0, 2).elevate(1, 1)
    assert mod.degree == [3, 3]
    assert np.allclose(evaluate(mod, pts), ref, atol=1e-14)
    mod.refine(0, [0.25, 0.5, 0.5001]).refine(1, [0.3, 0.9])
    assert mod.control.shape[:2] == (7, 6)
    assert np.allclose(evaluate(mod, pts), ref, atol=1e-13)

    ann = quarter_annulus(3, [5, 4], 2)
    assert ann.degree == [3, 3] and ann.control.shape == (8, 7, 3)
    X = evaluate(ann, pts)
    assert np.allclose(np.hypot(X[:, 0], X[:, 1]), 1.0 + pts[:, 0], atol=1e-13)   # r linear in xi_0
    assert np.all(X >= -1e-14)                                                     # first quadrant
    ann3 = quarter_annulus(3, [3, 4, 2], 3, height=0.5)
    p3 = rng.rand(6, 3) * 0.98 + 0.01
    X3 = evaluate(ann3, p3)
    assert np.allclose(np.hypot(X3[:, 0], X3[:, 1]), 1.0 + p3[:, 0], atol=1e-13)
    assert np.allclose(X3[:, 2], 0.5 * p3[:, 2], atol=1e-14)
    cm = NURBSControlMesh(ann3)
    n0, n1, n2 = ann3.control.shape[:3]
    assert cm.getNsd() == 3 and cm.controlNet().shape == (n0 * n1 * n2, 4)
    i, j, k = 2, 3, 1
    assert np.array_equal(cm.controlNet()[i + n0 * (j + n1 * k)], ann3.control[i, j, k])
    assert cm.getHomogeneousCoordinate(i + n0 * (j + n1 * k), 3) == ann3.control[i, j, k, 3]


def test_mutable_parameter_is_baked_at_compile_time():
    """ufl_lite.Parameter / symbolic.param: a leaf with zero derivative whose current
    value enters the program each time it is compiled (dolfin Constant.assign /
    Expression parameters between solves)."""
    from tigar_b200 import api as A
    x = U.Tensor(U.Scalar.coef(S.xi(0)))
    t = U.Parameter(0.25)
    e = U.sin(t * x) + t * t
    node = U.as_tensor(e).a[()].node()
    dnode = S.diff(node, 0)
    X = 0.7
    for val in (0.25, 1.5):
        t.assign(val)
        prog = S.compile_program([node, dnode], 1)
        got = run_program(prog, [X], 1.0, {})
        assert abs(got[0] - (math.sin(val * X) + val * val)) < 1e-15
        assert abs(got[1] - val * math.cos(val * X)) < 1e-15
        assert val in prog.consts
    assert float(t) == 1.5 and S.jets_of([node]) == []
    frozen = S.freeze_params(S.mul(S.param(t.pid), S.const(2.0)))
    assert frozen.is_const() and frozen.args[0] == 3.0
    # a linear combination with a parameter coefficient uses the current value
    V = A.FunctionSpace(_DummyOwner())
    u = A.Function(V)
    assert A.linear_combination(t * u) == [(u.fid, 1.5)]


def test_curvilinear_calculus_in_polar_coordinates():
    """tigar_b200.calculus curvilinear part (calculusUtils.py:26-54, 71-250, 307-346) in
    the polar chart g = diag(1, r^2): Christoffel symbols, metric compatibility,
    gradient / divergence / Laplace-Beltrami against the textbook formulas, index
    gymnastics, mapped normal, surface element and the three pushforwards."""
    from tigar_b200 import calculus as CU
    old = U.DEFAULT_DIM[0]
    U.DEFAULT_DIM[0] = 2
    try:
        r, t = (U.Tensor(U.Scalar.coef(S.xi(d))) for d in range(2))
        F = U.as_vector([r * U.cos(t), r * U.sin(t)])
        g = CU.getMetric(F)
        X = (1.3, 0.8)
        R, T_ = X
        ev = lambda e: _eval_coef(e, X)
        gam = CU.getChristoffel(g)
        ref = np.zeros((2, 2, 2))
        ref[0, 1, 1] = -R
        ref[1, 0, 1] = ref[1, 1, 0] = 1.0 / R
        for idx in np.ndindex(2, 2, 2):
            assert abs(ev(gam[idx]) - ref[idx]) < 1e-12, idx
        # metric compatibility: the covariant derivative of g vanishes
        Dg = CU.covariantDerivative(CU.CurvilinearTensor(g, g))
        assert Dg.lowered == [True, True, True]
        for idx in np.ndindex(2, 2, 2):
            assert abs(ev(Dg.T[idx])) < 1e-12
        # scalar: GRAD raises the derivative index, DIV GRAD = Laplace-Beltrami
        f = r * r * U.cos(2.0 * t)                   # x^2 - y^2: harmonic
        G = CU.curvilinearGrad(CU.CurvilinearTensor(f, g))
        assert G.lowered == [False]
        assert abs(ev(G.T[0]) - 2 * R * math.cos(2 * T_)) < 1e-12
        assert abs(ev(G.T[1]) - (-2 * R * R * math.sin(2 * T_)) / R ** 2) < 1e-12
        assert abs(ev(CU.curvilinearDiv(G).T)) < 1e-11
        lap = CU.curvilinearDiv(CU.curvilinearGrad(CU.CurvilinearTensor(r * r * r, g)))
        assert abs(ev(lap.T) - 9 * R) < 1e-11         # (1/r)(r f')' with f = r^3
        # contravariant vector: div v = d_r v^r + d_t v^t + v^r / r
        v = U.as_vector([r * r * U.sin(t), U.cos(t) / r])
        dv = CU.curvilinearDiv(CU.CurvilinearTensor(v, g, [False]))
        assert dv.rank() == 0
        assert abs(ev(dv.T) - (2 * R * math.sin(T_) - math.sin(T_) / R + R * math.sin(T_))) < 1e-12
        # index gymnastics and the inner product
        a = CU.CurvilinearTensor(U.as_vector([r * t, r + t]), g)            # covector
        up = a.sharp()
        assert up.lowered == [False] and abs(ev(up.T[1]) - (R + T_) / R ** 2) < 1e-13
        assert abs(ev(up.flat().T[1]) - (R + T_)) < 1e-13
        assert abs(ev(CU.curvilinearInner(a, a)) - ((R * T_) ** 2 + (R + T_) ** 2 / R ** 2)) < 1e-12
        s2 = 2.0 * a - a
        assert abs(ev(s2.T[0]) - R * T_) < 1e-14
        # geometry of the circle r = const and the pushforwards
        n = CU.mappedNormal(U.as_vector([1.0, 0.0]), F)
        assert abs(ev(n[0]) - math.cos(T_)) < 1e-13 and abs(ev(n[1]) - math.sin(T_)) < 1e-13
        assert abs(ev(CU.surfaceJacobian(g, U.as_vector([1.0, 0.0]))) - R) < 1e-13
        w = U.as_vector([r, t])
        rt = CU.cartesianPushforwardRT(w, F)         # DF w / J
        DFn = np.array([[math.cos(T_), -R * math.sin(T_)], [math.sin(T_), R * math.cos(T_)]])
        exp_rt = DFn @ np.array([R, T_]) / R
        exp_n = np.linalg.inv(DFn.T) @ np.array([R, T_])
        pn = CU.cartesianPushforwardN(w, F)
        for i in range(2):
            assert abs(ev(rt[i]) - exp_rt[i]) < 1e-13 and abs(ev(pn[i]) - exp_n[i]) < 1e-13
        assert abs(ev(CU.cartesianPushforwardW(r * t, F)) - T_) < 1e-13
    finally:
        U.DEFAULT_DIM[0] = old


def test_dof_list_is_a_list_that_remembers_its_arrays():
    """getSideDofs returns a real Python list (reference API, BSplines.py:599-649) whose
    numpy form is kept through ``+=`` / ``+`` and dropped by any other edit."""
    from tigar_b200.bsplines import DofList, BSpline, uniformKnots
    sp = BSpline([2, 3], [uniformKnots(2, 0.0, 1.0, 4), uniformKnots(3, 0.0, 1.0, 3)])
    a = sp.getSideDofs(0, 0)
    b = sp.getSideDofs(1, 1, nLayers=2)
    assert isinstance(a, list) and isinstance(a, DofList) and a == [0, 6, 12, 18, 24, 30]
    assert b == list(range(30, 36)) + list(range(24, 30))
    z = DofList()
    z += a
    z += b
    z += [99, 98]                                     # plain lists are fine
    assert z == a + b + [99, 98] and len(z._chunks) == 3
    assert z.asarray().tolist() == list(z) and z.asarray().dtype == np.int64
    c = a + b
    assert isinstance(c, DofList) and c.asarray().tolist() == list(a) + list(b)
    assert a.asarray().tolist() == list(a)            # operands untouched
    z.append(7)
    assert z._chunks is None and z.asarray().tolist() == list(z)
    z2 = DofList([5, 4])
    z2 += a
    assert z2.asarray().tolist() == [5, 4] + list(a)
    z2[0] = 1
    z2 += b
    assert z2.asarray().tolist() == list(z2)
    # generator bookkeeping: field offsets, duplicates at corners kept
    from tigar_b200 import api as A
    from tigar_b200.bsplines import ExplicitBSplineControlMesh
    cm = ExplicitBSplineControlMesh([2, 3], [uniformKnots(2, 0.0, 1.0, 4), uniformKnots(3, 0.0, 1.0, 3)])
    gen = A.EqualOrderSpline(2, cm)
    gen.addZeroDofs(0, a)
    gen.addZeroDofs(1, a)
    gen.addZeroDofs(1, [3, 4])
    assert list(gen.zeroDofs) == list(a) + [d + 36 for d in a] + [39, 40]
    assert gen.zeroDofs.asarray().tolist() == list(gen.zeroDofs)


def test_conditional_is_a_true_select():
    """UFL's conditional selects; an inf/NaN in the unselected branch (the usual
    use guards a singularity: conditional(gt(r, eps), 1/r, 0)) must not reach the result, and
    the derivative of a selected branch is the selected derivative."""
    x = S.xi(0)
    c = S.binary("gt", x, S.const(0.5))
    e = S.select(c, S.div(S.ONE, S.sub(x, S.const(0.25))), S.const(7.0))     # singular at 0.25
    prog = S.compile_program([e, S.diff(e, 0)], 1)
    v = run_program(prog, [0.25], 1.0, {})
    assert v[0] == 7.0 and v[1] == 0.0                       # no NaN from 0 * inf
    v = run_program(prog, [0.75], 1.0, {})
    assert abs(v[0] - 2.0) < 1e-15 and abs(v[1] + 4.0) < 1e-14
    # the form language uses it term by term
    r = U.Tensor(U.Scalar.coef(x))
    t = U.conditional(U.gt(r, 0.5), 1.0 / (r - 0.25), 0.0)
    p2 = S.compile_program([t.a[()].node() if hasattr(t.a[()], "node") else t.a.item().node()], 1)
    assert run_program(p2, [0.25], 1.0, {})[0] == 0.0
    # generated CUDA carries the ternary
    from tigar_b200 import jit
    src, _ = jit.generate(prog, 1, [3, 1, 1], [3, 1, 1], 1, [], 0)
    assert "!= 0.0) ?" in src


def test_fd_weight_model_selection():
    """solvers.choose_fd_weights: stiffness-only / mass-only / both are recovered exactly from
    the relative-error normal equations, and a stiffness operator seen through a smoothly
    varying coefficient (curved geometry) keeps sigma = 0 instead of the spurious mass term the
    full model would fit."""
    torch = pytest.importorskip("torch")
    try:
        from tigar_b200.solvers import choose_fd_weights
    except (ImportError, OSError) as e:
        pytest.skip("library not built: %s" % e)
    rng = np.random.default_rng(0)
    n = 4000
    cols = rng.uniform(0.5, 2.0, (n, 4))

    def sums(d):
        u = cols / d[:, None]
        G, r = u.T @ u, u.sum(0)
        return np.array([G[a, b] for a in range(4) for b in range(a, 4)] + list(r) + [n])
    c, s = choose_fd_weights(sums(cols[:, :3] @ np.array([2.0, 0.3, 1.5])), 3)
    assert np.allclose(c, [2.0, 0.3, 1.5], rtol=1e-10) and s == 0.0
    c, s = choose_fd_weights(sums(0.7 * cols[:, 3]), 3)
    assert c == [0.0, 0.0, 0.0] and abs(s - 0.7) < 1e-10
    c, s = choose_fd_weights(sums(cols @ np.array([2.0, 0.3, 1.5, 5.0])), 3)
    assert np.allclose(c + [s], [2.0, 0.3, 1.5, 5.0], rtol=1e-10)
    d = (cols[:, :3] @ np.array([2.0, 0.3, 1.5])) * rng.uniform(0.8, 1.25, n)
    c, s = choose_fd_weights(sums(d), 3)
    assert s == 0.0 and np.allclose(c, [2.0, 0.3, 1.5], rtol=0.05)
