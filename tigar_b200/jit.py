"""
Form compiler back end: turns a compiled Gauss-point program
(tigar_b200.symbolic.Program) plus the element shape into CUDA C and has the
library compile it with NVRTC for sm_100a (``tg_jit_compile``).  This is the
role FFC/dijitso play behind ``dolfin.assemble`` in the reference
(common.py:1215-1216): one generated kernel per integrand, cached.

Generated kernel: one CTA per cell, one thread per Gauss point.  Jets of the
coefficient functions are sum-factorised through shared memory with every loop
bound a compile-time constant; the scalar program is emitted as straight-line
FP64 code (no interpreter, no local-memory register file).
"""
import ctypes as C
import hashlib
import os

from . import dev
from . import symbolic as S
from ._lib import lib, check, c_vp, c_i32, c_i64

_UN = {"neg": "-(%s)", "sin": "sin(%s)", "cos": "cos(%s)", "exp": "exp(%s)", "log": "log(%s)",
       "sqrt": "sqrt(%s)", "abs": "fabs(%s)", "tan": "tan(%s)", "tanh": "tanh(%s)",
       "sinh": "sinh(%s)", "cosh": "cosh(%s)", "atan": "atan(%s)", "mov": "(%s)"}
_BIN = {"add": "(%s + %s)", "sub": "(%s - %s)", "mul": "(%s * %s)", "div": "(%s / %s)",
        "pow": "pow(%s, %s)", "max": "fmax(%s, %s)", "min": "fmin(%s, %s)",
        "gt": "((%s > %s) ? 1.0 : 0.0)", "selz": "((%s != 0.0) ? %s : 0.0)"}
_NAMES = {v: k for k, v in S.OPCODES.items()}

MAXFUN = 16


class QpArgs(C.Structure):
    _fields_ = [("tab", c_vp * 3), ("idx", c_vp * 3), ("wq", c_vp * 3), ("xq", c_vp * 3),
                ("coef", c_vp * MAXFUN), ("n", c_i32 * 3), ("nel", c_i32 * 3),
                ("cell0", c_i64), ("out", c_vp), ("sstride", c_i64), ("lc", c_i32),
                ("elast0", c_i32)]


class OpArgs(C.Structure):
    """QpArgs + the colour lattice of the launch and the output vector (fused
    matrix-free operator kernel, ``generate(..., op=...)``)."""
    _fields_ = [("tab", c_vp * 3), ("idx", c_vp * 3), ("wq", c_vp * 3), ("xq", c_vp * 3),
                ("coef", c_vp * MAXFUN), ("n", c_i32 * 3), ("nel", c_i32 * 3),
                ("cell0", c_i64), ("out", c_vp), ("sstride", c_i64), ("lc", c_i32),
                ("elast0", c_i32), ("co", c_i32 * 3), ("cs", c_i32 * 3),
                ("cn", c_i32 * 3), ("y", c_vp)]


_PRELUDE_OP = r"""
struct QpArgs {
  const double* tab[3]; const int* idx[3]; const double* wq[3]; const double* xq[3];
  const double* coef[%(MAXFUN)d]; int n[3]; int nel[3]; long long cell0; double* out;
  long long sstride; int lc; int elast0;
  int co[3]; int cs[3]; int cn[3]; double* y;
};
"""

_PRELUDE = r"""
struct QpArgs {
  const double* tab[3]; const int* idx[3]; const double* wq[3]; const double* xq[3];
  const double* coef[%(MAXFUN)d]; int n[3]; int nel[3]; long long cell0; double* out;
  long long sstride; int lc; int elast0;
};
"""


def generate(prog, dim, nloc, nq, nd, jets, nfun, op=None, diag=False, layout="cell"):
    """CUDA source of the kernel for ``prog``.  jets: list of (fpos, comp, al3)
    in register order; nloc/nq: per-direction sizes (padded to 3 with 1).

    ``op``: list of test multi-indices, one per program output.  The kernel (``tigar_op``)
    then does not store the outputs but contracts output ``s`` with ``D^op[s]`` of the
    test basis (sum-factorised through shared memory, as k_assemble_vector does) and adds
    the element vector into ``A.y``: the action of a bilinear form on a Function in ONE
    kernel, with no coefficient buffer (matrix-free operator, tigar_b200/matfree.py).
    Cells of a launch form the lattice ``co + cs * i`` (colouring: cs = nloc keeps the
    scatter exclusive).

    ``diag=True`` (with ``op`` = list of (test, trial) multi-index pairs, one per output):
    the kernel adds the DIAGONAL of the element matrices instead,
    ``y[g(a)] += sum_s sum_q out_s(q) D^test_s N_a(q) D^trial_s N_a(q)`` -- the Jacobi
    preconditioner of the matrix-free operator without ever assembling a matrix.

    ``layout="gsf"``: the outputs are stored for the global sum-factorised assembly
    (csrc/tg_gsf.cu), ``out[s*sstride + cellperm*NQP + qp]`` with the cells of the launch (whole
    layers ``elast0 .. elast0+lc`` of the last direction) permuted so that the FIRST direction is
    slowest: cellperm = (e0*nel1 + e1)*lc + (e2 - elast0)  [2-D: e0*lc + (e1 - elast0)]."""
    n0, n1, n2 = nloc
    q0, q1, q2 = nq
    nen, nqp = n0 * n1 * n2, q0 * q1 * q2
    nth = ((nqp + 31) // 32) * 32
    L = []
    w = L.append
    w((_PRELUDE if op is None else _PRELUDE_OP) % dict(MAXFUN=MAXFUN))
    w("#define N0 %d\n#define N1 %d\n#define N2 %d\n#define Q0 %d\n#define Q1 %d\n#define Q2 %d"
      % (n0, n1, n2, q0, q1, q2))
    w("#define ND %d\n#define NEN %d\n#define NQP %d\n#define NTH %d\n#define DIM %d"
      % (nd, nen, nqp, nth, dim))
    w('extern "C" __global__ void __launch_bounds__(NTH) %s(const QpArgs A) {'
      % ("tigar_qp" if op is None else "tigar_op"))
    # padded strides (bank conflicts): a table row per Gauss point is N*ND+1 doubles, a row of
    # coefficients N0+1, a plane of the first intermediate N1*Q0+S1PAD
    s1pad = 4 if (n1 * q0) % 16 == 0 else 0
    w("#define TS0 (N0*ND+1)\n#define TS1 (N1*ND+1)\n#define TS2 (N2*ND+1)")
    w("#define CFS ((N0+1)*N1*N2)\n#define S1PL (N1*Q0+%d)\n#define S1P (N2*S1PL)" % s1pad)
    w("  __shared__ double tb0[Q0*TS0], tb1[Q1*TS1], tb2[Q2*TS2];")
    # jets: all functions of a group go through the three contraction stages TOGETHER (one
    # barrier per stage and group instead of one per function and derivative order: the
    # per-function version spent its time in ~40 barriers of 4 FMAs each)
    byf = {}
    for k, (f, comp, al) in enumerate(jets):
        byf.setdefault((f, comp), []).append((k, al))
    S1, S2 = q0 * n1 * n2, q0 * q1 * n2
    groups, cur, cost = [], [], 0
    for key, lst in byf.items():
        c1 = len(set(al[0] for _, al in lst))
        c2 = len(set((al[0], al[1]) for _, al in lst))
        mine = 8 * (nen + c1 * S1 + c2 * S2)
        if cur and cost + mine > 28 * 1024:
            groups.append(cur)
            cur, cost = [], 0
        cur.append(key)
        cost += mine
    if cur:
        groups.append(cur)

    def gsizes(grp):
        c1 = sum(len(set(al[0] for _, al in byf[key])) for key in grp)
        c2 = sum(len(set((al[0], al[1]) for _, al in byf[key])) for key in grp)
        return len(grp), c1, c2
    mg = max([gsizes(g)[0] for g in groups] + [1])
    m1 = max([gsizes(g)[1] for g in groups] + [1])
    m2 = max([gsizes(g)[2] for g in groups] + [1])
    w("  __shared__ double cf[%d*CFS], s1[%d*S1P], s2[%d];" % (mg, m1, m2 * S2))
    w("  const int tid = threadIdx.x;")
    w("  const long long cl = blockIdx.x;")
    if op is None:
        # 32-bit cell arithmetic (launch() checks the range): a 64-bit division is ~100
        # instructions, and every thread of the block decodes the cell
        w("  const unsigned c = (unsigned)(A.cell0 + cl);")
        w("  int e0 = (int)(c % (unsigned)A.nel[0]), e1 = 0, e2 = 0;")
        if dim > 1:
            w("  { const unsigned r = c / (unsigned)A.nel[0]; e1 = (int)(r % (unsigned)A.nel[1]); " +
              ("e2 = (int)(r / (unsigned)A.nel[1]);" if dim > 2 else "") + " }")
    else:
        w("  __shared__ double cq[NQP], u1[N0*Q1*Q2], u2[N0*N1*Q2], accs[NEN];")
        w("  int e0 = A.co[0] + A.cs[0] * (int)(cl % A.cn[0]), e1 = 0, e2 = 0;")
        if dim > 1:
            w("  { long long r = cl / A.cn[0]; e1 = A.co[1] + A.cs[1] * (int)(r % A.cn[1]); " +
              ("e2 = A.co[2] + A.cs[2] * (int)(r / A.cn[1]);" if dim > 2 else "") + " }")
        w("  for (int a = tid; a < NEN; a += NTH) accs[a] = 0.0;")
    for d in range(3):
        Q, N, e = "Q%d" % d, "N%d" % d, "e%d" % d
        src = ("A.tab[%d][(long long)%s*%s*%s*ND + i]" % (d, e, Q, N)) if d < dim else "1.0"
        w("  for (int i = tid; i < %s*%s*ND; i += NTH) tb%d[(i / (%s*ND))*TS%d + i %% (%s*ND)] = %s;"
          % (Q, N, d, N, d, N, src))
    w("  const bool active = tid < NQP;")
    w("  const int qa = tid %% Q0, qb = (tid / Q0) %% Q1, qc = tid / (Q0*Q1);" .replace("%%", "%"))
    w("  double %s;" % ", ".join("j%d = 0.0" % k for k in range(max(len(jets), 1))))
    for grp in groups:
        # stage 0: coefficients of the cell's functions
        w("  __syncthreads();")
        w("  for (int a = tid; a < NEN; a += NTH) {")
        w("    int l0 = a %% N0, t = a / N0, l1 = t %% N1, l2 = t / N1;".replace("%%", "%"))
        w("    long long g = A.idx[0][e0*N0 + l0];")
        if dim > 1:
            w("    g += (long long)A.n[0] * A.idx[1][e1*N1 + l1];")
        if dim > 2:
            w("    g += (long long)A.n[0] * A.n[1] * A.idx[2][e2*N2 + l2];")
        for gi, (f, comp) in enumerate(grp):
            w("    cf[%d*CFS + a + a / N0] = A.coef[%d][g];" % (gi, f))
        w("  }")
        # stage 1: contract direction 0 for every (function, a0)
        slot1, slot2 = {}, {}
        for gi, key in enumerate(grp):
            for a0 in sorted(set(al[0] for _, al in byf[key])):
                slot1[(gi, a0)] = len(slot1)
            for a01 in sorted(set((al[0], al[1]) for _, al in byf[key])):
                slot2[(gi,) + a01] = len(slot2)
        w("  __syncthreads();")
        w("  for (int o = tid; o < Q0*N1*N2; o += NTH) {")
        w("    const int qq = o %% Q0, r = o / Q0;".replace("%%", "%"))
        w("    const int so = (r / N1)*S1PL + (r %% N1)*Q0 + qq;".replace("%%", "%"))
        for (gi, a0), sl in slot1.items():
            w("    { double acc = 0.0;")
            w("      #pragma unroll\n      for (int l = 0; l < N0; l++) acc += cf[%d*CFS + r*(N0+1) + l] * tb0[qq*TS0 + l*ND + %d];" % (gi, a0))
            w("      s1[%d*S1P + so] = acc; }" % sl)
        w("  }")
        # stage 2: contract direction 1 for every (function, a0, a1)
        w("  __syncthreads();")
        w("  for (int o = tid; o < Q0*Q1*N2; o += NTH) {")
        w("    const int x0 = o %% Q0, t = o / Q0, x1 = t %% Q1, l2 = t / Q1;".replace("%%", "%"))
        for (gi, a0, a1), sl in slot2.items():
            w("    { double acc = 0.0;")
            w("      #pragma unroll\n      for (int l = 0; l < N1; l++) acc += s1[%d*S1P + l2*S1PL + l*Q0 + x0] * tb1[x1*TS1 + l*ND + %d];" % (slot1[(gi, a0)], a1))
            w("      s2[%d + o] = acc; }" % (sl * S2))
        w("  }")
        # stage 3: the jets at this thread's Gauss point
        w("  __syncthreads();")
        w("  if (active) {")
        for gi, key in enumerate(grp):
            for k, al in byf[key]:
                w("    { double acc = 0.0;")
                w("      #pragma unroll\n      for (int l = 0; l < N2; l++) acc += s2[%d + (l*Q1 + qb)*Q0 + qa] * tb2[qc*TS2 + l*ND + %d];" % (slot2[(gi, al[0], al[1])] * S2, al[2]))
                w("      j%d = acc; }" % k)
        w("  }")
    if op is None:
        w("  if (!active) return;")
    else:
        if len(op) != len(prog.outregs):
            raise ValueError("one test multi-index per program output")
        w("  double %s;" % ", ".join("ov%d = 0.0" % i for i in range(len(op))))
        w("  if (active) {")
    # fixed registers
    names = {}
    for d in range(dim):
        q = ("qa", "qb", "qc")[d]
        e = ("e0", "e1", "e2")[d]
        Q = ("Q0", "Q1", "Q2")[d]
        w("  const double x%d = A.xq[%d][%s*%s + %s];" % (d, d, e, Q, q))
        names[d] = "x%d" % d
    w("  const double wqv = %s;" % " * ".join(
        "A.wq[%d][%s*%s + %s]" % (d, ("e0", "e1", "e2")[d], ("Q0", "Q1", "Q2")[d],
                                  ("qa", "qb", "qc")[d]) for d in range(dim)))
    names[dim] = "wqv"
    for k in range(len(jets)):
        names[dim + 1 + k] = "j%d" % k
    # straight-line program in SSA form (registers are re-used by the allocator,
    # so every definition gets a fresh C variable)
    ver = 0
    for (opc, dst, a, b) in prog.prog:
        nm = _NAMES[opc]
        var = "t%d" % ver
        ver += 1
        if nm == "const":
            expr = repr(float(prog.consts[a]))
            if expr in ("inf", "-inf", "nan"):
                raise ValueError("non-finite constant in form")
        elif nm in _UN:
            expr = _UN[nm] % names[a]
        else:
            expr = _BIN[nm] % (names[a], names[b])
        w("  const double %s = %s;" % (var, expr))
        names[dst] = var
    if op is None:
        if layout == "gsf":
            if dim == 3:
                w("  double* o = A.out + (((long long)e0 * A.nel[1] + e1) * A.lc + (e2 - A.elast0)) * NQP + tid;")
            elif dim == 2:
                w("  double* o = A.out + ((long long)e0 * A.lc + (e1 - A.elast0)) * NQP + tid;")
            else:
                raise ValueError("gsf layout needs a 2-D or 3-D patch")
            for s, r in enumerate(prog.outregs):
                w("  o[%d * A.sstride] = %s;" % (s, names[r]))
        else:
            w("  double* o = A.out + cl * (long long)%d * NQP + tid;" % len(prog.outregs))
            for s, r in enumerate(prog.outregs):
                w("  o[%d * NQP] = %s;" % (s, names[r]))
        w("}")
        return "\n".join(L), nth
    for s, r in enumerate(prog.outregs):
        w("  ov%d = %s;" % (s, names[r]))
    w("  }")
    # test-function contraction, one output slot at a time (u1, u2 as in k_assemble_vector);
    # diag: the same three sum-factorised stages with the PRODUCT of the test and trial tables
    # (N_a is a tensor product, so sum_q c(q) D^s N_a(q) D^t N_a(q) factorises too: 3*(p+1)^4
    # FMAs per slot and cell instead of (p+1)^6)
    def tabx(d, q, a, al):
        T, TS = "tb%d" % d, "TS%d" % d
        if diag:
            return "%s[%s*%s + %s*ND + %d] * %s[%s*%s + %s*ND + %d]" % (
                T, q, TS, a, al[0][d], T, q, TS, a, al[1][d])
        return "%s[%s*%s + %s*ND + %d]" % (T, q, TS, a, al[d])
    for s, al in enumerate(op):
        if diag:
            al = (tuple(al[0]) + (0,) * (3 - len(al[0])), tuple(al[1]) + (0,) * (3 - len(al[1])))
        else:
            al = tuple(al) + (0,) * (3 - len(al))
        w("  __syncthreads();")
        w("  if (active) cq[tid] = ov%d;" % s)
        w("  __syncthreads();")
        w("  for (int o = tid; o < N0*Q1*Q2; o += NTH) {")
        w("    int a0 = o %% N0, r = o / N0; double acc = 0.0;".replace("%%", "%"))
        w("    #pragma unroll\n    for (int q = 0; q < Q0; q++) acc += %s * cq[r*Q0 + q];"
          % tabx(0, "q", "a0", al))
        w("    u1[o] = acc;\n  }")
        w("  __syncthreads();")
        w("  for (int o = tid; o < N0*N1*Q2; o += NTH) {")
        w("    int a0 = o %% N0, t = o / N0, a1 = t %% N1, q2 = t / N1; double acc = 0.0;"
          .replace("%%", "%"))
        w("    #pragma unroll\n    for (int q = 0; q < Q1; q++) acc += %s * u1[(q2*Q1 + q)*N0 + a0];"
          % tabx(1, "q", "a1", al))
        w("    u2[o] = acc;\n  }")
        w("  __syncthreads();")
        w("  for (int a = tid; a < NEN; a += NTH) {")
        w("    int a01 = a %% (N0*N1), a2 = a / (N0*N1); double acc = 0.0;".replace("%%", "%"))
        w("    #pragma unroll\n    for (int q = 0; q < Q2; q++) acc += %s * u2[q*N0*N1 + a01];"
          % tabx(2, "q", "a2", al))
        w("    accs[a] += acc;\n  }")
    w("  __syncthreads();")
    w("  for (int a = tid; a < NEN; a += NTH) {")
    w("    int l0 = a %% N0, t = a / N0, l1 = t %% N1, l2 = t / N1;".replace("%%", "%"))
    w("    long long g = A.idx[0][e0*N0 + l0];")
    if dim > 1:
        w("    g += (long long)A.n[0] * A.idx[1][e1*N1 + l1];")
    if dim > 2:
        w("    g += (long long)A.n[0] * A.n[1] * A.idx[2][e2*N2 + l2];")
    w("    A.y[g] += accs[a];")
    w("  }")
    w("}")
    return "\n".join(L), nth


_cache = {}


def enabled():
    return not os.environ.get("TIGAR_B200_NO_JIT")


def _signature(prog, *rest):
    """Cheap identity of a kernel request: the register program itself (not the generated
    source: emitting 4 000 lines per Newton step costs more than the launch)."""
    return hash((tuple(map(tuple, prog.prog)), tuple(prog.consts), tuple(prog.outregs), rest))


def get_kernel(prog, dim, nloc, nq, nd, jets, nfun, layout="cell"):
    sig = ("qp", _signature(prog, dim, tuple(nloc), tuple(nq), nd, tuple(jets), nfun, layout))
    k = _cache.get(sig)
    if k is not None:
        return k
    src, nth = generate(prog, dim, nloc, nq, nd, jets, nfun, layout=layout)
    key = hashlib.sha1(src.encode()).hexdigest()
    k = _cache.get(key)
    if k is None:
        h = C.c_void_p()
        check(lib.tg_jit_compile(src.encode(), b"tigar_qp", C.byref(h)))
        k = (h, nth)
        _cache[key] = k
    _cache[sig] = k
    return k


def check_source(src):
    """Compile only (no GPU needed); returns the cubin size."""
    n = c_i64(0)
    check(lib.tg_jit_check(src.encode(), C.byref(n)))
    return n.value


def launch(kernel, B, coef_ptrs, cell0, ncells, out, gsf=None):
    """``gsf=(sstride, lc, elast0)`` for kernels generated with layout="gsf"."""
    h, nth = kernel
    a = QpArgs()
    b = B.c
    for d in range(3):
        a.tab[d], a.idx[d], a.wq[d], a.xq[d] = b.tab[d], b.idx[d], b.wq[d], b.xq[d]
        a.n[d], a.nel[d] = b.n[d], b.nel[d]
    for i, p in enumerate(coef_ptrs):
        a.coef[i] = p
    if cell0 + ncells >= 2 ** 32:
        raise ValueError("more than 2^32 cells in one patch")
    a.cell0 = cell0
    a.out = dev.ptr(out)
    if gsf is not None:
        a.sstride, a.lc, a.elast0 = int(gsf[0]), int(gsf[1]), int(gsf[2])
    check(lib.tg_jit_launch(h, ncells, nth, 0, C.byref(a), C.sizeof(a), dev.stream()))


# ---- fused matrix-free operator kernel (generate(..., op=...)) -------------------------------
def get_op_kernel(prog, dim, nloc, nq, nd, jets, nfun, op, diag=False):
    sig = ("op", _signature(prog, dim, tuple(nloc), tuple(nq), nd, tuple(jets), nfun,
                            tuple(map(tuple, op)) if op and not isinstance(op[0], int)
                            else tuple(op), diag))
    k = _cache.get(sig)
    if k is not None:
        return k
    src, nth = generate(prog, dim, nloc, nq, nd, jets, nfun, op=op, diag=diag)
    key = hashlib.sha1(src.encode()).hexdigest()
    k = _cache.get(key)
    if k is None:
        h = C.c_void_p()
        check(lib.tg_jit_compile(src.encode(), b"tigar_op", C.byref(h)))
        k = (h, nth)
        _cache[key] = k
    _cache[sig] = k
    return k


def launch_op(kernel, B, coef_ptrs, y, stride):
    """One launch per colour of the cell lattice (cells ``stride`` apart per direction never
    share a basis function when stride = nloc), accumulating into ``y``."""
    import itertools
    h, nth = kernel
    a = OpArgs()
    b = B.c
    for d in range(3):
        a.tab[d], a.idx[d], a.wq[d], a.xq[d] = b.tab[d], b.idx[d], b.wq[d], b.xq[d]
        a.n[d], a.nel[d] = b.n[d], b.nel[d]
    for i, p in enumerate(coef_ptrs):
        a.coef[i] = p
    a.y = dev.ptr(y)
    dim = int(b.dim)
    nel = [int(b.nel[d]) if d < dim else 1 for d in range(3)]
    st = [int(stride[d]) if d < dim else 1 for d in range(3)]
    for o in itertools.product(*[range(min(st[d], nel[d])) for d in range(3)]):
        grid = 1
        for d in range(3):
            a.co[d], a.cs[d] = o[d], st[d]
            a.cn[d] = (nel[d] - o[d] + st[d] - 1) // st[d]
            grid *= int(a.cn[d])
        check(lib.tg_jit_launch(h, grid, nth, 0, C.byref(a), C.sizeof(a), dev.stream()))
