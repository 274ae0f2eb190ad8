"""
Scalar expression DAG -> register-machine program for ``tg_qp_eval``.

In the reference, every integrand is a UFL expression that FFC compiles to a C
``tabulate_tensor`` (common.py:1215-1216, calculusUtils.py).  Neither UFL nor
FFC exists here; this module is the part of that pipeline the hot path needs:
hash-consed scalar expressions over
  * the parametric coordinates xi_d and the quadrature weight,
  * jets  D^alpha f_c  of coefficient functions (FE or spline basis),
with symbolic differentiation d/dxi_j, compiled to the (op,dst,a,b) int32
program the CUDA kernel ``k_qp_eval`` interprets once per Gauss point.
Host logic only: nothing here evaluates numbers.
"""
import math

# opcodes: keep in sync with tigar_b200/csrc/tg_qp.cu
OPCODES = dict(nop=0, const=1, mov=2, add=3, sub=4, mul=5, div=6, neg=7, sin=8, cos=9,
               exp=10, log=11, sqrt=12, pow=13, abs=14, tan=15, tanh=16, max=17, min=18,
               sinh=19, cosh=20, atan=21, gt=22, selz=23)
_UNARY = ("neg", "sin", "cos", "exp", "log", "sqrt", "abs", "tan", "tanh", "sinh", "cosh",
          "atan")
_BINARY = ("add", "sub", "mul", "div", "pow", "max", "min", "gt", "selz")

_table = {}
_counter = [0]


class Node(object):
    """Immutable, hash-consed scalar expression."""
    __slots__ = ("op", "args", "uid", "_d")

    def __init__(self, op, args):
        self.op = op
        self.args = args
        self.uid = _counter[0]
        _counter[0] += 1
        self._d = None

    # arithmetic sugar
    def __add__(self, o): return add(self, as_node(o))
    def __radd__(self, o): return add(as_node(o), self)
    def __sub__(self, o): return sub(self, as_node(o))
    def __rsub__(self, o): return sub(as_node(o), self)
    def __mul__(self, o): return mul(self, as_node(o))
    def __rmul__(self, o): return mul(as_node(o), self)
    def __truediv__(self, o): return div(self, as_node(o))
    def __rtruediv__(self, o): return div(as_node(o), self)
    def __neg__(self): return neg(self)
    def __pow__(self, o): return power(self, as_node(o))

    def is_const(self):
        return self.op == "const"

    def __repr__(self):
        if self.op == "const":
            return repr(self.args[0])
        if self.op in ("xi", "wq", "jet", "param"):
            return "%s%r" % (self.op, self.args)
        return "%s(%s)" % (self.op, ",".join(repr(a) for a in self.args))


def _mk(op, *args):
    key = (op,) + tuple(a.uid if isinstance(a, Node) else a for a in args)
    n = _table.get(key)
    if n is None:
        n = Node(op, args)
        _table[key] = n
    return n


def const(v):
    v = float(v)
    if v == 0.0:
        v = 0.0            # fold -0.0
    return _mk("const", v)


ZERO = const(0.0)
ONE = const(1.0)


def as_node(x):
    if isinstance(x, Node):
        return x
    return const(x)


def xi(d):
    return _mk("xi", int(d))


def wq():
    return _mk("wq")


PARAMS = {}          # pid -> current value of a mutable scalar parameter


def param(pid, value=None):
    """Mutable scalar parameter (a dolfin ``Constant`` / ``Expression`` parameter that is
    re-assigned between solves, e.g. the load-stepping time, timeIntegration.py:84-93):
    a leaf whose CURRENT value is baked into the program each time it is compiled."""
    pid = int(pid)
    if value is not None:
        PARAMS[pid] = float(value)
    return _mk("param", pid)


def freeze_params(n):
    """``n`` with every parameter leaf replaced by its current value."""
    leaves = []
    seen, stack = set(), [n]
    while stack:
        m = stack.pop()
        if m.uid in seen:
            continue
        seen.add(m.uid)
        if m.op == "param":
            leaves.append(m)
        elif m.op not in ("const", "xi", "wq", "jet"):
            stack.extend(m.args)
    if not leaves:
        return n
    return substitute(n, {m: const(PARAMS[m.args[0]]) for m in leaves})


def jet(fid, comp, alpha):
    """D^alpha of component ``comp`` of coefficient function ``fid``."""
    return _mk("jet", int(fid), int(comp), tuple(int(a) for a in alpha))


def _c(n):
    return n.args[0]


def add(a, b):
    if a.is_const() and b.is_const():
        return const(_c(a) + _c(b))
    if a is ZERO:
        return b
    if b is ZERO:
        return a
    if b.op == "neg":
        return sub(a, b.args[0])
    if a.op == "neg":
        return sub(b, a.args[0])
    if a.uid > b.uid:
        a, b = b, a
    return _mk("add", a, b)


def sub(a, b):
    if a.is_const() and b.is_const():
        return const(_c(a) - _c(b))
    if b is ZERO:
        return a
    if a is ZERO:
        return neg(b)
    if a is b:
        return ZERO
    if b.op == "neg":
        return add(a, b.args[0])
    return _mk("sub", a, b)


def mul(a, b):
    if a.is_const() and b.is_const():
        return const(_c(a) * _c(b))
    if a is ZERO or b is ZERO:
        return ZERO
    if a is ONE:
        return b
    if b is ONE:
        return a
    if a.is_const() and _c(a) == -1.0:
        return neg(b)
    if b.is_const() and _c(b) == -1.0:
        return neg(a)
    if a.op == "neg" and b.op == "neg":
        return mul(a.args[0], b.args[0])
    if a.op == "neg":
        return neg(mul(a.args[0], b))
    if b.op == "neg":
        return neg(mul(a, b.args[0]))
    if a.uid > b.uid:
        a, b = b, a
    return _mk("mul", a, b)


def div(a, b):
    if b.is_const():
        if _c(b) == 0.0:
            raise ZeroDivisionError("symbolic division by zero")
        if a.is_const():
            return const(_c(a) / _c(b))
        if b is ONE:
            return a
    if a is ZERO:
        return ZERO
    if a.op == "neg":
        return neg(div(a.args[0], b))
    return _mk("div", a, b)


def neg(a):
    if a.is_const():
        return const(-_c(a))
    if a.op == "neg":
        return a.args[0]
    if a.op == "sub":
        return sub(a.args[1], a.args[0])
    return _mk("neg", a)


def power(a, b):
    if b.is_const():
        e = _c(b)
        if e == 0.0:
            return ONE
        if e == 1.0:
            return a
        if e == 2.0:
            return mul(a, a)
        if e == 3.0:
            return mul(mul(a, a), a)
        if e == 0.5:
            return func("sqrt", a)
        if e == -1.0:
            return div(ONE, a)
        if a.is_const():
            return const(_c(a) ** e)
    return _mk("pow", a, b)


_PYF = dict(sin=math.sin, cos=math.cos, exp=math.exp, log=math.log, sqrt=math.sqrt,
            abs=abs, tan=math.tan, tanh=math.tanh, sinh=math.sinh, cosh=math.cosh,
            atan=math.atan)


def func(name, a):
    a = as_node(a)
    if name == "neg":
        return neg(a)
    if a.is_const():
        return const(_PYF[name](_c(a)))
    return _mk(name, a)


def binary(name, a, b):
    a, b = as_node(a), as_node(b)
    if name == "add":
        return add(a, b)
    if name == "sub":
        return sub(a, b)
    if name == "mul":
        return mul(a, b)
    if name == "div":
        return div(a, b)
    if name == "pow":
        return power(a, b)
    if name == "selz":
        return selz(a, b)
    if a.is_const() and b.is_const():
        x, y = _c(a), _c(b)
        return const(dict(max=max(x, y), min=min(x, y), gt=1.0 if x > y else 0.0)[name])
    return _mk(name, a, b)


def selz(cond, a):
    """(cond != 0) ? a : 0 -- a TRUE select: an inf/NaN in an unselected ``a`` does not reach
    the result (cond * a would give 0 * inf = NaN)."""
    cond, a = as_node(cond), as_node(a)
    if a is ZERO or cond is ZERO:
        return ZERO
    if cond.is_const():
        return a if _c(cond) != 0.0 else ZERO
    return _mk("selz", cond, a)


def select(cond, a, b):
    """cond in {0,1}: a where cond != 0, b elsewhere (UFL ``conditional`` is a select, not an
    arithmetic blend: the usual use guards a singularity in the other branch)."""
    cond, a, b = as_node(cond), as_node(a), as_node(b)
    if a is b:
        return a
    return add(selz(cond, a), selz(sub(ONE, cond), b))


# ------------------------------------------------------------ differentiation
def diff(n, j):
    """d n / d xi_j  (jets raise their multi-index; wq and constants are 0)."""
    if n._d is None:
        n._d = {}
    r = n._d.get(j)
    if r is None:
        r = _diff(n, j)
        n._d[j] = r
    return r


def _diff(n, j):
    op = n.op
    if op in ("const", "wq", "param"):
        return ZERO
    if op == "xi":
        return ONE if n.args[0] == j else ZERO
    if op == "jet":
        f, c, al = n.args
        al = list(al)
        al[j] += 1
        return jet(f, c, al)
    return _diff_rule(n, lambda c: diff(c, j))


def diff_leaf(n, leaf, _memo=None):
    """Partial derivative of ``n`` with respect to the leaf node ``leaf`` (a jet),
    every other leaf held fixed: the building block of the Gateaux derivative
    (UFL ``derivative``)."""
    if _memo is None:
        _memo = {}
    r = _memo.get(n.uid)
    if r is None:
        if n is leaf:
            r = ONE
        elif n.op in ("const", "wq", "xi", "jet", "param"):
            r = ZERO
        else:
            r = _diff_rule(n, lambda c: diff_leaf(c, leaf, _memo))
        _memo[n.uid] = r
    return r


def _diff_rule(n, d):
    """Differentiation rules of the interior nodes; ``d(child)`` differentiates
    a child."""
    op = n.op
    a = n.args[0]
    da = d(a)
    if op == "neg":
        return neg(da)
    if op == "selz":                     # the condition is piecewise constant
        return selz(a, d(n.args[1]))
    if op in ("add", "sub", "mul", "div", "pow", "max", "min", "gt"):
        b = n.args[1]
        db = d(b)
        if op == "add":
            return add(da, db)
        if op == "sub":
            return sub(da, db)
        if op == "mul":
            return add(mul(da, b), mul(a, db))
        if op == "div":
            # (da - n*db)/b
            return div(sub(da, mul(n, db)), b)
        if op == "pow":
            if b.is_const():
                return mul(mul(b, power(a, const(_c(b) - 1.0))), da)
            return mul(n, add(mul(db, func("log", a)), mul(b, div(da, a))))
        if op == "max":
            return select(binary("gt", a, b), da, db)
        if op == "min":
            return select(binary("gt", a, b), db, da)
        if op == "gt":
            return ZERO
    if da is ZERO:
        return ZERO
    if op == "sin":
        return mul(func("cos", a), da)
    if op == "cos":
        return neg(mul(func("sin", a), da))
    if op == "exp":
        return mul(n, da)
    if op == "log":
        return div(da, a)
    if op == "sqrt":
        return div(da, mul(const(2.0), n))
    if op == "abs":
        return mul(sub(mul(const(2.0), binary("gt", a, ZERO)), ONE), da)
    if op == "tan":
        return mul(add(ONE, mul(n, n)), da)
    if op == "tanh":
        return mul(sub(ONE, mul(n, n)), da)
    if op == "sinh":
        return mul(func("cosh", a), da)
    if op == "cosh":
        return mul(func("sinh", a), da)
    if op == "atan":
        return div(da, add(ONE, mul(a, a)))
    raise NotImplementedError("diff of " + op)


def substitute(n, mapping, _memo=None):
    """Replace nodes (by identity) according to ``mapping`` {Node: Node}."""
    if _memo is None:
        _memo = {}
    r = _memo.get(n.uid)
    if r is not None:
        return r
    if n in mapping:
        r = mapping[n]
    elif n.op in ("const", "xi", "wq", "jet", "param"):
        r = n
    else:
        args = [substitute(a, mapping, _memo) for a in n.args]
        if n.op in _UNARY:
            r = func(n.op, args[0])
        else:
            r = binary(n.op, args[0], args[1])
    _memo[n.uid] = r
    return r


def jets_of(nodes):
    """All jet leaves reachable from ``nodes`` (deterministic order)."""
    seen, out, stack = set(), [], list(nodes)
    while stack:
        n = stack.pop()
        if n.uid in seen:
            continue
        seen.add(n.uid)
        if n.op == "jet":
            out.append(n)
        elif n.op not in ("const", "xi", "wq", "param"):
            stack.extend(n.args)
    out.sort(key=lambda n: n.args)
    return out


def max_order(nodes):
    m = 0
    for jn in jets_of(nodes):
        m = max(m, max(jn.args[2]))
    return m


# ------------------------------------------------------------------ compiler
class Program(object):
    """Compiled register program.

    registers: 0..dim-1 = xi, dim = wq, dim+1.. = jets (in ``jets`` order),
    then temporaries (re-used once dead).
    """

    def __init__(self, dim, jets, prog, consts, nreg, outregs):
        self.dim = dim
        self.jets = jets          # list of (fid, comp, (a1,a2,a3))
        self.prog = prog          # list of (op, dst, a, b)
        self.consts = consts      # list of float
        self.nreg = nreg
        self.outregs = outregs


def share_reciprocals(outputs):
    """The DAGs with every denominator that divides two or more numerators inverted ONCE
    (``a/b -> a*(1/b)``): an FP64 division is ~15 instructions with a slow-path branch, and the
    geometry (P/w, g^-1) divides 9-12 numerators by the same weight / determinant.  Changes the
    rounding of those quotients by <= 1 ulp; the host interpreter runs the same program."""
    nodes, seen, stack = [], set(), list(outputs)
    while stack:
        n = stack.pop()
        if n.uid in seen:
            continue
        seen.add(n.uid)
        nodes.append(n)
        if n.op not in ("const", "xi", "wq", "jet", "param"):
            stack.extend(n.args)
    uses = {}
    for n in nodes:
        if n.op == "div" and not n.args[1].is_const():
            uses[n.args[1].uid] = uses.get(n.args[1].uid, 0) + 1
    shared = set(u for u, c in uses.items() if c >= 2)
    if not shared:
        return outputs
    memo = {}
    # children were created before their parents (hash-consing), so uid order is topological
    for n in sorted(nodes, key=lambda m: m.uid):
        if n.op in ("const", "xi", "wq", "jet", "param"):
            memo[n.uid] = n
            continue
        args = [memo[a.uid] for a in n.args]
        if n.op == "div" and n.args[1].uid in shared:
            r = mul(args[0], _mk("div", ONE, args[1]))
        elif all(a is b for a, b in zip(args, n.args)):
            r = n
        elif n.op in _UNARY:
            r = func(n.op, args[0])
        else:
            r = binary(n.op, args[0], args[1])
        memo[n.uid] = r
    return [memo[o.uid] for o in outputs]


def compile_program(outputs, dim):
    """Topologically order the DAG under ``outputs`` and allocate registers."""
    outputs = share_reciprocals([as_node(o) for o in outputs])
    order, seen = [], set()
    # iterative post-order
    for root in outputs:
        stack = [(root, 0)]
        while stack:
            n, i = stack.pop()
            if n.uid in seen:
                continue
            kids = n.args if n.op not in ("const", "xi", "wq", "jet", "param") else ()
            if i < len(kids):
                stack.append((n, i + 1))
                if kids[i].uid not in seen:
                    stack.append((kids[i], 0))
            else:
                seen.add(n.uid)
                order.append(n)
    jets = [n for n in order if n.op == "jet"]
    jets.sort(key=lambda n: n.args)
    reg = {}
    for n in order:
        if n.op == "xi":
            if n.args[0] >= dim:
                raise ValueError("xi index beyond parametric dimension")
            reg[n.uid] = n.args[0]
        elif n.op == "wq":
            reg[n.uid] = dim
    for k, n in enumerate(jets):
        reg[n.uid] = dim + 1 + k
    nfixed = dim + 1 + len(jets)
    # last use of each node
    last = {}
    for pos, n in enumerate(order):
        if n.op in ("const", "xi", "wq", "jet", "param"):
            continue
        for a in n.args:
            last[a.uid] = pos
    outset = set(o.uid for o in outputs)
    free, nreg = [], nfixed
    prog, consts, cidx = [], [], {}
    for pos, n in enumerate(order):
        if n.op in ("xi", "wq", "jet"):
            continue
        isc = n.op in ("const", "param")          # a parameter is a constant of THIS program
        if isc:
            v = n.args[0] if n.op == "const" else float(PARAMS[n.args[0]])
            if v not in cidx:
                cidx[v] = len(consts)
                consts.append(v)
        # release operands whose last use is here (their registers may be the dst)
        ops = [] if isc else [reg[a.uid] for a in n.args]
        if not isc:
            for a in n.args:
                if (last.get(a.uid) == pos and a.uid not in outset
                        and reg[a.uid] >= nfixed and reg[a.uid] not in free):
                    free.append(reg[a.uid])
        if free:
            dst = free.pop()
        else:
            dst = nreg
            nreg += 1
        reg[n.uid] = dst
        if isc:
            prog.append((OPCODES["const"], dst, cidx[v], 0))
        elif len(ops) == 1:
            prog.append((OPCODES[n.op], dst, ops[0], ops[0]))
        else:
            prog.append((OPCODES[n.op], dst, ops[0], ops[1]))
    outregs = [reg[o.uid] for o in outputs]
    return Program(dim, [n.args for n in jets], prog, consts, max(nreg, 1), outregs)
