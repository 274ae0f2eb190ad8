"""
A small form language with the UFL/DOLFIN names the tIGAr demos use
(TrialFunction, TestFunction, Function, Constant, inner, dot, grad, sin, ...,
``a == L``, lhs/rhs, assemble), so that ``demos/poisson/poisson.py`` and
``demos/biharmonic/biharmonic.py`` of the reference run unmodified.

In the reference these names come from ``from dolfin import *`` inside
``tIGAr/common.py:10`` and every integrand is compiled by FFC.  Here an
integrand is kept as a polynomial in the derivatives of the test and trial
basis functions,

    sum_k  c_k(xi) * D^{a_k} v * D^{b_k} u ,

whose coefficients c_k are hash-consed scalar DAGs (tigar_b200.symbolic) over
jets of coefficient functions; the CUDA kernels ``tg_qp_eval`` and
``tg_assemble_*`` consume exactly that representation.  Host logic only.
"""
import math

import numpy as np

from . import symbolic as S

pi = math.pi
DOLFIN_EPS = 3.0e-16

ZERO3 = (0, 0, 0)


def _inc(al, j):
    al = list(al)
    al[j] += 1
    return tuple(al)


class Scalar(object):
    """sum over keys (test, trial) of Node * D^test v * D^trial u; a key part is
    None (absent), a 3-multi-index of parametric derivatives (single-field space)
    or ``(a0, a1, a2, field)`` for a basis function of one field of a multi-field
    space (tigar_b200.multifield)."""
    __slots__ = ("terms",)

    def __init__(self, terms=None):
        self.terms = {}
        if terms:
            for k, v in terms.items():
                if v is not S.ZERO:
                    self.terms[k] = v

    @staticmethod
    def coef(node):
        return Scalar({(None, None): S.as_node(node)})

    def arity(self):
        a = 0
        for (t, u) in self.terms:
            a = max(a, (t is not None) + (u is not None))
        return a

    def is_coef(self):
        return all(k == (None, None) for k in self.terms)

    def node(self):
        if not self.is_coef():
            raise ValueError("expression depends on test/trial functions")
        return self.terms.get((None, None), S.ZERO)

    def add(self, o, sign=1.0):
        out = dict(self.terms)
        for k, v in o.terms.items():
            vv = v if sign > 0 else S.neg(v)
            out[k] = S.add(out[k], vv) if k in out else vv
        return Scalar(out)

    def neg(self):
        return Scalar({k: S.neg(v) for k, v in self.terms.items()})

    def mul(self, o):
        out = {}
        for (t1, u1), v1 in self.terms.items():
            for (t2, u2), v2 in o.terms.items():
                if (t1 is not None and t2 is not None) or (u1 is not None and u2 is not None):
                    raise ValueError("form is not linear in its test/trial functions")
                k = (t1 if t1 is not None else t2, u1 if u1 is not None else u2)
                v = S.mul(v1, v2)
                out[k] = S.add(out[k], v) if k in out else v
        return Scalar(out)

    def div(self, o):
        d = o.node()
        return Scalar({k: S.div(v, d) for k, v in self.terms.items()})

    def diff(self, j):
        """d/dxi_j: product rule between coefficient and basis derivative."""
        out = {}

        def acc(k, v):
            if v is S.ZERO:
                return
            out[k] = S.add(out[k], v) if k in out else v
        for (t, u), v in self.terms.items():
            acc((t, u), S.diff(v, j))
            if t is not None and u is not None:
                acc((_inc(t, j), u), v)
                acc((t, _inc(u, j)), v)
            elif t is not None:
                acc((_inc(t, j), u), v)
            elif u is not None:
                acc((t, _inc(u, j)), v)
        return Scalar(out)

    def apply(self, name):
        return Scalar.coef(S.func(name, self.node()))


def _as_scalar(x):
    if isinstance(x, Scalar):
        return x
    if isinstance(x, Tensor):
        if x.a.shape != ():
            raise ValueError("scalar expected")
        return x.a[()]
    if isinstance(x, S.Node):
        return Scalar.coef(x)
    return Scalar.coef(S.const(float(x)))


class Tensor(object):
    """Tensor-valued expression: numpy object array of ``Scalar``."""
    __array_priority__ = 1000

    def __init__(self, a):
        if isinstance(a, Scalar):
            arr = np.empty((), dtype=object)
            arr[()] = a
            a = arr
        self.a = a

    # -- structure
    @property
    def ufl_shape(self):
        return self.a.shape

    def rank(self):
        return self.a.ndim

    def __len__(self):
        return self.a.shape[0]

    def __iter__(self):
        for i in range(self.a.shape[0]):
            yield self[i]

    def __getitem__(self, idx):
        r = self.a[idx]
        if isinstance(r, Scalar):
            return Tensor(r)
        return Tensor(r)

    @property
    def T(self):
        return Tensor(self.a.T.copy())

    def _map(self, f):
        out = np.empty(self.a.shape, dtype=object)
        for i in np.ndindex(*self.a.shape):
            out[i] = f(self.a[i])
        return Tensor(out)

    def _zip(self, o, f):
        if self.a.shape != o.a.shape:
            raise ValueError("shape mismatch %r vs %r" % (self.a.shape, o.a.shape))
        out = np.empty(self.a.shape, dtype=object)
        for i in np.ndindex(*self.a.shape):
            out[i] = f(self.a[i], o.a[i])
        return Tensor(out)

    # -- arithmetic
    def __add__(self, o):
        o = as_tensor_like(o, self)
        return self._zip(o, lambda x, y: x.add(y))

    __radd__ = __add__

    def __sub__(self, o):
        o = as_tensor_like(o, self)
        return self._zip(o, lambda x, y: x.add(y, -1.0))

    def __rsub__(self, o):
        o = as_tensor_like(o, self)
        return o._zip(self, lambda x, y: x.add(y, -1.0))

    def __neg__(self):
        return self._map(lambda x: x.neg())

    def __pos__(self):
        return self

    def __mul__(self, o):
        if isinstance(o, Measure):
            return o.__rmul__(self)
        if isinstance(o, Form):
            return o.__rmul__(self)
        o = as_tensor(o)
        if self.a.ndim == 0:
            s = self.a[()]
            return o._map(lambda y: s.mul(y))
        if o.a.ndim == 0:
            s = o.a[()]
            return self._map(lambda x: x.mul(s))
        if self.a.ndim == 2 and o.a.ndim in (1, 2):      # UFL: A*B is a matrix product
            return dot(self, o)
        raise ValueError("product of rank-%d and rank-%d tensors" % (self.a.ndim, o.a.ndim))

    def __rmul__(self, o):
        return as_tensor(o).__mul__(self)

    def __truediv__(self, o):
        s = _as_scalar(o)
        return self._map(lambda x: x.div(s))

    def __rtruediv__(self, o):
        return as_tensor(o).__truediv__(self)

    def __pow__(self, e):
        s = _as_scalar(self)
        if isinstance(e, int) and e >= 1 and not s.is_coef():
            if e == 1:
                return self
            raise ValueError("form is not linear in its test/trial functions")
        ee = _as_scalar(e).node()
        return Tensor(Scalar.coef(S.power(s.node(), ee)))

    def __eq__(self, o):
        raise TypeError("only Forms can be equated (a == L)")

    __hash__ = None

    def dx(self, j):
        return self._map(lambda x: x.diff(j))


def as_tensor(x):
    if isinstance(x, Tensor):
        return x
    if isinstance(x, (list, tuple)):
        return as_vector(x)
    if isinstance(x, np.ndarray) and x.dtype != object:
        out = np.empty(x.shape, dtype=object)
        for i in np.ndindex(*x.shape):
            out[i] = _as_scalar(float(x[i]))
        return Tensor(out)
    return Tensor(_as_scalar(x))


def as_tensor_like(x, like):
    x = as_tensor(x)
    if x.a.shape == () and like.a.shape != ():
        raise ValueError("cannot add a scalar to a rank-%d tensor" % like.a.ndim)
    return x


def as_vector(comps):
    comps = [as_tensor(c) for c in comps]
    shp = comps[0].a.shape
    out = np.empty((len(comps),) + shp, dtype=object)
    for i, c in enumerate(comps):
        if c.a.shape != shp:
            raise ValueError("ragged tensor")
        out[i] = c.a if shp else c.a[()]
    return Tensor(out)


as_matrix = as_vector


def Constant(v):
    return as_tensor(np.asarray(v, dtype=float)) if np.ndim(v) else as_tensor(float(v))


_param_counter = [0]


class Parameter(Tensor):
    """Mutable scalar: what a dolfin ``Constant`` that is later ``assign``-ed, or an
    ``Expression("t", t=...)`` whose parameter is updated between solves, is in the
    reference's demos (timeIntegration.py:84-93).  Forms keep a symbolic leaf; the
    CURRENT value is baked in each time a form is compiled for assembly."""

    def __init__(self, value=0.0):
        _param_counter[0] += 1
        self.pid = _param_counter[0]
        Tensor.__init__(self, Scalar.coef(S.param(self.pid, float(value))))

    def assign(self, value):
        S.PARAMS[self.pid] = float(value)

    def __float__(self):
        return S.PARAMS[self.pid]

    def values(self):
        return np.array([S.PARAMS[self.pid]])


# ---------------------------------------------------------------- algebra
def inner(a, b):
    a, b = as_tensor(a), as_tensor(b)
    if a.a.shape != b.a.shape:
        raise ValueError("inner: shape mismatch")
    tot = Scalar()
    for i in np.ndindex(*a.a.shape):
        tot = tot.add(a.a[i].mul(b.a[i]))
    return Tensor(tot)


def dot(a, b):
    a, b = as_tensor(a), as_tensor(b)
    if a.a.ndim == 0 or b.a.ndim == 0:
        return a * b
    n = a.a.shape[-1]
    if b.a.shape[0] != n:
        raise ValueError("dot: dimension mismatch")
    shp = a.a.shape[:-1] + b.a.shape[1:]
    out = np.empty(shp, dtype=object)
    for i in np.ndindex(*a.a.shape[:-1]):
        for j in np.ndindex(*b.a.shape[1:]):
            tot = Scalar()
            for k in range(n):
                tot = tot.add(a.a[i + (k,)].mul(b.a[(k,) + j]))
            out[i + j] = tot
    return Tensor(out) if shp else Tensor(out[()])


def outer(a, b):
    a, b = as_tensor(a), as_tensor(b)
    out = np.empty(a.a.shape + b.a.shape, dtype=object)
    for i in np.ndindex(*a.a.shape):
        for j in np.ndindex(*b.a.shape):
            out[i + j] = a.a[i].mul(b.a[j])
    return Tensor(out)


def cross(a, b):
    """UFL ``cross`` of two 3-vectors (kl-shell-svk/dynamic-tspline.py:146)."""
    a, b = as_tensor(a), as_tensor(b)
    if a.a.shape != (3,) or b.a.shape != (3,):
        raise ValueError("cross: two 3-vectors expected")
    out = np.empty((3,), dtype=object)
    for i in range(3):
        j, k = (i + 1) % 3, (i + 2) % 3
        out[i] = a.a[j].mul(b.a[k]).add(a.a[k].mul(b.a[j]), -1.0)
    return Tensor(out)


def _compare(kind):
    def f(a, b):
        na, nb = _as_scalar(a).node(), _as_scalar(b).node()
        if kind == "gt":
            c = S.binary("gt", na, nb)
        elif kind == "lt":
            c = S.binary("gt", nb, na)
        elif kind == "ge":
            c = S.sub(S.ONE, S.binary("gt", nb, na))
        else:
            c = S.sub(S.ONE, S.binary("gt", na, nb))
        return Tensor(Scalar.coef(c))
    f.__name__ = kind
    return f


gt, lt, ge, le = _compare("gt"), _compare("lt"), _compare("ge"), _compare("le")


def conditional(cond, a, b):
    """UFL ``conditional``: ``cond`` is a 0/1 coefficient expression (gt/lt/ge/le); the
    branches may contain test/trial functions."""
    c = _as_scalar(cond)
    a, b = as_tensor(a), as_tensor(b)
    cn = c.terms.get((None, None), S.ZERO) if set(c.terms) <= {(None, None)} else None
    if cn is None:
        raise ValueError("conditional: the condition must not contain test/trial functions")
    ncn = S.sub(S.ONE, cn)

    def sel(x, y):
        # a true select on every coefficient: term-wise (cond ? x : 0) + (!cond ? y : 0)
        out = Scalar({k: S.selz(cn, v) for k, v in x.terms.items()})
        return out.add(Scalar({k: S.selz(ncn, v) for k, v in y.terms.items()}))
    return a._zip(b, sel)


def tr(a):
    a = as_tensor(a)
    tot = Scalar()
    for i in range(a.a.shape[0]):
        tot = tot.add(a.a[i, i])
    return Tensor(tot)


def transpose(a):
    return as_tensor(a).T


def _node_matrix(a):
    a = as_tensor(a)
    n, m = a.a.shape
    return [[a.a[i, j].node() for j in range(m)] for i in range(n)]


def det(a):
    m = _node_matrix(a)
    n = len(m)
    if n == 1:
        d = m[0][0]
    elif n == 2:
        d = m[0][0] * m[1][1] - m[0][1] * m[1][0]
    elif n == 3:
        d = (m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1])
             - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0])
             + m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]))
    else:
        raise NotImplementedError("det of %dx%d" % (n, n))
    return Tensor(Scalar.coef(d))


def inv(a):
    m = _node_matrix(a)
    n = len(m)
    d = det(a).a[()].node()
    if n == 1:
        adj = [[S.ONE]]
    elif n == 2:
        adj = [[m[1][1], -m[0][1]], [-m[1][0], m[0][0]]]
    elif n == 3:
        def cof(i, j):
            r = [x for x in range(3) if x != i]
            c = [x for x in range(3) if x != j]
            v = m[r[0]][c[0]] * m[r[1]][c[1]] - m[r[0]][c[1]] * m[r[1]][c[0]]
            return v if (i + j) % 2 == 0 else -v
        adj = [[cof(j, i) for j in range(3)] for i in range(3)]
    else:
        raise NotImplementedError("inv of %dx%d" % (n, n))
    out = np.empty((n, n), dtype=object)
    for i in range(n):
        for j in range(n):
            out[i, j] = Scalar.coef(S.div(adj[i][j], d))
    return Tensor(out)


def _unary(name):
    def f(x):
        if isinstance(x, (int, float)):
            return S._PYF[name](float(x))
        return as_tensor(x)._map(lambda s: s.apply(name))
    f.__name__ = name
    return f


sqrt = _unary("sqrt")
sin = _unary("sin")
cos = _unary("cos")
tan = _unary("tan")
exp = _unary("exp")
ln = _unary("log")
tanh = _unary("tanh")
sinh = _unary("sinh")
cosh = _unary("cosh")
atan = _unary("atan")


def abs_(x):
    return as_tensor(x)._map(lambda s: s.apply("abs"))


DEFAULT_DIM = [None]      # parametric dimension of the most recent ExtractedSpline


def grad(f):
    """dolfin ``grad``: derivative w.r.t. the mesh (= parametric) coordinates."""
    if DEFAULT_DIM[0] is None:
        raise RuntimeError("grad() needs an ExtractedSpline to fix the parametric dimension")
    return parametric_grad(f, DEFAULT_DIM[0])


def parametric_grad(f, dim):
    """UFL ``grad`` on the parametric mesh: appends an axis of length dim."""
    f = as_tensor(f)
    out = np.empty(f.a.shape + (dim,), dtype=object)
    for i in np.ndindex(*f.a.shape):
        for j in range(dim):
            out[i + (j,)] = f.a[i].diff(j)
    return Tensor(out)


def curl_from_grad(f, gradf):
    """``cartesianCurl`` (calculusUtils.py:278-302) given ``gradf[a, j] = d f_a / d x_j``
    (or ``gradf[j]`` for a scalar ``f``): rank 1 in 3-D -> vector
    ``eps_ijk gradf[k, j]``; rank 1 in 2-D -> scalar ``gradf[1,0] - gradf[0,1]``;
    scalar in 2-D -> vector ``(-gradf[1], gradf[0])``."""
    f, g = as_tensor(f), as_tensor(gradf)
    if f.a.ndim == 1:
        m = f.a.shape[0]
        if g.a.shape != (m, m):
            raise ValueError("curl: gradient of shape %r for a %d-vector" % (g.a.shape, m))
        if m == 3:
            return as_vector([Tensor(g.a[2, 1].add(g.a[1, 2], -1.0)),
                              Tensor(g.a[0, 2].add(g.a[2, 0], -1.0)),
                              Tensor(g.a[1, 0].add(g.a[0, 1], -1.0))])
        if m == 2:
            return Tensor(g.a[1, 0].add(g.a[0, 1], -1.0))
        raise ValueError("Unsupported dimension of argument to curl.")
    if f.a.ndim == 0:
        if g.a.shape != (2,):
            raise ValueError("curl of a scalar is defined in 2-D only")
        return as_vector([Tensor(g.a[1].neg()), Tensor(g.a[0])])
    raise ValueError("Unsupported rank of argument to curl.")


def expression_from_string(expr, coords):
    """``Expression(expr, degree=...)`` of the reference (common.py:1111-1117) for the
    C++-style scalar expressions the demos use: ``x[i]`` are the given coordinate
    tensors, ``pi`` / ``DOLFIN_PI``, ``pow``, ``sqrt``, ``exp``, ``log``, ``sin``,
    ``cos``, ``tan``, ``fabs``/``abs`` are available.  A tuple/list of strings gives
    a vector."""
    import math
    if isinstance(expr, (tuple, list)):
        return as_vector([expression_from_string(e, coords) for e in expr])
    ns = {"x": list(coords), "pi": math.pi, "DOLFIN_PI": math.pi, "pow": lambda a, b: a ** b,
          "sqrt": sqrt, "exp": exp, "log": ln, "sin": sin, "cos": cos, "tan": tan,
          "fabs": abs_, "abs": abs_, "__builtins__": {}}
    return as_tensor(eval(str(expr), ns))     # noqa: S307 (user-supplied formula, as in the reference)


# ------------------------------------------------------------ forms
class Measure(object):
    """Weighted volume measure (tIGArMeasure, calculusUtils.py:351-410):
    ``f*dx`` integrates f*J over the patch with the spline's Gauss rule."""

    def __init__(self, J, owner, kind="dx"):
        self.J = J            # Tensor scalar or None
        self.owner = owner    # the ExtractedSpline (assembly context)
        self.kind = kind

    def __rmul__(self, other):
        if self.kind != "dx":
            raise NotImplementedError("boundary integrals (ds) are outside the built hot path")
        f = _as_scalar(other)
        if self.J is not None:
            f = f.mul(_as_scalar(self.J))
        return Form([(f, self.owner)])

    def __call__(self, *a, **k):
        return self


class Form(object):
    def __init__(self, integrals):
        self.integrals = integrals      # list of (Scalar, owner)

    def owner(self):
        for _, o in self.integrals:
            return o
        return None

    def scalar(self):
        tot = Scalar()
        for s, _ in self.integrals:
            tot = tot.add(s)
        return tot

    def arity(self):
        return self.scalar().arity()

    def empty(self):
        return not self.scalar().terms

    def __add__(self, o):
        if isinstance(o, (int, float)) and o == 0:
            return self
        return Form(self.integrals + o.integrals)

    __radd__ = __add__

    def __neg__(self):
        return Form([(s.neg(), o) for s, o in self.integrals])

    def __sub__(self, o):
        return self + (-o)

    def __mul__(self, c):
        cs = _as_scalar(c)
        return Form([(s.mul(cs), o) for s, o in self.integrals])

    __rmul__ = __mul__

    def __eq__(self, o):
        return Equation(self, o)

    __hash__ = None

    def part(self, arity):
        out = []
        for s, o in self.integrals:
            t = {k: v for k, v in s.terms.items()
                 if (k[0] is not None) + (k[1] is not None) == arity}
            if t:
                out.append((Scalar(t), o))
        return Form(out)


class Equation(object):
    def __init__(self, lhs, rhs):
        self.lhs = lhs
        self.rhs = rhs


def lhs(form):
    return form.part(2)


def rhs(form):
    return -form.part(1)


def gateaux(form, fid, test=False):
    """d/d(eps) form(u + eps*du) at eps=0 for the coefficient function with id
    ``fid`` (UFL ``derivative(form, u, du)``).  ``du`` is the trial function, or, with
    ``test=True``, the test function (first variation of an energy functional,
    kl-shell-svk/dynamic-tspline.py:232); the differentiated terms must not already
    contain that argument.  For a multi-field Function ``fid`` is a dict
    {component fid: field}: the new key part then carries the field,
    ``(a0, a1, a2, field)``."""
    fields = fid if isinstance(fid, dict) else {fid: None}
    out = []
    for sc, owner in form.integrals:
        acc = Scalar()
        for (t, tr), coef in sc.terms.items():
            if (t if test else tr) is not None:
                raise ValueError("derivative() of a form that already has a %s function"
                                 % ("test" if test else "trial"))
            for jn in S.jets_of([coef]):
                if jn.args[0] not in fields:
                    continue
                dc = S.diff_leaf(coef, jn)
                if dc is not S.ZERO:
                    fld = fields[jn.args[0]]
                    key = tuple(jn.args[2]) if fld is None else tuple(jn.args[2]) + (fld,)
                    acc = acc.add(Scalar({((key, tr) if test else (t, key)): dc}))
        if acc.terms:
            out.append((acc, owner))
    return Form(out)
