"""
Time integrators of the reference's ``tIGAr/timeIntegration.py`` on the
``ufl_lite`` form language: backward Euler, the generalized-alpha method for
first- and second-order systems (implicit midpoint at rho_inf = 1), and a load
stepper.  Every rate is a LINEAR combination of the unknown ``Function`` and the
previous step's data, written once as a coefficient table and turned into a
form-language expression on demand; ``advance()`` moves the data with
``Function.assign`` of those combinations (api.linear_combination).

Not built: ``LinearDGSpaceTimeIntegrator`` (needs mixed function spaces, SURVEY
8f n1).  ``LoadStepper.t`` is a ``ufl_lite.Parameter``: forms written with it see
the new value at their next assembly, like the DOLFIN ``Expression`` parameter the
reference mutates in place (timeIntegration.py:74-93).
"""
from . import ufl_lite as U


def _combo(pairs):
    """sum_k c_k f_k as a form-language expression (zero coefficients skipped)."""
    out = None
    for c, f in pairs:
        if c == 0.0:
            continue
        term = U.Constant(float(c)) * f
        out = term if out is None else out + term
    if out is None:
        raise ValueError("empty linear combination")
    return out


def x_alpha(alpha, x, x_old):
    """alpha-level value alpha x + (1 - alpha) x_old (timeIntegration.py:95-100)."""
    return _combo([(alpha, x), (1.0 - alpha, x_old)])


class _Integrator(object):
    """Shared bookkeeping: ``rates()`` returns, per derivative order, the
    coefficients of (x, x_old, xdot_old[, xddot_old])."""

    def _fields(self):
        f = [self.x, self.x_old]
        if hasattr(self, "xdot_old"):
            f.append(self.xdot_old)
        if hasattr(self, "xddot_old"):
            f.append(self.xddot_old)
        return f

    def _expr(self, coefs):
        return _combo(list(zip(coefs, self._fields())))

    def _advance(self, new_values):
        """new_values: list of (target Function, expression); every expression is
        evaluated with the OLD data before any target is overwritten."""
        from .api import Function
        tmp = []
        for target, expr in new_values:
            t = Function(self.x.function_space())
            t.assign(expr)
            tmp.append((target, t))
        for target, t in tmp:
            target.assign(t)
        self.t += float(self.DELTA_T)


class BackwardEulerIntegrator(_Integrator):
    """timeIntegration.py:13-69.  ``oldFunctions`` = [x_old] (first order) or
    [x_old, xdot_old] (second order)."""

    def __init__(self, DELTA_T, x, oldFunctions, t=0.0):
        self.systemOrder = len(oldFunctions)
        if self.systemOrder not in (1, 2):
            raise ValueError("oldFunctions: [x_old] or [x_old, xdot_old]")
        self.DELTA_T = DELTA_T
        self.x = x
        self.x_old = oldFunctions[0]
        if self.systemOrder == 2:
            self.xdot_old = oldFunctions[1]
        self.t = t + float(DELTA_T)

    def _xdot_coefs(self):
        r = 1.0 / float(self.DELTA_T)
        return [r, -r] + ([0.0] if self.systemOrder == 2 else [])

    def xdot(self):
        return self._expr(self._xdot_coefs())

    def xddot(self):
        if self.systemOrder != 2:
            raise ValueError("xddot of a first-order system")
        r = 1.0 / float(self.DELTA_T)
        v = self._xdot_coefs()
        return self._expr([r * v[0], r * v[1], -r])

    def advance(self):
        new = [(self.x_old, self.x)]
        if self.systemOrder == 2:
            new.append((self.xdot_old, self.xdot()))
        self._advance(new)


class LoadStepper(object):
    """Pseudo-time for problems without time derivatives (timeIntegration.py:71-93)."""

    def __init__(self, DELTA_T, t=0.0):
        self.DELTA_T = DELTA_T
        self.tval = t
        self.t = U.Parameter(t)
        self.advance()

    def advance(self):
        self.tval += float(self.DELTA_T)
        self.t.assign(self.tval)


class GeneralizedAlphaIntegrator(_Integrator):
    """timeIntegration.py:102-247.  ``oldFunctions`` = [x_old, xdot_old] (first
    order) or [x_old, xdot_old, xddot_old] (second order); ``RHO_INF`` is the
    spectral radius of the amplification matrix as the step goes to infinity."""

    def __init__(self, RHO_INF, DELTA_T, x, oldFunctions, t=0.0, useFirstOrderAlphaM=False):
        self.RHO_INF = RHO_INF
        self.DELTA_T = DELTA_T
        self.systemOrder = len(oldFunctions) - 1
        if self.systemOrder not in (1, 2):
            raise ValueError("oldFunctions: [x_old, xdot_old] or [x_old, xdot_old, xddot_old]")
        rho = float(RHO_INF)
        if useFirstOrderAlphaM or self.systemOrder == 1:
            self.ALPHA_M = 0.5 * (3.0 - rho) / (1.0 + rho)
        else:
            self.ALPHA_M = (2.0 - rho) / (1.0 + rho)
        self.ALPHA_F = 1.0 / (1.0 + rho)
        self.GAMMA = 0.5 + self.ALPHA_M - self.ALPHA_F
        self.BETA = 0.25 * (1.0 + self.ALPHA_M - self.ALPHA_F) ** 2
        self.x = x
        self.x_old, self.xdot_old = oldFunctions[0], oldFunctions[1]
        if self.systemOrder == 2:
            self.xddot_old = oldFunctions[2]
        self.t = t + float(DELTA_T)

    # coefficient tables over (x, x_old, xdot_old[, xddot_old])
    def _xdot_coefs(self):
        dt, g, b = float(self.DELTA_T), self.GAMMA, self.BETA
        if self.systemOrder == 1:
            # Newmark velocity update solved for xdot_{n+1}
            return [1.0 / (g * dt), -1.0 / (g * dt), (g - 1.0) / g]
        # Newmark displacement update solved for xddot_{n+1}, inserted in the velocity update
        return [g / (b * dt), -g / (b * dt), 1.0 - g / b,
                (1.0 - g) * dt - (1.0 - 2.0 * b) * dt * g / (2.0 * b)]

    def _xddot_coefs(self):
        if self.systemOrder != 2:
            raise ValueError("xddot of a first-order system")
        dt, g = float(self.DELTA_T), self.GAMMA
        v = self._xdot_coefs()
        s = 1.0 / (dt * g)
        return [s * v[0], s * v[1], s * v[2] - s, s * v[3] - (1.0 - g) / g]

    def xdot(self):
        return self._expr(self._xdot_coefs())

    def xddot(self):
        return self._expr(self._xddot_coefs())

    def x_alpha(self):
        return x_alpha(self.ALPHA_F, self.x, self.x_old)

    def xdot_alpha(self):
        a = self.ALPHA_M if self.systemOrder == 1 else self.ALPHA_F
        return x_alpha(a, self.xdot(), self.xdot_old)

    def xddot_alpha(self):
        return x_alpha(self.ALPHA_M, self.xddot(), self.xddot_old)

    def sameVelocityPredictor(self):
        if self.systemOrder == 1:
            return self.x_old
        dt, g, b = float(self.DELTA_T), self.GAMMA, self.BETA
        return _combo([(1.0, self.x_old), (dt, self.xdot_old),
                       (0.5 * dt * dt * ((1.0 - 2.0 * b) + 2.0 * b * (g - 1.0) / g),
                        self.xddot_old)])

    def advance(self):
        new = [(self.x_old, self.x), (self.xdot_old, self.xdot())]
        if self.systemOrder == 2:
            new.append((self.xddot_old, self.xddot()))
        self._advance(new)


class LinearDGSpaceTimeIntegrator(object):
    def __init__(self, *a, **k):
        raise NotImplementedError("space-time DG needs mixed function spaces (SURVEY 8f n1)")
