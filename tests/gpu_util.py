"""Helpers shared by the -m gpu parity tests: build the same problem in the
CUDA product (through the tIGAr API) and in the oracle."""
import math

import numpy as np

from oracle import bsplines as OB
from oracle import pipeline as OP


def make_pair(deg, kvecs, quadDeg=None, nLayers=1, form="poisson", mode="csr"):
    """(generator, ExtractedSpline, oracle Problem) for an explicit B-spline patch."""
    from tIGAr import EqualOrderSpline, ExtractedSpline
    from tIGAr.BSplines import ExplicitBSplineControlMesh
    cm = ExplicitBSplineControlMesh(deg, kvecs)
    gen = EqualOrderSpline(1, cm)
    sp = gen.getScalarSpline(0)
    for d in range(len(deg)):
        for side in (0, 1):
            gen.addZeroDofs(0, sp.getSideDofs(d, side, nLayers))
    qd = 2 * max(deg) if quadDeg is None else quadDeg
    spline = ExtractedSpline(gen, qd, mode=mode)
    pr = OP.Problem(deg, kvecs, form=form, nLayers=nLayers, quadDeg=qd)
    return gen, spline, pr


def uk(p, n, lo=0.0, hi=1.0):
    return OB.uniform_knots(p, lo, hi, n)


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def relm(A, B):
    import scipy.sparse.linalg as spla
    return spla.norm(A - B) / spla.norm(B)
