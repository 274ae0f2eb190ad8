"""``tIGAr.BSplines`` of the reference (BSplines.py), served by ``tigar_b200``."""
from tIGAr.common import *                                      # noqa: F401,F403
from tigar_b200.bsplines import (                               # noqa: F401
    uniformKnots, KNOT_NEAR_EPS, basisFuncsInner, BSpline1, ij2dof, ijk2dof, dof2ij, dof2ijk, BSpline,
    ExplicitBSplineControlMesh, TensorMesh)
