"""Names the reference's ``tIGAr/common.py`` exports (its own classes plus the
DOLFIN/UFL/numpy names it re-exports at :10-25), served by ``tigar_b200``."""
import math                                                     # noqa: F401
import sys                                                      # noqa: F401
from numpy import array, zeros, arange                          # noqa: F401

from tigar_b200.api import (                                    # noqa: F401
    AbstractExtractionGenerator, AbstractCoordinateChartSpline, AbstractMultiFieldSpline,
    EqualOrderSpline, FieldListSpline, ExtractedSpline, FunctionSpace, Function,
    TrialFunction, TestFunction, VectorFunction, split, assemble, derivative, File, KrylovSolver, PETScKrylovSolver, SubDomain,
    MPI, worldcomm, selfcomm, mpisize, mpirank, norm, INDEX_TYPE, DEFAULT_PREALLOC,
    DEFAULT_DO_PERMUTATION, DEFAULT_BASIS_FUNC_IGNORE_EPS, USE_DG_DEFAULT, FORM_MT,
    EXTRACTION_DATA_FILE, EXTRACTION_INFO_FILE, EXTRACTION_ZERO_DOFS_FILE,
    EXTRACTION_MAT_FILE, EXTRACTION_MAT_FILE_CTRL)
from tigar_b200.bsplines import (                               # noqa: F401
    AbstractScalarBasis, AbstractControlMesh, DOLFIN_EPS, USE_RECT_ELEM_DEFAULT, near)
from tigar_b200.ufl_lite import (                               # noqa: F401
    pi, inner, dot, outer, cross, conditional, gt, lt, ge, le, tr, det, inv, transpose, grad, sqrt, sin, cos, tan, exp, ln, tanh,
    sinh, cosh, atan, as_vector, as_matrix, as_tensor, Constant, Parameter, lhs, rhs, Form,
    Equation)
from tigar_b200.calculus import (                               # noqa: F401
    getMetric, pinvD, volumeJacobian, cartesianGrad, cartesianDiv, cartesianCurl, getQuadRule,
    getQuadRuleInterval, getChristoffel, CurvilinearTensor, curvilinearInner, covariantDerivative,
    curvilinearGrad, curvilinearDiv, mappedNormal, surfaceJacobian, cartesianPushforwardN,
    cartesianPushforwardRT, cartesianPushforwardW)
