"""``tIGAr.timeIntegration`` of the reference, served by ``tigar_b200.time_integration``."""
from tIGAr.common import *                                      # noqa: F401,F403
from tigar_b200.time_integration import (                       # noqa: F401
    BackwardEulerIntegrator, LoadStepper, x_alpha, GeneralizedAlphaIntegrator,
    LinearDGSpaceTimeIntegrator)
