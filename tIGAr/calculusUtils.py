"""``tIGAr.calculusUtils`` of the reference, served by ``tigar_b200.calculus``."""
from tigar_b200.calculus import (                               # noqa: F401
    getMetric, pinvD, volumeJacobian, cartesianGrad, cartesianDiv, cartesianCurl, getQuadRule,
    getQuadRuleInterval, getChristoffel, CurvilinearTensor, curvilinearInner, covariantDerivative,
    curvilinearGrad, curvilinearDiv, mappedNormal, surfaceJacobian, cartesianPushforwardN,
    cartesianPushforwardRT, cartesianPushforwardW)
