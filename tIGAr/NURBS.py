"""``tIGAr.NURBS`` of the reference (NURBS.py), served by ``tigar_b200``.
igakit is not required: any object with ``degree``, ``knots``, ``control``
(igakit's NURBS attributes) is accepted; ``tigar_b200.nurbs.NURBS`` is a
minimal stand-in with ``refine`` / ``elevate``."""
from tIGAr.common import *                                      # noqa: F401,F403
from tIGAr.BSplines import *                                    # noqa: F401,F403
from tigar_b200.nurbs import (NURBSControlMesh, NURBS, quarter_annulus,     # noqa: F401
                               cylindrical_roof)
