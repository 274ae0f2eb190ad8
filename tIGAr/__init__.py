"""
Drop-in ``tIGAr`` package backed by the B200-native CUDA library
(``tigar_b200``).  Same import surface as the reference's
``tIGAr/__init__.py:1`` (``from tIGAr.common import *``), so
``demos/poisson/poisson.py`` and ``demos/biharmonic/biharmonic.py`` run
unmodified.
"""
from tIGAr.common import *          # noqa: F401,F403
