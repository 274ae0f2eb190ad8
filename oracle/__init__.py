"""
CPU oracle for the tIGAr extraction -> assemble -> M^T A M -> solve hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tigar_b200/`` (the product) may
import this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker or as the timed CPU baseline.

Parity status: the reference ships no tests or golden vectors.
The B-spline layer of this oracle (knots, spans, Cox-de Boor, tensor-product
``getNodesAndEvals``, Greville points, side DoFs) is pinned against outputs of
the *real* reference code: ``tests/golden/gen_reference_golden.py`` imports
``/root/reference/tIGAr/BSplines.py`` under stub ``dolfin``/``petsc4py``
modules, compiles the reference's own embedded C++ ``basisFuncsInner`` and
writes ``tests/golden/bspline_reference.npz``.  The FE-assembly / PtAP / solve
layer delegates to FEniCS/PETSc in the reference (not installed, not
buildable here) and is therefore "parity unpinned" against the reference
itself; it is pinned by the known-answer tests of SURVEY.md 8c (independent
direct-IGA Galerkin assembly, partition of unity, manufactured solutions,
analytic eigenvalues).
"""
