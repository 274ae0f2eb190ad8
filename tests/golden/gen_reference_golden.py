"""
Generates ``tests/golden/bspline_reference.npz`` from the REAL reference code.

Run in the build container only (needs /root/reference):
    python tests/golden/gen_reference_golden.py

How: ``/root/reference/tIGAr/BSplines.py`` is executed *unmodified* as module
``tIGAr.BSplines`` after a stub ``tIGAr.common`` has been planted in
``sys.modules``.  The stub supplies exactly the names BSplines.py pulls from
``from tIGAr.common import *`` (which in turn re-exports dolfin): numpy's
``array``/``zeros``, ``DOLFIN_EPS``, ``near``, ``INDEX_TYPE``, the abstract
base classes (empty), and ``compile_cpp_code`` -- which here compiles the
reference's embedded C++ ``basisFuncsInner`` with g++/pybind11
(oracle/build_ref.py), i.e. the reference's own native code produces the
numbers.  dolfin meshes / PETSc are never touched by the calls made here.

Recorded cases cover: uniformKnots (open / periodic / continuity drop),
BSpline1 state (unique knots, multiplicities, ncp, ghost knots, Greville),
getKnotSpan / getNodes / basisFuncs at interior points, exactly on knots and at
the patch ends, tensor-product getNodesAndEvals in 1/2/3-D (ordering included),
getSideDofs, ExplicitBSplineControlMesh.getHomogeneousCoordinate, getDegree /
needsDG / getPrealloc / nel.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/tIGAr"

from oracle import build_ref  # noqa: E402


def load_reference_bsplines():
    common = types.ModuleType("tIGAr.common")
    DOLFIN_EPS = 3.0e-16

    def near(x, x0, eps=DOLFIN_EPS):
        return (x0 - eps <= x) and (x <= x0 + eps)

    class AbstractScalarBasis(object):
        pass

    class AbstractControlMesh(object):
        pass

    def compile_cpp_code(code):
        build_ref.compile_cxx(code)
        return build_ref.load()

    common.__dict__.update(dict(
        array=np.array, zeros=np.zeros, DOLFIN_EPS=DOLFIN_EPS, near=near,
        INDEX_TYPE="int32", USE_RECT_ELEM_DEFAULT=True, worldcomm=None, selfcomm=None,
        mpirank=0, mpisize=1, AbstractScalarBasis=AbstractScalarBasis,
        AbstractControlMesh=AbstractControlMesh, compile_cpp_code=compile_cpp_code))
    common.__all__ = [k for k in common.__dict__ if not k.startswith("__")]
    pkg = types.ModuleType("tIGAr")
    pkg.__path__ = []
    pkg.common = common
    sys.modules["tIGAr"] = pkg
    sys.modules["tIGAr.common"] = common
    mod = types.ModuleType("tIGAr.BSplines")
    mod.__file__ = os.path.join(REF, "BSplines.py")
    with open(mod.__file__) as f:
        src = f.read()
    sys.modules["tIGAr.BSplines"] = mod
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    return mod


# ----- case definitions (shared with the tests through the npz itself) -------
def knot_cases(R):
    uk = R.uniformKnots
    cases = {
        "p1_n4": (1, uk(1, 0.0, 1.0, 4)),
        "p2_n8": (2, uk(2, 0.0, 1.0, 8)),
        "p3_n10": (3, uk(3, 0.0, 1.0, 10)),
        "p3_n7_shift": (3, uk(3, -1.5, 2.25, 7)),
        "p4_n7": (4, uk(4, -1.0, 1.0, 7)),
        "p5_n3": (5, uk(5, 0.0, 2.0, 3)),
        "p3_n6_cd1": (3, uk(3, 0.0, 1.0, 6, False, 1)),
        "p4_n5_cd2": (4, uk(4, 0.0, 3.0, 5, False, 2)),
        "p2_n6_periodic": (2, uk(2, 0.0, 1.0, 6, True)),
        "p3_n8_periodic": (3, uk(3, 0.0, 2.0, 8, True)),
        "p2_nonuniform": (2, [0, 0, 0, 0.1, 0.35, 0.4, 0.75, 1, 1, 1]),
        "p3_nonuniform_rep": (3, [0, 0, 0, 0, 0.2, 0.5, 0.5, 0.7, 0.9, 0.9, 0.9, 1, 1, 1, 1]),
    }
    return cases


def sample_points(s):
    """Interior points, every unique knot, FE-node-like fractions, patch ends."""
    uk = s.uniqueKnots
    pts = list(uk)
    for e in range(len(uk) - 1):
        h = uk[e + 1] - uk[e]
        for a in range(1, s.p):
            pts.append(uk[e] + (a * h) / s.p)
        pts.append(uk[e] + 0.3819660112501051 * h)
        pts.append(uk[e] + 0.9 * h)
    return np.array(pts, dtype=np.float64)


def main():
    R = load_reference_bsplines()
    out = {}
    names = []
    # uniformKnots
    uk_args = [(1, 0.0, 1.0, 4, False, 0), (2, 0.0, 1.0, 64, False, 0),
               (3, 0.0, 1.0, 10, False, 0), (4, -1.0, 1.0, 20, False, 0),
               (3, 0.25, 4.0, 9, False, 2), (2, 0.0, 1.0, 6, True, 0),
               (5, -2.0, 3.0, 3, False, 4)]
    out["uk_args"] = np.array(uk_args, dtype=np.float64)
    for i, a in enumerate(uk_args):
        out["uk_%d" % i] = np.array(
            R.uniformKnots(int(a[0]), a[1], a[2], int(a[3]), bool(a[4]), int(a[5])))

    for name, (p, knots) in knot_cases(R).items():
        names.append(name)
        s = R.BSpline1(p, knots)
        pre = "s1_%s_" % name
        out[pre + "p"] = np.int64(p)
        out[pre + "knots"] = np.array(knots, dtype=np.float64)
        out[pre + "uniqueKnots"] = np.array(s.uniqueKnots)
        out[pre + "multiplicities"] = np.array(s.multiplicities, dtype=np.int64)
        out[pre + "ncp"] = np.int64(s.getNcp())
        out[pre + "nel"] = np.int64(s.nel)
        out[pre + "ghostKnots"] = np.array(s.ghostKnots)
        out[pre + "disc"] = np.int64(bool(s.isDiscontinuous()))
        out[pre + "greville"] = np.array([s.greville(i) for i in range(s.getNcp())])
        pts = sample_points(s)
        spans = np.array([int(s.getKnotSpan(u)) for u in pts], dtype=np.int64)
        nodes = np.array([s.getNodes(u) for u in pts], dtype=np.int64)
        vals = np.array([s.basisFuncs(int(sp), u) for sp, u in zip(spans, pts)])
        out[pre + "pts"] = pts
        out[pre + "spans"] = spans
        out[pre + "nodes"] = nodes
        out[pre + "vals"] = vals
    out["s1_names"] = np.array(names)

    # tensor-product splines
    uk = R.uniformKnots
    tp_cases = {
        "tp1": ([3], [uk(3, 0.0, 1.0, 5)]),
        "tp2": ([2, 3], [uk(2, 0.0, 1.0, 4), uk(3, -1.0, 1.0, 3)]),
        "tp2_nonuni": ([2, 2], [[0, 0, 0, 0.1, 0.35, 0.4, 0.75, 1, 1, 1], uk(2, 0.0, 2.0, 3)]),
        "tp3": ([2, 1, 3], [uk(2, 0.0, 1.0, 3), uk(1, 0.0, 1.0, 2), uk(3, 0.0, 1.0, 2)]),
        "tp3_cubic": ([3, 3, 3], [uk(3, 0.0, 1.0, 3)] * 3),
    }
    tnames = []
    rng = np.random.RandomState(20261017)
    for name, (deg, kv) in tp_cases.items():
        tnames.append(name)
        b = R.BSpline(deg, kv)
        pre = "tp_%s_" % name
        out[pre + "deg"] = np.array(deg, dtype=np.int64)
        for d, k in enumerate(kv):
            out[pre + "kv%d" % d] = np.array(k, dtype=np.float64)
        out[pre + "ncp"] = np.int64(b.getNcp())
        out[pre + "nel"] = np.int64(b.nel)
        out[pre + "degree"] = np.int64(b.getDegree())
        out[pre + "needsDG"] = np.int64(bool(b.needsDG()))
        out[pre + "prealloc"] = np.int64(b.getPrealloc())
        # points: random interior + points with coordinates on knots
        npts = 24
        X = np.zeros((npts, len(deg)))
        for d, s in enumerate(b.splines):
            lo, hi = s.uniqueKnots[0], s.uniqueKnots[-1]
            X[:, d] = lo + (hi - lo) * rng.rand(npts)
            X[::4, d] = s.uniqueKnots[rng.randint(0, len(s.uniqueKnots), size=len(X[::4]))]
        idx = []
        val = []
        for x in X:
            ne = b.getNodesAndEvals(x)
            idx.append([int(a[0]) for a in ne])
            val.append([float(a[1]) for a in ne])
        out[pre + "X"] = X
        out[pre + "idx"] = np.array(idx, dtype=np.int64)
        out[pre + "val"] = np.array(val)
        for direction in range(len(deg)):
            for side in (0, 1):
                for nl in (1, 2):
                    out[pre + "side_%d_%d_%d" % (direction, side, nl)] = np.array(
                        b.getSideDofs(direction, side, nl), dtype=np.int64)
        cm = R.ExplicitBSplineControlMesh(deg, kv, extraDim=1 if len(deg) == 2 else 0)
        nsd = cm.getNsd()
        P = np.array([[cm.getHomogeneousCoordinate(n, d) for d in range(nsd + 1)]
                      for n in range(b.getNcp())])
        out[pre + "P"] = P
        out[pre + "nsd"] = np.int64(nsd)
    out["tp_names"] = np.array(tnames)
    # index helpers
    out["ij2dof"] = np.array([R.ij2dof(i, j, 7) for i in range(7) for j in range(3)])
    out["ijk2dof"] = np.array([R.ijk2dof(i, j, k, 5, 4) for i in range(5) for j in range(4)
                               for k in range(3)])
    out["dof2ijk"] = np.array([R.dof2ijk(d, 5, 4) for d in range(60)])
    out["dof2ij"] = np.array([R.dof2ij(d, 7) for d in range(21)])
    path = os.path.join(HERE, "bspline_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "with", len(out), "arrays")


if __name__ == "__main__":
    main()
