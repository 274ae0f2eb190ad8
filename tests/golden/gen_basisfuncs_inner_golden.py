"""
Generates ``tests/golden/basisfuncs_inner_reference.npz``: the reference's native routine
``basisFuncsInner(ghostKnots,nGhost,u,pl,i,ndu,left,right,ders)`` (BSplines.py:73-120, 135-145),
compiled from the reference's OWN embedded C++ source (oracle/build_ref.py, as
gen_reference_golden.py does) and called through the reference's own Python wrapper with
caller-chosen indices ``i`` -- including spans the point does not lie in, which the batched
``getKnotSpan``-based golden vectors never exercise.

    python tests/golden/gen_basisfuncs_inner_golden.py      # needs /root/reference (read-only)

TEST INFRASTRUCTURE: the .npz is the committed fixture; nothing reads /root/reference at test time.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from gen_reference_golden import load_reference_bsplines  # noqa: E402


def main():
    R = load_reference_bsplines()
    rng = np.random.default_rng(20261017)
    out, names = {}, []
    cases = {
        "p1_uniform": (1, R.uniformKnots(1, 0.0, 1.0, 6)),
        "p2_uniform": (2, R.uniformKnots(2, -1.0, 2.0, 7)),
        "p3_uniform": (3, R.uniformKnots(3, 0.0, 1.0, 9)),
        "p3_nonuniform": (3, [0, 0, 0, 0, 0.1, 0.35, 0.4, 0.8, 1, 1, 1, 1]),
        "p4_drop": (4, R.uniformKnots(4, 0.0, 2.0, 5, False, 1)),
        "p5_uniform": (5, R.uniformKnots(5, -2.0, 3.0, 4)),
    }
    for name, (p, knots) in cases.items():
        s = R.BSpline1(p, knots)
        lo, hi = float(s.knots[0]), float(s.knots[-1])
        us = list(rng.uniform(lo, hi, 10)) + [lo, hi] + [float(k) for k in s.uniqueKnots[1:-1][:3]]
        U, I, D = [], [], []
        smin = int(s.multiplicities[0]) - 1
        smax = len(s.knots) - 1 - (int(s.multiplicities[-1]) - 1) - 1
        for u in us:
            span = s.getKnotSpan(u)
            for sp in sorted(set([span, max(smin, span - 1), min(smax, span + 1)])):
                ndu = np.zeros((p + 1, p + 1))
                left, right, ders = np.zeros(p + 1), np.zeros(p + 1), np.zeros(p + 1)
                with np.errstate(all="ignore"):
                    R.basisFuncsInner(s.ghostKnots, s.nGhost, u, p, sp + 1, ndu, left, right, ders)
                if not np.all(np.isfinite(ders)):
                    continue                      # 0/0 at a repeated knot: not a defined output
                U.append(u)
                I.append(sp + 1)
                D.append(ders.copy())
        names.append(name)
        pre = name + "_"
        out[pre + "p"] = np.int64(p)
        out[pre + "ghostKnots"] = np.array(s.ghostKnots, dtype=np.float64)
        out[pre + "nGhost"] = np.int64(s.nGhost)
        out[pre + "u"] = np.array(U)
        out[pre + "i"] = np.array(I, dtype=np.int32)
        out[pre + "ders"] = np.array(D)
    out["names"] = np.array(names)
    path = os.path.join(HERE, "basisfuncs_inner_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {n: out[n + "_u"].shape[0] for n in names})


if __name__ == "__main__":
    main()
