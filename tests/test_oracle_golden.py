"""Pins the oracle's B-spline layer against outputs of the REAL reference
(tests/golden/bspline_reference.npz, written by gen_reference_golden.py, which
runs /root/reference/tIGAr/BSplines.py and its embedded C++).  Bit-exact."""
import numpy as np

from oracle import bsplines as B


def test_uniform_knots(golden):
    for i, a in enumerate(golden["uk_args"]):
        got = B.uniform_knots(int(a[0]), a[1], a[2], int(a[3]), bool(a[4]), int(a[5]))
        assert np.array_equal(np.array(got), golden["uk_%d" % i])


def test_spline1_state(golden):
    for name in golden["s1_names"]:
        pre = "s1_%s_" % name
        s = B.Spline1(int(golden[pre + "p"]), golden[pre + "knots"])
        assert s.nel == int(golden[pre + "nel"])
        assert s.ncp == int(golden[pre + "ncp"])
        assert np.array_equal(s.uniqueKnots, golden[pre + "uniqueKnots"])
        assert np.array_equal(s.multiplicities, golden[pre + "multiplicities"])
        assert np.array_equal(s.ghostKnots, golden[pre + "ghostKnots"])
        assert bool(s.isDiscontinuous()) == bool(golden[pre + "disc"])
        grev = np.array([s.greville(i) for i in range(s.ncp)])
        assert np.array_equal(grev, golden[pre + "greville"])


def test_span_nodes_basis_scalar_and_vectorised(golden):
    for name in golden["s1_names"]:
        pre = "s1_%s_" % name
        s = B.Spline1(int(golden[pre + "p"]), golden[pre + "knots"])
        pts = golden[pre + "pts"]
        spans = np.array([s.getKnotSpan(u) for u in pts])
        assert np.array_equal(spans, golden[pre + "spans"]), name
        nodes = np.array([s.getNodes(u) for u in pts])
        assert np.array_equal(nodes, golden[pre + "nodes"]), name
        vals = np.array([s.basisFuncs(sp, u) for sp, u in zip(spans, pts)])
        assert np.array_equal(vals, golden[pre + "vals"]), name      # bit-exact
        assert np.array_equal(s.spans_vec(pts), golden[pre + "spans"])
        assert np.array_equal(s.basis_vec(s.spans_vec(pts), pts), golden[pre + "vals"])


def test_tensor_product(golden):
    for name in golden["tp_names"]:
        pre = "tp_%s_" % name
        deg = [int(x) for x in golden[pre + "deg"]]
        kv = [golden[pre + "kv%d" % d] for d in range(len(deg))]
        t = B.TensorSpline(deg, kv)
        assert t.getNcp() == int(golden[pre + "ncp"])
        assert t.nel == int(golden[pre + "nel"])
        assert t.getDegree() == int(golden[pre + "degree"])
        assert bool(t.needsDG()) == bool(golden[pre + "needsDG"])
        for x, idx, val in zip(golden[pre + "X"], golden[pre + "idx"], golden[pre + "val"]):
            ne = t.getNodesAndEvals(x)
            assert [a[0] for a in ne] == list(idx)
            assert np.array_equal(np.array([a[1] for a in ne]), val)
        for direction in range(len(deg)):
            for side in (0, 1):
                for nl in (1, 2):
                    assert np.array_equal(
                        np.array(t.getSideDofs(direction, side, nl)),
                        golden[pre + "side_%d_%d_%d" % (direction, side, nl)])
        extra = int(golden[pre + "nsd"]) - len(deg)
        assert np.array_equal(B.explicit_control_net(t, extra), golden[pre + "P"])


def test_index_helpers(golden):
    assert np.array_equal(np.array([B.ij2dof(i, j, 7) for i in range(7) for j in range(3)]),
                          golden["ij2dof"])
    assert np.array_equal(np.array([B.ijk2dof(i, j, k, 5, 4) for i in range(5)
                                    for j in range(4) for k in range(3)]), golden["ijk2dof"])
    assert np.array_equal(np.array([B.dof2ijk(d, 5, 4) for d in range(60)]), golden["dof2ijk"])
    assert np.array_equal(np.array([B.dof2ij(d, 7) for d in range(21)]), golden["dof2ij"])


def test_basis_funcs_inner_with_caller_chosen_index(golden_inner):
    """The oracle's restatement of basisFuncsInner (BSplines.py:73-120) against the reference's
    own compiled routine called with arbitrary i = span+1 (neighbouring spans included):
    bit-exact."""
    g = golden_inner
    n = 0
    for name in g["names"]:
        pre = str(name) + "_"
        p, nG = int(g[pre + "p"]), int(g[pre + "nGhost"])
        for u, i, ders in zip(g[pre + "u"], g[pre + "i"], g[pre + "ders"]):
            got = B.basis_funcs_inner(g[pre + "ghostKnots"], nG, float(u), p, int(i))
            assert np.array_equal(got, ders), (name, u, i)
            n += 1
    assert n > 150
