"""GPU parity, layer (i): batched B-spline evaluation and the extraction
operator M -- bit-exact against the reference golden vectors and the oracle."""
import numpy as np
import pytest

from oracle import bsplines as OB
from oracle import extraction as OX

pytestmark = pytest.mark.gpu


def test_eval_batch_bit_exact_vs_reference_golden(golden):
    from tigar_b200.bsplines import BSpline1
    from tigar_b200 import dev
    for name in golden["s1_names"]:
        pre = "s1_%s_" % name
        s = BSpline1(int(golden[pre + "p"]), golden[pre + "knots"])
        span, nodes, vals = s.evalBatch(golden[pre + "pts"])
        assert np.array_equal(dev.to_np(span), golden[pre + "spans"]), name
        assert np.array_equal(dev.to_np(nodes), golden[pre + "nodes"]), name
        assert np.array_equal(dev.to_np(vals), golden[pre + "vals"]), name     # bit-exact
        # scalar API of the reference
        u = float(golden[pre + "pts"][3])
        assert s.getKnotSpan(u) == int(golden[pre + "spans"][3])
        assert s.getNodes(u) == list(golden[pre + "nodes"][3])
        assert np.array_equal(s.basisFuncs(s.getKnotSpan(u), u), golden[pre + "vals"][3])


def test_eval_batch_empty_and_large():
    from tigar_b200.bsplines import BSpline1, uniformKnots
    from tigar_b200 import dev
    s = BSpline1(3, uniformKnots(3, 0.0, 1.0, 257))
    span, nodes, vals = s.evalBatch(np.zeros(0))
    assert span.numel() == 0 and vals.numel() == 0
    rng = np.random.RandomState(3)
    u = rng.rand(200001)
    u[::7] = s.uniqueKnots[rng.randint(0, 258, size=len(u[::7]))]
    o = OB.Spline1(3, OB.uniform_knots(3, 0.0, 1.0, 257))
    sp = o.spans_vec(u)
    span, nodes, vals = s.evalBatch(u)
    assert np.array_equal(dev.to_np(span), sp)
    assert np.array_equal(dev.to_np(vals), o.basis_vec(sp, u))
    assert np.abs(dev.to_np(vals).sum(axis=1) - 1).max() < 1e-14


def test_get_nodes_and_evals_matches_golden(golden):
    from tigar_b200.bsplines import BSpline
    for name in golden["tp_names"]:
        pre = "tp_%s_" % name
        deg = [int(x) for x in golden[pre + "deg"]]
        kv = [golden[pre + "kv%d" % d] for d in range(len(deg))]
        b = BSpline(deg, kv)
        for x, idx, val in list(zip(golden[pre + "X"], golden[pre + "idx"], golden[pre + "val"]))[:6]:
            ne = b.getNodesAndEvals(x)
            assert [a[0] for a in ne] == list(idx)
            assert np.array_equal(np.array([a[1] for a in ne]), val)


CASES = [([2], [8]), ([3, 3], [5, 4]), ([2, 3], [4, 3]), ([4, 4], [3, 5]), ([2, 2, 2], [3, 2, 4]),
         ([3, 3, 3], [3, 3, 2]), ([1, 2], [3, 3])]


@pytest.mark.parametrize("deg,nels", CASES)
def test_M_bit_exact_vs_oracle(deg, nels):
    from tigar_b200.engine import TensorPatch
    from tigar_b200 import dev
    kv = [OB.uniform_knots(p, -0.5, 1.5, n) for p, n in zip(deg, nels)]
    ts = OB.TensorSpline(deg, kv)
    Mo = OX.build_M_kron(ts)
    patch = TensorPatch(deg, kv)
    M = patch.build_M()
    Mg = M.to_scipy(drop_eps=1e-15)
    assert Mg.shape == Mo.shape and Mg.nnz == Mo.nnz
    assert np.array_equal(Mg.indptr, Mo.indptr) and np.array_equal(Mg.indices, Mo.indices)
    assert np.array_equal(Mg.data, Mo.data)                      # bit-exact
    # FE node coordinates
    assert np.array_equal(patch.fe_node_coords(), OX.fe_node_coords(ts))
    # M*P (cpFuncs) and M^T b
    rng = np.random.RandomState(0)
    x = rng.rand(Mo.shape[1])
    b = rng.rand(Mo.shape[0])
    y = dev.to_np(M.matvec(dev.from_np(x)))
    assert np.abs(y - Mo @ x).max() < 1e-14
    z = dev.to_np(patch.mt_vec(M, dev.from_np(b)))
    assert np.abs(z - Mo.T @ b).max() < 1e-13


def test_M_nonuniform_knots():
    from tigar_b200.engine import TensorPatch
    kv = [[0, 0, 0, 0.1, 0.35, 0.4, 0.75, 1, 1, 1], [0, 0, 0, 0, 0.2, 0.5, 0.6, 1, 1, 1, 1]]
    ts = OB.TensorSpline([2, 3], kv)
    Mo = OX.build_M_loop(ts)                      # literal reference loop
    Mg = TensorPatch([2, 3], kv).build_M().to_scipy(drop_eps=1e-15)
    assert np.array_equal(Mg.indices, Mo.indices) and np.array_equal(Mg.data, Mo.data)


def test_unsupported_inputs_fail_loudly():
    from tigar_b200.engine import TensorPatch
    with pytest.raises(NotImplementedError):
        TensorPatch([2], [OB.uniform_knots(2, 0.0, 1.0, 6, True)])            # periodic
    # a C^-1 knot: the spline is supported on the element-fused path (no FE space), but the
    # csr path's operands (A_FE, M on a DG space) are not built
    tp = TensorPatch([2], [[0, 0, 0, 0.5, 0.5, 0.5, 1, 1, 1]])
    assert tp.window("C").nnz == 2 * 9
    with pytest.raises(NotImplementedError):
        tp.window("A")


@pytest.mark.gpu
def test_basis_funcs_inner_is_the_reference_native_routine():
    """tIGAr.BSplines.basisFuncsInner(ghostKnots,nGhost,u,pl,i,ndu,left,right,ders)
    (BSplines.py:73-120, 135-145): the caller's index i = span+1, result in ``ders`` --
    bit-exact with the oracle recurrence, also for a span the point does not lie in."""
    from tIGAr.BSplines import basisFuncsInner, BSpline1, uniformKnots
    from oracle import bsplines as OB
    for p in (1, 2, 3, 5):
        kv = uniformKnots(p, -1.0, 2.0, 7)
        ours, ref = BSpline1(p, kv), OB.Spline1(p, kv)
        rng = np.random.default_rng(p)
        for u in list(rng.uniform(-1.0, 2.0, 6)) + [-1.0, 2.0, ref.uniqueKnots[3]]:
            span = ref.getKnotSpan(u)
            for sp in (span, min(span + 1, ref.getKnotSpan(2.0))):
                ders = np.zeros(p + 1)
                basisFuncsInner(ref.ghostKnots, ref.nGhost, u, p, sp + 1, np.zeros((p + 1, p + 1)),
                                np.zeros(p + 1), np.zeros(p + 1), ders)
                assert np.array_equal(ders, ref.basisFuncs(sp, u))
                assert np.array_equal(ours.basisFuncs(sp, u), ders)


def test_basis_funcs_inner_equals_the_reference_native_code(golden_inner):
    """tg_basis_funcs_inner (scalar wrapper and batched) against vectors produced by the
    reference's OWN compiled C++ basisFuncsInner for caller-chosen indices: bit-exact."""
    from tIGAr.BSplines import basisFuncsInner
    g = golden_inner
    for name in g["names"]:
        pre = str(name) + "_"
        p, nG = int(g[pre + "p"]), int(g[pre + "nGhost"])
        gk, us, ii, ref = g[pre + "ghostKnots"], g[pre + "u"], g[pre + "i"], g[pre + "ders"]
        out = np.zeros_like(ref)
        basisFuncsInner(gk, nG, us, p, ii, None, None, None, out)          # batched
        assert np.array_equal(out, ref), name
        d1 = np.zeros(p + 1)
        basisFuncsInner(gk, nG, float(us[0]), p, int(ii[0]), np.zeros((p + 1, p + 1)),
                        np.zeros(p + 1), np.zeros(p + 1), d1)              # reference call form
        assert np.array_equal(d1, ref[0]), name
