"""cProfile of the host side of the EXTRACT stage of a bench step (generator, zero DoFs,
TensorPatch with and without a slab partition, control net)."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from tIGAr import EqualOrderSpline, ExtractedSpline
from tigar_b200.engine import TensorPatch
nel = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kv, cm, _ = B.build_inputs(nel, net=False)


def extract(part):
    gen = EqualOrderSpline(1, cm)
    sp = gen.getScalarSpline(0)
    for d in range(3):
        for side in (0, 1):
            gen.addZeroDofs(0, sp.getSideDofs(d, side))
    if part is None:
        return ExtractedSpline(gen, 6, mode="fused")
    tp = gen._tensor_spline(-1)
    return TensorPatch([s.p for s in tp.splines], None, quadDeg=6, splines=tp.splines, part=part)


for part in (None, (3, 8)):
    for _ in range(2):
        extract(part)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    extract(part)
    torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
    print("==== part =", part)
    print(s.getvalue()[:5500])
