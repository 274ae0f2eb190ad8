"""cProfile of the host side of one bench step (where does the Python time go)."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
nel = int(sys.argv[1]) if len(sys.argv) > 1 else 128
kv, cm, pinned = B.build_inputs(nel, net=False)
cols = None
for _ in range(2):
    B.one_step(kv, cm, cols, "fused", 1e-10, False)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
out = B.one_step(kv, cm, cols, "fused", 1e-10, False)
pr.disable()
torch.cuda.synchronize()
ev = out[2]
print("stage ms:", [round(ev[i].elapsed_time(ev[i + 1]), 2) for i in range(3)])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
