"""Time the stages of the Kronecker PtAP separately (CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tigar_b200.engine import TensorPatch, WinMatrix
from tigar_b200 import dev
from tIGAr.BSplines import uniformKnots
nel = int(sys.argv[1]) if len(sys.argv) > 1 else 48
patch = TensorPatch([3] * 3, [uniformKnots(3, 0.0, 1.0, nel)] * 3)
A = WinMatrix(patch.window("A")); A.vals.fill_(1.0)
for rep in range(3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record(); C = patch.ptap_kron(A); ev[1].record(); torch.cuda.synchronize()
    print("total ms", ev[0].elapsed_time(ev[1]))
for k in ("A", "P", "K0", "K1", "K2", "C"):
    w = patch._win[k] if k in patch._win else patch.window(k)
    print(k, "rows", w.nrows, "nnz", w.nnz, "GB", 8e-9 * w.nnz)
