"""Time the global M^T A M (march passes vs the row-wise Kronecker kernels) with
CUDA events, per pass, and check both agree."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tigar_b200.engine import TensorPatch, WinMatrix
from tigar_b200._lib import lib, check
from tigar_b200 import dev
from tIGAr.BSplines import uniformKnots
nel = int(sys.argv[1]) if len(sys.argv) > 1 else 48
p = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 3
VARIANT = int(os.environ.get("TIGAR_B200_MARCH_VARIANT", "2"))
patch = TensorPatch([p] * dim, [uniformKnots(p, 0.0, 1.0, nel)] * dim)
A = WinMatrix(patch.window("A"))
A.vals.copy_(torch.rand(A.window.nnz, dtype=torch.float64, device="cuda"))
dirs, passes = patch._march_setup()
peak = 6553.6
wA, wM, wC = (patch.window(k) for k in "AMC")
alg = sum(12 * w.nnz + 4 * (w.nrows + 1) for w in (wA, wM, wC))
for rep in range(4):
    X = A.vals
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(passes) + 1)]
    ev[0].record()
    for k, P_ in enumerate(passes):
        D = dirs[P_["d"]]
        Y = dev.empty(P_["wY"].nnz)
        if VARIANT == 2:
            check(lib.tg_ptap_march_w(P_["wX"].ref(), dev.ptr(X), P_["wY"].ref(), dev.ptr(Y),
                                      P_["d"], D["p"], D["GMAX"], dev.ptr(D["irec"]),
                                      dev.ptr(D["Sx"]), dev.ptr(D["jrec"]),
                                      dev.ptr(D["cpad"]), dev.ptr(D["grp"]), dev.ptr(D["slo"]),
                                      dev.ptr(D["shi"]), dev.ptr(P_["tasks"]), P_["ntask"],
                                      dev.ptr(P_["segw"]), P_["nsegw"], P_["maxnodes"],
                                      P_["maxrows"], P_["maxgroups"], P_["maxpieces"], P_["wpc"], dev.stream()))
        else:
          check(lib.tg_ptap_march(P_["wX"].ref(), dev.ptr(X), P_["wY"].ref(), dev.ptr(Y),
                                P_["d"], D["p"], D["KA"], P_["KAmax"], dev.ptr(D["first"]),
                                dev.ptr(D["mrow"]), dev.ptr(D["tabc"]), dev.ptr(D["slo"]),
                                dev.ptr(D["shi"]), dev.ptr(P_["ga"]), P_["nga"],
                                dev.ptr(P_["gb"]), P_["ngb"], dev.ptr(P_["seg"]),
                                P_["nseg"], P_["stage"], P_["outd"], P_["maxlines"],
                                VARIANT, dev.stream()))
        ev[k + 1].record()
        X = Y
    torch.cuda.synchronize()
    if rep == 3:
        tot = ev[0].elapsed_time(ev[-1])
        for k, P_ in enumerate(passes):
            ms = ev[k].elapsed_time(ev[k + 1])
            by = 8 * (P_["wX"].nnz + P_["wY"].nnz)
            print("pass %d: %.3f ms  in %.3f GB out %.3f GB  streamed %.0f GB/s (%.2f of peak) grid %dx%dx%d tasks %d x %d"
                  % (k, ms, 8e-9 * P_["wX"].nnz, 8e-9 * P_["wY"].nnz, by / ms * 1e-6,
                     by / ms * 1e-6 / peak, P_["nga"], P_["ngb"], P_["nseg"], P_["ntask"],
                     P_["nsegw"]))
        print("march total %.3f ms; algorithmic CSR bytes %.3f GB -> %.0f GB/s = %.3f of %.1f"
              % (tot, alg * 1e-9, alg / tot * 1e-6, alg / tot * 1e-6 / peak, peak))
Cm = X
if os.environ.get("PROBE_NO_KRON"):
    sys.exit(0)
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
Ck = patch.ptap_kron(A)
e[0].record(); Ck = patch.ptap_kron(A); e[1].record(); torch.cuda.synchronize()
print("kron (row-wise) total %.3f ms" % e[0].elapsed_time(e[1]))
err = (Cm - Ck.csr_values()).abs().max().item() / Ck.csr_values().abs().max().item()
print("max rel diff march vs kron: %.3e" % err)
