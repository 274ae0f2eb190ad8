"""Time the system-matrix SpMV variants on one assembled matrix (CUDA events)."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nel = int(sys.argv[1]) if len(sys.argv) > 1 else 128
if len(sys.argv) > 2:      # child: one variant
    import torch, ctypes as C
    from tigar_b200.engine import TensorPatch, WinMatrix
    from tigar_b200 import dev
    from tigar_b200._lib import lib, check
    from tIGAr.BSplines import uniformKnots
    patch = TensorPatch([3] * 3, [uniformKnots(3, 0.0, 1.0, nel)] * 3)
    W = patch.window("C")
    A = WinMatrix(W); A.vals.copy_(torch.rand(A.vals.numel(), dtype=torch.float64, device='cuda'))
    x = dev.from_np(__import__("numpy").random.rand(W.ncols)); y = dev.empty(W.nrows)
    scratch = dev.empty(lib.tg_cg_scratch_len()); out = dev.zeros(1)
    ts = []
    for rep in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.tg_win_spmv_dot(W.ref(), dev.ptr(A.vals), dev.ptr(x), 0, dev.ptr(y), dev.ptr(scratch), dev.ptr(out), dev.stream()))
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts[2:])[len(ts[2:]) // 2]
    b = 8 * W.nnz + 16 * W.nrows
    print("%-28s %8.3f ms  %7.1f GB/s  frac %.3f" % (sys.argv[2], ms, b / ms / 1e6, b / ms / 1e6 / 6553.6))
    sys.exit(0)
for name, env in [("rowmajor + cp.async row prefetch", {"TIGAR_B200_SPMV_PF": "1"}),
                  ("tma rows + smem x tiles", {"TIGAR_B200_TMA_SPMV": "1"}),
                  ("rowmajor U=4 (default)", {"TIGAR_B200_SPMV_U": "4"}),
                  ("rowmajor U=4 L1::no_allocate", {"TIGAR_B200_SPMV_LD": "1"}),
                  ("rowmajor U=4 no_alloc+L2::256B", {"TIGAR_B200_SPMV_LD": "2"}),
                  ("sell G=4", {"TIGAR_B200_LAYOUT": "1", "TIGAR_B200_SELL_G": "4"})]:
    e = dict(os.environ); e.update(env)
    subprocess.run([sys.executable, __file__, str(nel), name], env=e)
