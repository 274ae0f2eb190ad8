"""Run a GENERATED CUDA kernel (tigar_b200.jit.generate) on the CPU: the source is compiled
unchanged with g++ behind a small prelude that maps the CUDA execution model onto host
threads -- one std::thread per CUDA thread of a block, ``__syncthreads()`` = std::barrier,
``__shared__`` = function-static storage, blocks run one after the other.  Numerics of the
form compiler's back end can then be checked without a device (the device run of the same
source is covered by tests/test_gpu_pipeline.py::test_jit_kernel_equals_interpreter).
TEST INFRASTRUCTURE ONLY.
"""
import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np

PRELUDE = r"""
#include <barrier>
#include <cmath>
#include <thread>
#include <vector>
using std::sin; using std::cos; using std::exp; using std::log; using std::sqrt; using std::fabs;
using std::tan; using std::tanh; using std::sinh; using std::cosh; using std::atan; using std::pow;
using std::fmax; using std::fmin;
#define __global__
#define __shared__ static
#define __launch_bounds__(x)
#define __align__(n) __attribute__((aligned(n)))
struct double2 { double x, y; };
struct EmuIdx { int x; };
static thread_local EmuIdx threadIdx, blockIdx;
static std::barrier<>* emu_bar = nullptr;
static inline void __syncthreads() { emu_bar->arrive_and_wait(); }
"""

RUNNER = r"""
extern "C" void emu_run(const QpArgs* A, int nblocks, int nth) {
  for (int b = 0; b < nblocks; b++) {
    std::barrier<> bar(nth);
    emu_bar = &bar;
    std::vector<std::thread> th;
    for (int t = 0; t < nth; t++)
      th.emplace_back([=]() { threadIdx.x = t; blockIdx.x = b; KERNEL(*A);
                              });
    for (auto& x : th) x.join();
  }
}
"""


def build(src, kernel="tigar_qp"):
    """Compile generated CUDA source for the host; returns the loaded library."""
    code = PRELUDE + src + RUNNER.replace("KERNEL", kernel)
    key = hashlib.sha1(code.encode()).hexdigest()[:16]
    d = os.path.join(tempfile.gettempdir(), "tigar_cuda_emu")
    os.makedirs(d, exist_ok=True)
    so = os.path.join(d, key + ".so")
    if not os.path.exists(so):
        cpp = os.path.join(d, key + ".cpp")
        with open(cpp, "w") as f:
            f.write(code)
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-shared", "-fPIC", "-pthread",
                               "-ffp-contract=off", cpp, "-o", so])
    return C.CDLL(so)


def run_qp_kernel(src, nth, tabs, coefs, nout, ncells, cell0=0):
    """tabs: oracle Tab1D per direction (T [nel,nq,nloc,nd], idx, w, x, n); coefs: list of
    global coefficient vectors in kernel order.  Returns out [ncells, nout, nqp]."""
    from tigar_b200.jit import QpArgs
    lib = build(src)
    dim = len(tabs)
    keep = []

    def ptr(a, dt):
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data
    A = QpArgs()
    nqp = 1
    for d in range(3):
        if d < dim:
            tb = tabs[d]
            A.tab[d], A.idx[d] = ptr(tb.T, np.float64), ptr(tb.idx, np.int32)
            A.wq[d], A.xq[d] = ptr(tb.w, np.float64), ptr(tb.x, np.float64)
            A.n[d], A.nel[d] = int(tb.n), int(tb.T.shape[0])
            nqp *= tb.T.shape[1]
        else:
            A.n[d], A.nel[d] = 1, 1
    for i, c in enumerate(coefs):
        A.coef[i] = ptr(c, np.float64)
    out = np.zeros((ncells, nout, nqp))
    keep.append(out)
    A.out = out.ctypes.data
    A.cell0 = cell0
    lib.emu_run.argtypes = [C.POINTER(QpArgs), C.c_int, C.c_int]
    lib.emu_run(C.byref(A), int(ncells), int(nth))
    return out


def run_op_kernel(src, nth, tabs, coefs, y, stride):
    """Fused operator kernel (jit.generate(..., op=...), ``tigar_op``): one launch per colour
    of the cell lattice (``stride`` cells apart per direction), accumulating into ``y``."""
    import itertools
    from tigar_b200.jit import OpArgs
    lib = build(src, "tigar_op")
    dim = len(tabs)
    keep = []

    def ptr(a, dt):
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data
    A = OpArgs()
    for d in range(3):
        if d < dim:
            tb = tabs[d]
            A.tab[d], A.idx[d] = ptr(tb.T, np.float64), ptr(tb.idx, np.int32)
            A.wq[d], A.xq[d] = ptr(tb.w, np.float64), ptr(tb.x, np.float64)
            A.n[d], A.nel[d] = int(tb.n), int(tb.T.shape[0])
        else:
            A.n[d], A.nel[d] = 1, 1
    for i, c in enumerate(coefs):
        A.coef[i] = ptr(c, np.float64)
    assert y.dtype == np.float64 and y.flags.c_contiguous
    A.y = y.ctypes.data
    lib.emu_run.argtypes = [C.POINTER(OpArgs), C.c_int, C.c_int]
    nel = [int(A.nel[d]) for d in range(3)]
    st = list(stride) + [1] * (3 - len(stride))
    launches = 0
    for o in itertools.product(*[range(min(st[d], nel[d])) for d in range(3)]):
        cn = [(nel[d] - o[d] + st[d] - 1) // st[d] for d in range(3)]
        for d in range(3):
            A.co[d], A.cs[d], A.cn[d] = o[d], st[d], cn[d]
        lib.emu_run(C.byref(A), int(np.prod(cn)), int(nth))
        launches += 1
    return launches
