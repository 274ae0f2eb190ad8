"""
Recipe that compiles the reference's ONLY native routine, ``basisFuncsInner``
(the C++/pybind11 string embedded at /root/reference/tIGAr/BSplines.py:48-127),
from where it lies in the read-only reference tree into
``oracle/_ref/tigar_ref_basisfuncs*.so``.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Nothing is copied into the
repository: the source string is read from the reference file at build time,
written to a temporary directory, compiled with g++ against pybind11 and a
one-line shim for the (unused) ``<dolfin/common/Array.h>`` include, and only the
resulting shared object is kept (``oracle/_ref/`` is git-ignored).  The
reference does the same thing at import time through
``dolfin.compile_cpp_code`` (BSplines.py:131), which passes the module name as
``-DSIGNATURE=...``.

The rest of the reference's hot path (assemble / PtAP / solve) lives in
FEniCS/PETSc, which are neither installed nor buildable here (they need cmake,
MPI, PETSc, Eigen, Boost, generated code) -- see DESIGN.md.
"""
import os
import re
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_BSPLINES = "/root/reference/tIGAr/BSplines.py"
MODNAME = "tigar_ref_basisfuncs"


def so_path():
    return os.path.join(REF_DIR, MODNAME + sysconfig.get_config_var("EXT_SUFFIX"))


def reference_cxx_string(path=REFERENCE_BSPLINES):
    """Return the C++ source the reference embeds (BSplines.py:48-127)."""
    with open(path, "r") as f:
        text = f.read()
    m = re.search(r'basisFuncsCXXString\s*=\s*"""(.*?)"""', text, re.S)
    if m is None:
        raise RuntimeError("basisFuncsCXXString not found in " + path)
    return m.group(1)


def compile_cxx(code, modname=MODNAME, out=None):
    """What ``dolfin.compile_cpp_code`` does, minus DOLFIN: g++ + pybind11."""
    import pybind11
    out = out or so_path()
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        shim = os.path.join(tmp, "dolfin", "common")
        os.makedirs(shim)
        with open(os.path.join(shim, "Array.h"), "w") as f:
            f.write("// shim: the reference routine no longer uses dolfin::Array\n")
        src = os.path.join(tmp, modname + ".cpp")
        with open(src, "w") as f:
            f.write(code)
        cmd = ["g++", "-O2", "-shared", "-fPIC", "-std=c++17",
               "-DSIGNATURE=" + modname, "-I" + tmp, "-I" + pybind11.get_include(),
               "-I" + sysconfig.get_paths()["include"], src, "-o", out]
        subprocess.check_call(cmd)
    return out


def build(force=False):
    """Build oracle/_ref/<module>.so if the reference tree is present."""
    if os.path.exists(so_path()) and not force:
        return so_path()
    if not os.path.exists(REFERENCE_BSPLINES):
        return None
    return compile_cxx(reference_cxx_string())


REFERENCE_DEMOS = {"poisson.py": "/root/reference/demos/poisson/poisson.py",
                   "biharmonic.py": "/root/reference/demos/biharmonic/biharmonic.py"}
DEMO_DIR = os.path.join(REF_DIR, "demos")


def stage_demos():
    """The north_star's acceptance scripts (demos/poisson/poisson.py, demos/biharmonic/
    biharmonic.py) must run UNMODIFIED on the device, but /root/reference does not exist on the
    GPU box.  Like the compiled reference routine above, byte-identical copies are staged under
    ``oracle/_ref/demos/`` (git-ignored: never in the history; not gpurun-ignored: they travel
    with the snapshot).  tests/test_gpu_reference_demos.py runs them there and checks their
    sha256 against the digests recorded here when the reference tree is present."""
    import hashlib
    import shutil
    out = {}
    for name, src in REFERENCE_DEMOS.items():
        if not os.path.exists(src):
            continue
        os.makedirs(DEMO_DIR, exist_ok=True)
        dst = os.path.join(DEMO_DIR, name)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            out[name] = hashlib.sha256(f.read()).hexdigest()
    return out


def load():
    """Import the compiled reference routine, or None if it was never built."""
    p = so_path()
    if not os.path.exists(p):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location(MODNAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(stage_demos())
