"""Run a script written against the tIGAr API on the HOST stand-ins of the device layer
(tests/test_scalar_glue_cpu.ScalarFakePatch for engine.TensorPatch, numpy twins of the C-ABI
calls): checks the API layer -- form language, generators, ExtractedSpline drivers -- without
a GPU.  TEST INFRASTRUCTURE ONLY; nothing in the product imports it.

    python tests/run_emulated.py <script.py> [args...]
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import torch
from tigar_b200 import api as A, dev, _lib
import test_multifield_glue_cpu as G
import test_scalar_glue_cpu as SG
fake = G.FakeLib()
dev.device = lambda: torch.device("cpu")
dev.stream = lambda: None
_lib.lib = fake
_lib.check = lambda rc: None
A.lib = fake; A.check = _lib.check; A.WinMatrix = G.FakeWinMatrix
A.TensorPatch = lambda *a, **k: SG.ScalarFakePatch(*a, lib=fake, **k)
script = sys.argv[1]
sys.argv = [script] + sys.argv[2:]
os.environ.setdefault("TIGAR_B200_MODE", "fused")
runpy.run_path(script, run_name="__main__")
