"""Per-case kernel variants and errors of the tcgen05 path against the rounding-point emulator (usage: emu_errors.py [T] [case ...])"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import *
import test_dynca_bf16_gpu as tb
from nca_b200 import functional as Fn, _lib
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cases = sys.argv[2:] or DYNCA_CASES
for name in cases:
    t, m = load_case(name)
    kind, cc = {"cpe": (_lib.NCA_COND_CPE, 2), "edges": (_lib.NCA_COND_TENSOR, 3), None: (_lib.NCA_COND_NONE, 0)}[m["cond"]]
    cfg = Fn.DyncaConfig(m["C"], m["fc"], m["pad"], m["scales"], kind, cc, precision="bf16")
    fv, bv = (Fn.dynca_kernel_variant(cfg, m["B"], m["H"], m["W"], backward=bw) for bw in (False, True))
    es, gmax, grms, _ = tb._emu_errors(name, T)
    print(name, (m["B"], m["C"], m["H"], m["W"], m["fc"]), m["scales"], m["pad"], "variants", fv, bv, "state %.2e" % es,
          {k: "%.1e" % v for k, v in gmax.items()}, flush=True)
