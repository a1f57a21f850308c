import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
from test_dynca_bf16_gpu import _emu_errors, DYNCA_CASES
from nca_b200 import functional as Fn
for T in (1, 4):
    for name in DYNCA_CASES:
        es, gmax, grms, rest = _emu_errors(name, T)
        print(T, name, "state %.1e" % es, {k: "%.1e" % v for k, v in gmax.items()}, {k: "%.1e" % v for k, v in grms.items()})
