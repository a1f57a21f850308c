import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from helpers import DYNCA_CASES, load_case, rel_err
from test_dynca_gpu import build_model
from test_dynca_bf16_gpu import _grads
DEV = "cuda"
for name in DYNCA_CASES:
    t, m = load_case(name)
    mb, mf = build_model(m, t, precision="bf16"), build_model(m, t, precision="fp32")
    x0, masks = t["x0"].to(DEV), t["masks"].to(DEV)
    kwargs = dict(cond_img=t["cond_img"].to(DEV) if "cond_img" in t else None) if m["flavour"] == "cd" else {}
    for T in (1, min(m["T"], 6)):
        taps = [tp for tp in m["taps"] if tp <= T]
        coefs = [t["coef_final"].to(DEV)] + [t[f"coef_tap{tp}"].to(DEV) for tp in taps]
        gb = _grads(mb, x0, T, masks[:T], taps, coefs, **kwargs)
        gf = _grads(mf, x0, T, masks[:T], taps, coefs, **kwargs)
        print(name, "T=%d" % T, " ".join("%s=%.2e" % (n, rel_err(a, b)) for a, b, n in zip(gb, gf, ("x0", "w1", "b1", "w2", "b2"))),
              "rms:", " ".join("%.1e" % float((a - b).norm() / b.norm()) for a, b in zip(gb, gf)))
