import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
import test_parity_baseline_gpu as T
name = "c2_256x256_C16_fc128_cpe_ms_B1"
one_o, fin_o, g_o = T._cpu(name, T.O.dynca_step_aten, True)
for prec in ("fp32", "f16x3"):
    one, fin, g, cfg = T._gpu(name, prec, True)
    d = (g[0] - g_o[0]).abs()
    print(prec, "max", float(d.max()), "ref max", float(g_o[0].abs().max()), "rms rel", T._rel_rms(g[0], g_o[0]))
    idx = (d > 0.1 * d.max()).nonzero()
    print("  n bad", idx.shape[0], "first", idx[:12].tolist())
    ys = sorted(set(idx[:, 2].tolist())); xs = sorted(set(idx[:, 3].tolist()))
    print("  rows", ys[:20], "cols", xs[:20])
    print("  other grads", [T._rel_max(a, b) for a, b in zip(g[1:], g_o[1:])])
