"""Golden vectors for the ImageEncoder (EncoderConditioning/encoder.py), made from the UNMODIFIED reference module.
Run in the build container (needs /root/reference):  python oracle/make_golden_encoder.py  -> tests/golden/encoder.npz"""
import importlib.util
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("ref_encoder", "/root/reference/EncoderConditioning/encoder.py")
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)

torch.manual_seed(0)
enc = mod.ImageEncoder(16, 3)
g = torch.Generator().manual_seed(1)
x = torch.rand(2, 3, 20, 28, generator=g)
coef = torch.randn(2, 16, 20, 28, generator=g)
out = enc(x)
(out * coef).sum().backward()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "encoder.npz"),
                    x=x.numpy(), coef=coef.numpy(), out=out.detach().numpy(),
                    w1=enc.embed[0].weight.detach().numpy(), b1=enc.embed[0].bias.detach().numpy(), w2=enc.embed[2].weight.detach().numpy(),
                    g_w1=enc.embed[0].weight.grad.numpy(), g_b1=enc.embed[0].bias.grad.numpy(), g_w2=enc.embed[2].weight.grad.numpy(),
                    gauss=enc.gaussian_blur.weight.detach().numpy())
print("wrote encoder.npz", out.shape)
