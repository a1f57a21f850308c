"""Does the BPTT read uninitialised workspace / history bytes?  Same forward history, backward with differently pre-filled buffers."""
import os, sys, ctypes as Ct, torch
sys.path.insert(0, '.')
import nca_b200
from nca_b200 import functional as Fn, _lib
DEV = "cuda"
B, C, fc, H, W, T, pad, scales = 2, 16, 128, 64, 96, int(sys.argv[1]), "replicate", (0, 1)
g = torch.Generator().manual_seed(5)
cfg = Fn.DyncaConfig(C, fc, pad, list(scales), _lib.NCA_COND_CPE, 2, precision="bf16")
params = [p.to(DEV) for p in (torch.randn(fc, 4 * C + 2, generator=g) * 0.15, torch.randn(fc, generator=g) * 0.1,
                              torch.randn(C, fc, generator=g) * 0.1, torch.randn(C, generator=g) * 0.02)]
x0 = (torch.rand(B, C, H, W, generator=g) - 0.5).to(DEV)
masks = (torch.rand(T, B, 1, H, W, generator=g) + 0.5).floor().to(DEV)
cf = torch.randn(B, C, H, W, generator=g).to(DEV)
lib = nca_b200.load_library()
hist, coarse, ops = Fn._dynca_forward_raw(cfg, x0, *params, None, masks, 0, T, 0.5, True, want_ops=True)
d = cfg.desc(B, H, W, 0.5, True)
nbytes = lib.nca_dynca_workspace_bytes(Ct.byref(d), 1)
wst = Fn._weights_struct(*params)
stream = Ct.c_void_p(torch.cuda.current_stream().cuda_stream)
def bwd(fill, use_ops=True):
    ws = torch.empty(nbytes, device=DEV, dtype=torch.uint8)
    if fill == "rand": ws.random_(0, 256)
    else: ws.fill_(fill)
    gouts = [torch.full_like(p, float("nan")) for p in params]
    gst = Fn._weights_struct(*gouts)
    gx0 = torch.full_like(x0, float("nan"))
    Fn.check(lib.nca_dynca_backward(Ct.byref(d), Ct.byref(wst), None, masks.data_ptr(), Ct.c_uint64(0), 0, T, hist.data_ptr(), coarse.data_ptr(),
                                    ops.data_ptr() if use_ops else None, cf.data_ptr(), (Ct.c_void_p * 1)(), (Ct.c_int32 * 1)(), 0, 1, 2.0,
                                    gx0.data_ptr(), Ct.byref(gst), ws.data_ptr(), nbytes, stream))
    torch.cuda.synchronize()
    return gx0.cpu(), [t.cpu() for t in gouts]
def rel(a, b): return float((a - b).abs().max() / (b.abs().max() + 1e-30))
ref, refw = bwd(0)
for fill in (0, 0):
    for use_ops in (True, False):
        gx, gw = bwd(fill, use_ops)
        dd = (gx - ref).abs(); idx = (dd > 1e-5 * ref.abs().max()).nonzero()
        if idx.shape[0]: print("   cells", idx.shape[0], "b", sorted(set(idx[:,0].tolist())), "y %d..%d x %d..%d" % (int(idx[:,2].min()), int(idx[:,2].max()), int(idx[:,3].min()), int(idx[:,3].max())))
        print("T", T, "ws fill", fill, "ops" if use_ops else "recompute", "gx0 rel %.1e" % rel(gx, ref), "gw1 rel %.1e" % rel(gw[0], refw[0]), "nan" if torch.isnan(gx).any() else "")
