"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules on CPU.

Run in the build container only (needs /root/reference):  python oracle/make_golden.py
The reference modules are imported by file path (``models/__init__.py`` pulls in gdown /
progressbar, which are absent).  The reference draws its fire mask from the
global torch generator inside ``forward``; we re-seed and pre-draw the same sequence so the
mask can be stored and SUPPLIED to the oracle / CUDA path.

Nothing from /root/reference is copied into the repo: only numeric inputs/outputs are saved.
"""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
CPU = torch.device("cpu")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def ref_modules():
    ec = _load("ref_ec_dynca", f"{REF}/ExtraChannels/models/dynca.py")
    cd = _load("ref_cd_dynca", f"{REF}/ConditioneDyNCA/models/dynca.py")
    sys.path.insert(0, f"{REF}/EncoderConditioning")
    enc = _load("ref_enc_nca", f"{REF}/EncoderConditioning/nca.py")
    return ec, cd, enc


def draw_masks(seed, T, B, H, W, rate):
    torch.manual_seed(seed)
    return torch.stack([(torch.rand(B, 1, H, W) + rate).floor() for _ in range(T)])


def dynca_case(name, mod, flavour, C, fc, H, W, B, T, pad, scales, cond, edge_transform="tanh",
               taps=(), weights=None, seed=0, x0_scale=1.0):
    torch.manual_seed(1000 + seed)
    kw = dict(c_in=C, c_out=3, fc_dim=fc, padding_mode=pad, perception_scales=list(scales), device=CPU)
    if flavour == "ec":
        model = mod.DyNCA(pos_emb=("CPE" if cond == "cpe" else None), **kw)
    else:
        model = mod.DyNCA(conditioning={"cpe": "pos_emb", "edges": "edges", None: None}[cond],
                          edge_transform=edge_transform, **kw)
    if weights is not None:
        with torch.no_grad():
            model.w1.weight.copy_(weights["w1"].reshape(model.w1.weight.shape))
            model.w1.bias.copy_(weights["b1"])
            model.w2.weight.copy_(weights["w2"].reshape(model.w2.weight.shape))
            model.w2.bias.copy_(weights["b2"])
    else:
        with torch.no_grad():  # make biases / w2 non-degenerate so every gradient path is exercised
            model.w2.bias.normal_(0, 0.02)
            model.w2.weight.mul_(4.0)
    x0 = (x0_scale * (torch.rand(B, C, H, W) - 0.5)).requires_grad_(True)
    cond_img = (torch.rand(B, 1, H, W) * 2 - 1) if cond == "edges" else None
    rate = 0.5
    masks = draw_masks(77 + seed, T, B, H, W, rate)
    torch.manual_seed(77 + seed)
    kwargs = dict(cond_img=cond_img) if flavour == "cd" else {}
    state, rgb, mids = model.forward_nsteps(x0, T, update_rate=rate, return_middle_feature=True, **kwargs)
    # loss touches the final state and rgb taps at chosen steps (fit_video_motion.py:230-235 pattern)
    gen = torch.Generator().manual_seed(5 + seed)
    cf = torch.randn(state.shape, generator=gen)
    loss = (state * cf).sum()
    tap_coefs = {}
    for t in taps:
        ct = torch.randn(mids[t - 1].shape, generator=gen)
        tap_coefs[t] = ct
        loss = loss + (mids[t - 1] * ct).sum()
    loss.backward()
    # one extra reference-only probe: single-step perception (return_perception=True)
    torch.manual_seed(3)
    with torch.no_grad():
        _, _, percept = model(x0.detach(), update_rate=rate, return_perception=True, **kwargs)
    d = dict(
        x0=x0.detach().numpy(), masks=masks.numpy(), final=state.detach().numpy(),
        rgb_last=rgb.detach().numpy(), percept=percept.numpy(),
        w1=model.w1.weight.detach().reshape(fc, -1).numpy(), b1=model.w1.bias.detach().numpy(),
        w2=model.w2.weight.detach().reshape(C, fc).numpy(), b2=model.w2.bias.detach().numpy(),
        g_w1=model.w1.weight.grad.reshape(fc, -1).numpy(), g_b1=model.w1.bias.grad.numpy(),
        g_w2=model.w2.weight.grad.reshape(C, fc).numpy(), g_b2=model.w2.bias.grad.numpy(),
        g_x0=x0.grad.numpy(), coef_final=cf.numpy(),
        meta=json.dumps(dict(flavour=flavour, C=C, fc=fc, H=H, W=W, B=B, T=T, pad=pad, scales=list(scales),
                             cond=cond, edge_transform=edge_transform, taps=list(taps), rate=rate)),
    )
    for t, ct in tap_coefs.items():
        d[f"coef_tap{t}"] = ct.numpy()
        d[f"rgb_tap{t}"] = mids[t - 1].detach().numpy()
    if cond_img is not None:
        d["cond_img"] = cond_img.numpy()
        with torch.no_grad():
            d["cond_mat"] = model.cond_layer(cond_img).numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "loss", float(loss), "max|state|", float(state.abs().max()))


def load_webgl_weights(path):
    """docs/data/**.json: w = (data - center) * scale ; rows = [w.T ; bias]."""
    d = json.load(open(path))
    out = []
    for L in d["layers"]:
        a = (np.asarray(L["data_flatten"], np.float32) - np.float32(L["center"])) * np.float32(L["scale"])
        out.append(a.reshape(L["shape"]))
    l1, l2 = out
    return dict(w1=torch.tensor(l1[:-1].T.copy()), b1=torch.tensor(l1[-1].copy()),
                w2=torch.tensor(l2[:-1].T.copy()), b2=torch.tensor(l2[-1].copy())), d.get("n_perception_scales", 1)


def enc_case(name, enc, H, W, B, T, seed=0, living=(0.4, -0.1)):
    torch.manual_seed(2000 + seed)
    nca = enc.ConditionedNCA(target_shape=(3, H, W), num_hidden_channels=16, living_channel_dim=3)
    C = nca.num_channels
    with torch.no_grad():
        nca.update_net.out[0].bias.normal_(0, 0.05)
        nca.update_net.out[2].bias.normal_(0, 0.05)
    x0 = nca.generate_seed(B, size=H).clone()
    x0 = x0 + 0.3 * torch.randn(B, C, H, W)
    x0[:, 3] = torch.rand(B, H, W) * living[0] + living[1]   # living channel straddles the 0.1 threshold
    x0.requires_grad_(True)
    goal = torch.rand(B, 3, H, W)
    with torch.no_grad():
        ge = nca.encoder(goal)
    ge = torch.nn.functional.pad(ge, (0, 0, 0, 0, C - 16, 0)).detach().requires_grad_(True)
    rate = nca.cell_fire_rate
    torch.manual_seed(88 + seed)
    fires = torch.stack([(torch.rand(B, 1, H, W) < rate).float() for _ in range(T)])
    torch.manual_seed(88 + seed)
    x = x0
    for _ in range(T):
        x, _ = nca.forward((x, ge))
    cf = torch.randn(x.shape, generator=torch.Generator().manual_seed(9 + seed))
    loss = (x * cf).sum()
    loss.backward()
    un = nca.update_net.out
    d = dict(
        x0=x0.detach().numpy(), goal_enc=ge.detach().numpy(), fires=fires.numpy(), final=x.detach().numpy(),
        coef_final=cf.numpy(),
        wp=nca.perception_net.weight.detach().numpy(),
        wa=un[0].weight.detach().reshape(64, 3 * C).numpy(), ba=un[0].bias.detach().numpy(),
        wb=un[2].weight.detach().reshape(64, 64).numpy(), bb=un[2].bias.detach().numpy(),
        wc=un[4].weight.detach().reshape(C, 64).numpy(),
        g_wp=nca.perception_net.weight.grad.numpy(),
        g_wa=un[0].weight.grad.reshape(64, 3 * C).numpy(), g_ba=un[0].bias.grad.numpy(),
        g_wb=un[2].weight.grad.reshape(64, 64).numpy(), g_bb=un[2].bias.grad.numpy(),
        g_wc=un[4].weight.grad.reshape(C, 64).numpy(),
        g_x0=x0.grad.numpy(), g_goal=ge.grad.numpy(),
        meta=json.dumps(dict(C=C, H=H, W=W, B=B, T=T, rate=rate, living_dim=3, thr=0.1)),
    )
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "loss", float(loss), "alive frac", float((x.detach().abs().sum(1) > 0).float().mean()))


def main():
    os.makedirs(OUT, exist_ok=True)
    ec, cd, enc = ref_modules()
    # c1-like: EC flavour, CPE, replicate (ctor default), single scale
    dynca_case("ec_c12_cpe_replicate", ec, "ec", 12, 96, 16, 16, 2, 6, "replicate", (0,), "cpe", taps=(1, 4), seed=1)
    # experiments.py default: EC flavour C=13, no pos-emb, circular
    dynca_case("ec_c13_none_circular", ec, "ec", 13, 96, 12, 20, 2, 5, "circular", (0,), None, taps=(2,), seed=2)
    # c2-like: C=16 fc=128 CPE, two scales, circular, non-square
    dynca_case("ec_c16_cpe_ms_circular", ec, "ec", 16, 128, 16, 24, 2, 5, "circular", (0, 1), "cpe", taps=(1, 3), seed=3)
    dynca_case("ec_c16_cpe_ms_replicate", ec, "ec", 16, 128, 12, 16, 1, 4, "replicate", (0, 1), "cpe", seed=4)
    dynca_case("ec_c12_none_ms_constant", ec, "ec", 12, 96, 8, 12, 1, 3, "constant", (0, 1), None, seed=5)
    dynca_case("ec_c12_cpe_reflect", ec, "ec", 12, 96, 10, 14, 1, 3, "reflect", (0,), "cpe", seed=6)
    dynca_case("ec_c12_cpe_ms_reflect", ec, "ec", 12, 96, 8, 12, 1, 3, "reflect", (0, 1), "cpe", seed=12)
    # c3-like: CD flavour, edge conditioning (tanh and identity), circular
    dynca_case("cd_c12_edges_tanh_circular", cd, "cd", 12, 96, 16, 16, 2, 5, "circular", (0,), "edges", "tanh", taps=(2,), seed=7)
    dynca_case("cd_c12_edges_none_replicate", cd, "cd", 12, 96, 12, 16, 2, 4, "replicate", (0,), "edges", "None", seed=8)
    dynca_case("cd_c12_posemb_circular", cd, "cd", 12, 96, 12, 12, 1, 3, "circular", (0,), "cpe", seed=9)
    # trained weights (docs/data): CPE, two perception scales; and an edge-conditioned one
    w, ns = load_webgl_weights(f"{REF}/docs/data/video_models/small/ants.json")
    np.savez_compressed(os.path.join(OUT, "weights_video_small_ants.npz"), **{k: v.numpy() for k, v in w.items()})
    dynca_case("trained_ants_cpe_ms", ec, "ec", 12, 96, 32, 32, 1, 24, "replicate", (0, 1) if ns == 2 else (0,), "cpe",
               weights=w, seed=10, x0_scale=0.2)
    w, ns = load_webgl_weights(f"{REF}/docs/data/vec_field_models/large/starry-night.json")
    np.savez_compressed(os.path.join(OUT, "weights_vecfield_large_starry.npz"), **{k: v.numpy() for k, v in w.items()})
    dynca_case("trained_starry_edges", cd, "cd", 12, 96, 32, 32, 1, 24, "circular", (0,), "edges", "None",
               weights=w, seed=11, x0_scale=0.2)
    # ENC
    enc_case("enc_c20_16x16", enc, 16, 16, 2, 5, seed=1)
    enc_case("enc_c20_12x12", enc, 12, 12, 1, 8, seed=2, living=(0.4, -0.27))


if __name__ == "__main__":
    main()
