"""Fused ImageEncoder (csrc/enc_encoder.cu; reference EncoderConditioning/encoder.py:5-64) through the C ABI: forward and weight
gradients against the golden vectors made from the unmodified reference, against the CPU oracle at the config-4 shape, and through
ConditionedNCA.grow (the encoder is trained through the rollout: conditioned_trainer.py:61, 118-137).  fp32: 1e-5 / 1e-4."""
import os

import numpy as np
import pytest
import torch

import nca_b200
from nca_b200 import functional as Fn
from oracle import nca_oracle as O
from helpers import rel_err, GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _golden():
    d = np.load(os.path.join(GOLDEN, "encoder.npz"))
    return {k: torch.from_numpy(d[k]) for k in d.files}


def test_fused_encoder_matches_reference_golden():
    t = _golden()
    ps = [t[k].clone().to(DEV).requires_grad_(True) for k in ("w1", "b1", "w2")]
    out = Fn.image_encoder(t["x"].to(DEV), *ps, 16)
    assert rel_err(out.detach().cpu(), t["out"]) < 1e-5
    (out * t["coef"].to(DEV)).sum().backward()
    for p, k in zip(ps, ("w1", "b1", "w2")):
        assert rel_err(p.grad.cpu(), t["g_" + k]) < 1e-4, k
    # zero-padded layout of the rollout's goal tensor: leading channels zero, embedding last
    padded = Fn.image_encoder(t["x"].to(DEV), *[p.detach() for p in ps], 20)
    assert padded.shape[1] == 20 and float(padded[:, :4].abs().max()) == 0.0
    assert torch.equal(padded[:, 4:], out.detach())


@pytest.mark.parametrize("shape", [(4, 64, 64), (1, 37, 45), (2, 16, 130)])
def test_fused_encoder_against_oracle(shape):
    B, H, W = shape
    g = torch.Generator().manual_seed(3)
    x = torch.rand(B, 3, H, W, generator=g)
    w1 = torch.randn(16, 6, 3, 3, generator=g) * 0.2
    b1 = torch.randn(16, generator=g) * 0.1
    w2 = torch.randn(16, 16, 3, 3, generator=g) * 0.1
    coef = torch.randn(B, 20, H, W, generator=g)
    po = [p.clone().requires_grad_(True) for p in (w1, b1, w2)]
    oo = O.image_encoder(x, *po)
    (oo * coef[:, 4:]).sum().backward()
    pg = [p.clone().to(DEV).requires_grad_(True) for p in (w1, b1, w2)]
    og = Fn.image_encoder(x.to(DEV), *pg, 20)
    assert rel_err(og[:, 4:].detach().cpu(), oo.detach()) < 1e-5
    (og * coef.to(DEV)).sum().backward()
    for a, b, n in zip(pg, po, ("w1", "b1", "w2")):
        assert rel_err(a.grad.cpu(), b.grad) < 1e-4, (n, rel_err(a.grad.cpu(), b.grad))


def test_grow_trains_the_encoder_through_the_fused_kernels():
    """ConditionedNCA.grow: goal image -> fused encoder -> rollout -> loss; the encoder's gradients must equal those of the same
    computation with the CPU oracles end to end"""
    torch.manual_seed(0)
    B, H, T = 2, 16, 3
    nca = nca_b200.ConditionedNCA(target_shape=(3, H, H), num_hidden_channels=16, living_channel_dim=3).to(DEV)
    with torch.no_grad():
        for p in nca.update_net.parameters():
            p.mul_(0.5)
    x0 = torch.zeros(B, 20, H, H)
    x0[:, 3:, H // 2 - 2:H // 2 + 2, H // 2 - 2:H // 2 + 2] = 1.0
    goal = torch.rand(B, 3, H, H)
    fires = (torch.rand(T, B, 1, H, H) < 0.5).float()
    cf = torch.randn(B, 20, H, H)
    out = nca.grow(x0.to(DEV), T, goal.to(DEV), masks=fires.to(DEV))
    (out * cf.to(DEV)).sum().backward()
    e = nca.encoder.embed
    got = [e[0].weight.grad.cpu(), e[0].bias.grad.cpu(), e[2].weight.grad.cpu()]
    # oracle end to end on the CPU
    ew = [e[0].weight.detach().cpu().clone().requires_grad_(True), e[0].bias.detach().cpu().clone().requires_grad_(True),
          e[2].weight.detach().cpu().clone().requires_grad_(True)]
    ws = [w.detach().cpu() for w in nca._w()]
    enc = torch.nn.functional.pad(O.image_encoder(goal, *ew), (0, 0, 0, 0, 4, 0))
    fo = O.enc_rollout(x0, enc, ws[0], ws[1].reshape(64, -1), ws[2], ws[3].reshape(64, 64), ws[4], ws[5].reshape(20, 64), fires)
    (fo * cf).sum().backward()
    assert rel_err(out.detach().cpu(), fo.detach()) < 1e-5
    for a, b, n in zip(got, ew, ("w1", "b1", "w2")):
        assert rel_err(a, b.grad) < 1e-4, (n, rel_err(a, b.grad))
