"""CPU restatement of the CALLERS either side of the NCA step.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs; the product path
(video-stylization-with-nca_b200/) never imports this module.

Every function restates the reference lines it cites with plain torch / numpy ops on CPU.  The reference's trainer loops are
inline in `main()` / class methods that need datasets and downloaded loss networks, so they cannot be imported as functions;
what CAN be pinned is pinned in tests/test_callers_cpu.py: the optimizer restatement against `torch.optim.Adam` +
`MultiStepLR` themselves (the objects the reference constructs, experiments.py:157,170-172), the grayscale against the
reference's own `RGBToGrayscale` where /root/reference is present, and the uint8 packing against `VideoWriter.add`'s arithmetic.
"""
import math

import numpy as np
import torch


def pool_gather(pool, idx, extra=None, seed_state=None, inject_n=0):
    """ExtraChannels/experiments.py:203-211:
        input_states = nca_pool[batch_idx]; input_states[:1] = seed_inject[:1]; input_states = cat((input_states, aux_gs), 1)
    EncoderConditioning/conditioned_trainer.py:167: batch[:2] = generate_seed(2)."""
    batch = pool[torch.as_tensor(idx, dtype=torch.long)].clone()
    if inject_n:
        batch[:inject_n] = 0.0 if seed_state is None else seed_state.unsqueeze(0)
    if extra is not None:
        batch = torch.cat((batch, extra), 1)
    return batch


def enc_sample_batch(pool_list, idx, seed_state, living_dim, alive_thr=0.1, inject_n=2):
    """EncoderConditioning/conditioned_trainer.py:100-113 + :167 on a list-backed SamplePool (sample_pool.py:14-33):
        batch = pool[idxs]; None / dead (`torch.sum(nca.alive(batch[i])) == 0`) entries -> seed; stack; batch[:2] = seed
    with alive() = max_pool2d(x[:, d:d+1], 3, 1, 1) > thr (nca.py:152-163)."""
    import torch.nn.functional as F
    batch = [pool_list[int(i)] for i in idx]
    for i in range(len(batch)):
        if batch[i] is None:
            batch[i] = seed_state.clone()
        else:
            alive = F.max_pool2d(batch[i].unsqueeze(0)[:, living_dim:living_dim + 1], kernel_size=3, stride=1, padding=1) > alive_thr
            if torch.sum(alive) == 0.0:
                batch[i] = seed_state.clone()
    batch = torch.stack(batch)
    batch[:inject_n] = seed_state.unsqueeze(0)
    return batch


def pool_scatter(pool, idx, states):
    """experiments.py:259: nca_pool[batch_idx] = nca_states_after[:, :12, :, :]."""
    pool = pool.clone()
    pool[torch.as_tensor(idx, dtype=torch.long)] = states[:, :pool.shape[1]]
    return pool


def normalize_grads(grads, eps=1e-8):
    """experiments.py:252-253: p.grad /= (p.grad.norm() + 1e-8); conditioned_trainer.py:134-136 (eps 1e-10)."""
    return [g / (g.norm() + eps) for g in grads]


def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (experiments.py:157, conditioned_trainer.py:61: defaults, amsgrad off, no weight decay), restated from
    torch/optim/adam.py `_single_tensor_adam`.  Returns new (params, exp_avg, exp_avg_sq); step counts from 1."""
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    out_p, out_m, out_v = [], [], []
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        m = m + (g - m) * (1.0 - beta1)
        v = v * beta2 + (1.0 - beta2) * g * g
        denom = v.sqrt() / bc2_sqrt + eps
        out_p.append(p - step_size * m / denom)
        out_m.append(m)
        out_v.append(v)
    return out_p, out_m, out_v


def multistep_lr(base_lr, milestones, gamma, epoch):
    """torch.optim.lr_scheduler.MultiStepLR (experiments.py:170-172: gamma 0.5): lr after `epoch` scheduler steps."""
    return base_lr * gamma ** sum(1 for m in milestones if m <= epoch)


def overflow_loss(x):
    """ExtraChannels/utils/loss/loss.py:33-36: (nca_state - nca_state.clamp(-1.0, 1.0)).abs().mean()."""
    return (x - x.clamp(-1.0, 1.0)).abs().mean()


def overflow_grad(x, scale=1.0):
    """d overflow_loss / dx = sign(x) [|x| > 1] / numel (what autograd returns for the expression above)."""
    g = torch.zeros_like(x)
    g[x > 1.0] = 1.0
    g[x < -1.0] = -1.0
    return g * (scale / x.numel())


def rgb_to_grayscale(rgb):
    """ExtraChannels/utils/misc/preprocess_texture.py:178-179: torch.mean(rgb, dim=1, keepdim=True)."""
    return torch.mean(rgb, dim=1, keepdim=True)


def frame_to_cond_channel(state, frame_rgb, ch):
    """video_utils.py:72 with the conditioning channel kept inside the state buffer: state[:, ch] = gray(frame)."""
    out = state.clone()
    out[:, ch:ch + 1] = rgb_to_grayscale(frame_rgb)
    return out


def state_to_rgb8(state, scale=2.0):
    """dynca.py:130-131 (to_rgb: x[:, :3] * 2), video_utils.py:78-82, VideoWriter.add :20-27, for every sample of the batch:
        img = z.numpy()[0].transpose(1, 2, 0); img = np.clip(img, -1, 1); img = (img + 1.0) / 2.0; np.uint8(img.clip(0, 1) * 255)."""
    z = (state[:, :3] * scale).detach().cpu().numpy()
    out = []
    for b in range(z.shape[0]):
        img = z[b].transpose(1, 2, 0)
        img = np.clip(img, -1.0, 1.0)
        img = (img + 1.0) / 2.0
        out.append(np.uint8(img.clip(0, 1) * 255))
    return np.stack(out)
