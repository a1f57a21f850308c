"""Drop-in for ExtraChannels/models/dynca.py: DyNCA (positional-encoding flavour) and CPE2D."""
import torch
import torch.nn as nn

from ._dynca_base import DyNCABase
from ._lib import NCA_COND_CPE, NCA_COND_NONE


class DyNCA(DyNCABase):
    """Same constructor, methods and state_dict as the reference's DyNCA (dynca.py:7-167).

    Extra keyword-only arguments (not in the reference): ``precision`` ('fp32' | 'bf16' MLP arithmetic) on the
    constructor; ``masks`` (supplied fire masks [T,B,1,H,W]) and ``seed`` (Philox key) on forward /
    forward_nsteps.  Without them the fire mask comes from the in-kernel Philox generator keyed from torch's
    default generator."""

    def __init__(self, c_in, c_out, fc_dim=96,
                 padding_mode='replicate',
                 seed_mode='zeros', pos_emb='CPE',
                 perception_scales=[0],
                 device=torch.device("cuda:0"), *, precision='fp32'):
        super().__init__()
        self.pos_emb = pos_emb
        if pos_emb == 'CPE':
            self.pos_emb_2d = CPE2D()
            c_cond = 2
        else:
            self.pos_emb_2d = None
            c_cond = 0
        self._init_common(c_in, c_out, fc_dim, padding_mode, seed_mode, perception_scales, device, c_cond, precision)

    def perceive_multiscale(self, x, pos_emb_mat=None):
        return self._perceive_multiscale(x, pos_emb_mat)

    def _kind(self):
        return (NCA_COND_CPE, 2) if self.pos_emb_2d else (NCA_COND_NONE, 0)

    def forward(self, x, update_rate=0.5, return_perception=False, *, masks=None, seed=None):
        kind, cc = self._kind()
        if return_perception:      # a diagnostic output: returned detached (no caller of the reference differentiates it)
            with torch.no_grad():
                y_percept = self.perceive_multiscale(x, pos_emb_mat=self.pos_emb_2d(x) if self.pos_emb_2d else None)
        x, _ = self._rollout(x, 1, update_rate, kind, cc, None, masks, seed, False)
        if return_perception:
            return x, self.to_rgb(x), y_percept
        return x, self.to_rgb(x)

    def seed(self, n, size=128):
        # the reference's EC flavour seeds c_in - 1 channels: the caller appends the conditioning channel
        # (dynca.py:140, experiments.py:211)
        return self._seed(n, size, self.c_in - 1)

    def forward_nsteps(self, input_state, step_n, update_rate=0.5, return_middle_feature=False, *, masks=None,
                       seed=None):
        kind, cc = self._kind()
        state, taps = self._rollout(input_state, step_n, update_rate, kind, cc, None, masks, seed,
                                    return_middle_feature)
        feature = self.to_rgb(state)
        if return_middle_feature:
            return state, feature, taps
        return state, feature


class CPE2D(nn.Module):
    """Cartesian positional encoding 2D (dynca.py:170-207).  The step kernels compute these two channels from
    the cell coordinates; this module only serves callers that ask for the tensor."""

    def __init__(self):
        super(CPE2D, self).__init__()
        self.cached_penc = None
        self.last_tensor_shape = None

    def forward(self, tensor):
        if len(tensor.shape) != 4:
            raise RuntimeError("The input tensor has to be 4d!")
        if self.cached_penc is not None and self.last_tensor_shape == tensor.shape:
            return self.cached_penc
        self.cached_penc = None
        batch_size, orig_ch, h, w = tensor.shape
        xs = torch.arange(h, device=tensor.device) / h
        ys = torch.arange(w, device=tensor.device) / w
        xs = 2.0 * (xs - 0.5 + 0.5 / h)
        ys = 2.0 * (ys - 0.5 + 0.5 / w)
        emb = torch.zeros((2, h, w), device=tensor.device)
        emb[:1] = xs[None, :, None]
        emb[1:2] = ys[None, None, :]
        self.cached_penc = emb.unsqueeze(0).repeat(batch_size, 1, 1, 1)
        self.last_tensor_shape = tensor.shape
        return self.cached_penc
