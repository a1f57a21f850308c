"""Drop-in for ConditioneDyNCA/models/dynca.py: DyNCA (edge / pos-emb conditioned flavour) and EdgeExtractor."""
import torch
import torch.nn as nn

from . import functional as Fn
from ._dynca_base import DyNCABase
from ._lib import NCA_COND_CPE, NCA_COND_NONE, NCA_COND_TENSOR
from .dynca_ec import CPE2D


class DyNCA(DyNCABase):
    """Same constructor, methods and state_dict as the reference's conditioned DyNCA
    (ConditioneDyNCA/models/dynca.py:7-178).  The edge map of ``cond_img`` is constant over a rollout, so
    forward_nsteps computes it once (the reference recomputes it every step, :122-124) and feeds it to the
    step kernel as 3 extra perception inputs."""

    def __init__(self, c_in, c_out, fc_dim=96,
                 padding_mode='replicate',
                 seed_mode='zeros', conditioning='edges',
                 edge_transform='tanh',
                 perception_scales=[0],
                 device=torch.device("cuda:0"), *, precision='fp32'):
        super().__init__()
        self.conditioning = conditioning
        if conditioning == 'pos_emb':
            self.cond_layer = CPE2D()
            c_cond = 2
        elif conditioning == 'edges':
            self.cond_layer = EdgeExtractor(edge_transform).to(device)
            c_cond = 3
        else:
            self.cond_layer = None
            c_cond = 0
        self._init_common(c_in, c_out, fc_dim, padding_mode, seed_mode, perception_scales, device, c_cond, precision)

    def perceive_multiscale(self, x, cond_mat=None):
        return self._perceive_multiscale(x, cond_mat)

    def _cond(self, cond_img):
        if self.conditioning == 'pos_emb':
            return NCA_COND_CPE, 2, None
        if self.conditioning == 'edges':
            with torch.no_grad():
                return NCA_COND_TENSOR, 3, self.cond_layer(cond_img)
        return NCA_COND_NONE, 0, None

    def forward(self, x, update_rate=0.5, return_perception=False, cond_img=None, *, masks=None, seed=None):
        kind, cc, cond = self._cond(cond_img)
        if return_perception:      # a diagnostic output: returned detached (no caller of the reference differentiates it)
            with torch.no_grad():
                cm = self.cond_layer(x) if self.conditioning == 'pos_emb' else cond
                y_percept = self.perceive_multiscale(x, cond_mat=cm)
        x, _ = self._rollout(x, 1, update_rate, kind, cc, cond, masks, seed, False)
        if return_perception:
            return x, self.to_rgb(x), y_percept
        return x, self.to_rgb(x)

    def seed(self, n, size=128):
        return self._seed(n, size, self.c_in)

    def forward_nsteps(self, input_state, step_n, update_rate=0.5, return_middle_feature=False,
                       cond_img=None, *, masks=None, seed=None):
        kind, cc, cond = self._cond(cond_img)
        state, taps = self._rollout(input_state, step_n, update_rate, kind, cc, cond, masks, seed,
                                    return_middle_feature)
        feature = self.to_rgb(state)
        if return_middle_feature:
            return state, feature, taps
        return state, feature


class EdgeExtractor(nn.Module):
    """Zero-padded Sobel-x / Sobel-y / Laplacian of a one-channel image, optional tanh
    (ConditioneDyNCA/models/dynca.py:182-213).  The frozen filter parameters are kept so the state_dict
    matches the reference (cond_layer.{sobel_x,sobel_y,laplacian}.weight); the arithmetic is the
    nca_edge_extract kernel, which hard-codes them."""

    def __init__(self, transform):
        super(EdgeExtractor, self).__init__()
        sobel_x_weight = torch.tensor([[[[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]]]], dtype=torch.float32)
        sobel_y_weight = torch.tensor([[[[-1, -2, -1], [0, 0, 0], [1, 2, 1]]]], dtype=torch.float32)
        laplacian_weight = torch.tensor([[[[1, 2, 1], [2, -12, 2], [1, 2, 1]]]], dtype=torch.float32)
        self.sobel_x = nn.Conv2d(1, 1, kernel_size=3, padding=1, bias=False)
        self.sobel_y = nn.Conv2d(1, 1, kernel_size=3, padding=1, bias=False)
        self.laplacian = nn.Conv2d(1, 1, kernel_size=3, padding=1, bias=False)
        self.sobel_x.weight = nn.Parameter(sobel_x_weight, requires_grad=False)
        self.sobel_y.weight = nn.Parameter(sobel_y_weight, requires_grad=False)
        self.laplacian.weight = nn.Parameter(laplacian_weight, requires_grad=False)
        self.transform = transform
        self.edge_transform = nn.Identity()
        if transform == 'tanh':
            self.edge_transform = nn.Tanh()

    def forward(self, x):
        return Fn.edge_extract(x, self.transform == 'tanh')
