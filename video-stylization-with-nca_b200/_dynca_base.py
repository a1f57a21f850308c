"""Shared implementation of the two DyNCA flavours of the reference (ExtraChannels/models/dynca.py and
ConditioneDyNCA/models/dynca.py).  The arithmetic of perceive / forward / forward_nsteps runs in
libnca_b200.so; this class only mirrors the nn.Module surface (names, arguments, return tuples,
state_dict layout)."""
import numpy as np
import torch

from . import functional as Fn
from ._lib import NCA_COND_CPE, NCA_COND_NONE, NCA_COND_TENSOR, NcaError


class DyNCABase(torch.nn.Module):
    SEED_MODES = ['random', 'center_on', 'zeros']

    def _init_common(self, c_in, c_out, fc_dim, padding_mode, seed_mode, perception_scales, device, c_cond,
                     precision):
        self.c_in = c_in
        self.c_out = c_out
        self.perception_scales = perception_scales
        self.fc_dim = fc_dim
        self.padding_mode = padding_mode
        assert seed_mode in DyNCABase.SEED_MODES
        self.seed_mode = seed_mode
        self.random_seed = 42
        self.device = device
        self.expand = 4
        self.c_cond = c_cond
        self.precision = precision
        # same parameter containers / init as the reference (dynca.py:56-61): state_dict keys w1.*, w2.*
        self.w1 = torch.nn.Conv2d(self.c_in * self.expand + self.c_cond, self.fc_dim, 1, device=self.device)
        torch.nn.init.xavier_normal_(self.w1.weight, gain=0.2)
        self.w2 = torch.nn.Conv2d(self.fc_dim, self.c_in, 1, bias=True, device=self.device)
        torch.nn.init.xavier_normal_(self.w2.weight, gain=0.1)
        torch.nn.init.zeros_(self.w2.bias)
        # the fixed perception filters (dynca.py:63-69); kept as attributes for API compatibility, the
        # kernels hard-code them
        self.sobel_filter_x = torch.FloatTensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]]).to(self.device)
        self.sobel_filter_y = self.sobel_filter_x.T
        self.identity_filter = torch.FloatTensor([[0, 0, 0], [0, 1, 0], [0, 0, 0]]).to(self.device)
        self.laplacian_filter = torch.FloatTensor([[1.0, 2.0, 1.0], [2.0, -12, 2.0], [1.0, 2.0, 1.0]]).to(self.device)

    # -- configuration of the CUDA path ---------------------------------------------------------------
    def _cfg(self, cond_kind, cc, scales=None):
        return Fn.DyncaConfig(self.c_in, self.fc_dim, self.padding_mode,
                              self.perception_scales if scales is None else scales, cond_kind, cc, self.precision)

    def _w(self):
        return self.w1.weight, self.w1.bias, self.w2.weight, self.w2.bias

    # -- reference API --------------------------------------------------------------------------------
    def perceive_torch(self, x, scale=0):
        """[B,C,H,W] -> [B,4C,H,W] (dynca.py:71-96).  scale must be 0 or 1."""
        assert scale in [0, 1, 2, 3, 4, 5]
        if scale not in (0, 1):
            raise NcaError("perceive_torch: only scales 0 and 1 are implemented on the CUDA path")
        p0 = Fn.dynca_perceive(self._cfg(NCA_COND_NONE, 0, [0]), x)
        if scale == 0:
            return p0
        return 2.0 * Fn.dynca_perceive(self._cfg(NCA_COND_NONE, 0, [0, 1]), x) - p0

    def _perceive_multiscale(self, x, cond_mat):
        if cond_mat is None:
            return Fn.dynca_perceive(self._cfg(NCA_COND_NONE, 0), x)
        return Fn.dynca_perceive(self._cfg(NCA_COND_TENSOR, cond_mat.shape[1]), x, cond_mat)

    def to_rgb(self, x):
        return x[:, :self.c_out, ...] * 2.0

    def _seed(self, n, size, channels):
        if isinstance(size, int):
            size_x, size_y = size, size
        else:
            size_x, size_y = size
        # zeros are created on the device (the reference builds them on the host and copies: 100 MB per 1080p seed)
        if self.seed_mode == 'zeros':
            return torch.zeros(n, channels, size_y, size_x, device=self.device)
        elif self.seed_mode == 'center_on':
            sd = torch.zeros(n, channels, size_y, size_x, device=self.device)
            sd[:, :, size_y // 2, size_x // 2] = 1.0
            return sd
        elif self.seed_mode == 'random':
            np.random.seed(self.random_seed)
            torch.manual_seed(self.random_seed)
            torch.cuda.manual_seed_all(self.random_seed)
            sd = (torch.rand(1, channels, size_y, size_x) - 0.5)
        else:
            sd = None
        return torch.cat([sd.clone() for _ in range(n)]).to(self.device)

    def _rollout(self, x, step_n, update_rate, cond_kind, cc, cond, masks, seed, return_taps):
        cfg = self._cfg(cond_kind, cc)
        return Fn.dynca_rollout(cfg, x, *self._w(), step_n, rate=update_rate, cond=cond, masks=masks, seed=seed,
                                c_out=self.c_out, return_taps=return_taps)
