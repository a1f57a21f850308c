"""Drop-in for EncoderConditioning/nca.py (ConditionedNCA, UpdateNet) and encoder.py (ImageEncoder).

The T-step loop of ``grow`` / ``forward`` runs in libnca_b200.so (csrc/enc_f32.cu); the modules only hold the
parameters in the reference's containers (same state_dict keys / shapes) and mirror its methods.  The
ImageEncoder runs once per rollout: forward and weight-gradient pass are one fused kernel each (csrc/enc_encoder.cu) that
write / read the zero-padded goal tensor of the rollout directly; other encoder shapes fall back to the reference's torch ops."""
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as Fn


class ImageEncoder(nn.Module):
    """Frame -> per-pixel embedding (encoder.py:5-64): [sobel_x, sobel_y, laplacian of the gray image, 5x5 gaussian
    blur of every colour channel] -> conv3x3 + ReLU + conv3x3."""

    def __init__(self, embedding_dim, channels):
        super().__init__()
        self.channels = channels

        def fixed(w, k, pad):
            conv = nn.Conv2d(1, 1, kernel_size=k, padding=pad, bias=False)
            conv.weight = nn.Parameter(w.float().view(1, 1, k, k), requires_grad=False)
            return conv

        sx = torch.tensor([[-1., 0., 1.], [-2., 0., 2.], [-1., 0., 1.]])
        self.sobel_x = fixed(sx, 3, 1)
        self.sobel_y = fixed(sx.t().contiguous(), 3, 1)
        r = torch.arange(5, dtype=torch.float64) - 2
        gk = torch.exp(-(r[:, None] ** 2 + r[None, :] ** 2) / 2.0) / (2 * np.pi)
        self.gaussian_blur = fixed(gk / gk.sum(), 5, 2)
        self.laplacian = fixed(torch.tensor([[1., 2., 1.], [2., -12., 2.], [1., 2., 1.]]), 3, 1)
        self.embed = nn.Sequential(
            nn.Conv2d(channels + 3, embedding_dim, kernel_size=3, padding=1),
            nn.ReLU(),
            nn.Conv2d(embedding_dim, embedding_dim, kernel_size=3, padding=1, bias=False),
        )

    def fused_ok(self, x):
        """the one-kernel CUDA path covers the reference's configuration (3 colour channels, 16-wide embedding, float32 on CUDA)"""
        return (x.is_cuda and x.dtype == torch.float32 and not (x.requires_grad and torch.is_grad_enabled())
                and Fn.image_encoder_supported(self.channels, self.embed[2].out_channels) and self.embed[0].out_channels == self.embed[2].out_channels)

    def forward(self, x, goal_channels=None):
        """goal_channels: when given, the embedding comes back already zero-padded to that many channels (leading zeros,
        nca.py:199-203) - the tensor ConditionedNCA.grow feeds to the rollout."""
        if self.fused_ok(x):
            E = self.embed[2].out_channels
            out = Fn.image_encoder(x, self.embed[0].weight, self.embed[0].bias, self.embed[2].weight, E if goal_channels is None else goal_channels)
            return out
        # other shapes / CPU tensors: the reference's own sequence of torch ops (not on the NCA hot path)
        gray = x.mean(dim=1, keepdim=True)
        feats = [self.sobel_x(gray), self.sobel_y(gray), self.laplacian(gray)]
        feats += [self.gaussian_blur(x[:, i:i + 1]) for i in range(self.channels)]
        out = self.embed(torch.cat(feats, dim=1))
        if goal_channels is not None and goal_channels > out.size(1):
            out = F.pad(out, (0, 0, 0, 0, goal_channels - out.size(1), 0))
        return out


class UpdateNet(nn.Module):
    """Parameter container of the update MLP 3C -> 64 -> 64 -> C (nca.py:29-58); evaluated inside the step kernels."""

    def __init__(self, in_channels: int, out_channels: int, zero_bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.out = nn.Sequential(
            nn.Conv2d(in_channels, 64, 1), nn.ReLU(), nn.Conv2d(64, 64, 1), nn.ReLU(),
            nn.Conv2d(64, out_channels, 1, bias=False))
        if zero_bias:
            with torch.no_grad():
                for m in self.out:
                    if isinstance(m, nn.Conv2d) and m.bias is not None:
                        nn.init.zeros_(m.bias)

    def forward(self, x):   # not used by the CUDA path; kept for API compatibility
        return self.out(x)


class ConditionedNCA(nn.Module):
    def __init__(self, encoder: nn.Module = None, target_shape: Tuple[int] = (3, 64, 64), num_hidden_channels=16,
                 use_living_channel: bool = True, living_channel_dim: Optional[int] = None,
                 alpha_living_threshold: float = 0.1, cell_fire_rate: float = 0.5, zero_bias=True, *, precision='fp32'):
        super().__init__()
        self.precision = precision
        self.target_shape = target_shape
        self.num_target_channels = target_shape[0]
        self.image_size = target_shape[-1]
        self.num_hidden_channels = num_hidden_channels
        self.use_living_channel = use_living_channel
        self.living_channel_dim = living_channel_dim if living_channel_dim is not None else self.num_target_channels
        self.num_channels = self.num_target_channels + self.num_hidden_channels + 1
        self.alpha_living_threshold = alpha_living_threshold
        self.cell_fire_rate = cell_fire_rate
        self.zero_bias = zero_bias
        self.perception_net = nn.Conv2d(self.num_channels, self.num_channels * 3, 3, stride=1, padding=1,
                                        groups=self.num_channels, bias=False)
        self.update_net = UpdateNet(self.num_channels * 3, self.num_channels, zero_bias)
        self.encoder = encoder if encoder is not None else ImageEncoder(num_hidden_channels, self.num_target_channels)

    # -- CUDA path plumbing ----------------------------------------------------------------------------
    def _cfg(self):
        return Fn.EncConfig(self.num_channels, self.living_channel_dim if self.use_living_channel else -1,
                            self.alpha_living_threshold, self.cell_fire_rate, precision=self.precision)

    def _w(self):
        o = self.update_net.out
        return (self.perception_net.weight, o[0].weight, o[0].bias, o[2].weight, o[2].bias, o[4].weight)

    def _pad_goal(self, goal_encoding):
        if goal_encoding.size(1) == self.num_hidden_channels:
            goal_encoding = F.pad(goal_encoding, (0, 0, 0, 0, self.num_channels - self.num_hidden_channels, 0))
        return goal_encoding

    # -- reference API ---------------------------------------------------------------------------------
    def encode(self, images):
        return self.encoder(images)

    def generate_seed(self, num_seeds, device: Optional[torch.device] = None, size: Optional[int] = None):
        if device is not None:          # (sic) the reference always builds the seed on the CPU (nca.py:136-137)
            device = torch.device("cpu")
        if size is None:
            size = self.image_size
        seed = torch.zeros(num_seeds, self.num_channels, size, size, device=device)
        seed[:, self.living_channel_dim:, size // 2, size // 2] = 1.0
        return seed

    def alive(self, x):
        if not self.use_living_channel:
            return torch.ones_like(x, dtype=torch.bool, device=x.device)
        d = self.living_channel_dim
        return F.max_pool2d(x[:, d:d + 1], kernel_size=3, stride=1, padding=1) > self.alpha_living_threshold

    def get_stochastic_update_mask(self, x, seed=None):
        B, _, H, W = x.shape
        return Fn.philox_mask(B, H, W, self.cell_fire_rate, Fn.new_seed() if seed is None else seed, 1, enc=True,
                              device=x.device)[0]

    def forward(self, x, *, masks=None, seed=None):
        x, goal_encoding = x[0], x[1]
        x = Fn.enc_rollout(self._cfg(), x, goal_encoding, *self._w(), 1, masks=masks, seed=seed)
        return x, goal_encoding

    def grow(self, x: torch.Tensor, num_steps: int, goal: torch.Tensor, *, masks=None, seed=None) -> torch.Tensor:
        if isinstance(self.encoder, ImageEncoder):      # embedding written straight into the zero-padded goal tensor
            goal_encoding = self.encoder(goal, goal_channels=self.num_channels)
        else:
            goal_encoding = self._pad_goal(self.encoder(goal))
        return Fn.enc_rollout(self._cfg(), x, goal_encoding, *self._w(), num_steps, masks=masks, seed=seed)

    def save(self, path: str):
        torch.save(self.state_dict(), path)

    def load(self, path: str):
        self.load_state_dict(torch.load(path))
