"""numpy restatement of the in-kernel counter-based fire mask.  TEST INFRASTRUCTURE ONLY.

The reference draws its fire mask from torch's global generator
(ExtraChannels/models/dynca.py:121, EncoderConditioning/nca.py:165-174); the production CUDA
path instead uses Philox4x32-10 keyed on (seed, t, b, y, x) so the backward pass can regenerate
it (SURVEY.md §8b RNG row).  This file restates that generator (Salmon et al., SC'11, the
published Philox4x32-10 constants) so tests can check the kernel's mask bit-for-bit.

Counter layout (must match csrc/nca_common.cuh: nca_philox_fire):
    ctr = (quad, b, t, 0x4E434131)  with quad = (y*W + x) >> 2 ; the word used is r[(y*W+x) & 3]
    key = (seed & 0xffffffff, seed >> 32)
Fire rules:
    DyNCA  (floor(u+rate))     : fire = r >= ceil((1-rate) * 2^32)
    ENC    (u < rate)          : fire = r <  ceil(rate * 2^32)
"""
import math

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
STREAM = 0x4E434131
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n1 = lo1
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        n3 = lo0
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def fire_threshold(rate: float, enc: bool) -> int:
    if enc:
        return min(max(int(math.ceil(rate * 4294967296.0)), 0), 1 << 32)
    return min(max(int(math.ceil((1.0 - rate) * 4294967296.0)), 0), 1 << 32)


def fire_mask(seed: int, t0: int, T: int, B: int, H: int, W: int, rate: float, enc: bool = False):
    """float32 [T,B,1,H,W] mask identical to the kernel's."""
    p = np.arange(H * W, dtype=np.uint64)
    quad, lane = p >> np.uint64(2), (p & np.uint64(3)).astype(np.int64)
    thr = np.uint64(fire_threshold(rate, enc))
    out = np.zeros((T, B, 1, H, W), np.float32)
    for ti in range(T):
        for b in range(B):
            r = philox4x32_10(quad, np.full_like(quad, b), np.full_like(quad, t0 + ti),
                              np.full_like(quad, STREAM), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
            r = np.stack(r, axis=1)[np.arange(H * W), lane]
            f = (r < thr) if enc else (r >= thr)
            out[ti, b, 0] = f.reshape(H, W).astype(np.float32)
    return out
