"""Host mirror of the reference's inference loop: `save_video` in
ExtraChannels/utils/misc/video_utils.py:50-82 (extra-channel flavour) and ConditioneDyNCA/utils/misc/video_utils.py:50-82
(`cond_img=` flavour) as a frame stream.

Per target frame the reference runs `torch.cat` (append the grayscale frame), `forward_nsteps`, a channel slice, a blocking
`.cpu().numpy()` and numpy clip / scale / uint8 on the host.  Here the state lives in one persistent two-slot device buffer with
the conditioning channel inside it (`nca_frame_to_cond_channel` rewrites it in place), the T steps go out back to back on one
stream, `nca_state_to_rgb8` packs the frame on the device (1/4 of the bytes cross PCIe/C2C) and the device->host copy of frame i
runs on a second stream under the rollout of frame i+1.
"""
import ctypes as C

import torch

from . import functional as Fn
from ._lib import NCA_COND_CPE, NCA_COND_NONE, NCA_COND_TENSOR, NcaError, check, load_library
from .functional import _need_cuda, _ptr, _stream


def frame_to_cond_channel(state, frame_rgb, ch=-1):
    """state[:, ch] = mean over the 3 channels of frame_rgb (RGBToGrayscale, preprocess_texture.py:178-179), in place."""
    _need_cuda(state, frame_rgb)
    if not (state.is_contiguous() and frame_rgb.is_contiguous()):
        raise NcaError("frame_to_cond_channel needs contiguous tensors")
    B, Cc, H, W = state.shape
    if tuple(frame_rgb.shape) != (B, 3, H, W):
        raise NcaError(f"frame must be [{B},3,{H},{W}], got {tuple(frame_rgb.shape)}")
    ch = ch + Cc if ch < 0 else ch
    with torch.cuda.device(state.device):
        check(load_library().nca_frame_to_cond_channel(B, Cc, H, W, _ptr(frame_rgb), _ptr(state), ch, _stream()))
    return state


def rgb_to_grayscale(frame_rgb):
    """RGBToGrayscale (preprocess_texture.py:178-179): [B,3,H,W] -> [B,1,H,W]."""
    B, _, H, W = frame_rgb.shape
    out = torch.empty(B, 1, H, W, device=frame_rgb.device, dtype=torch.float32)
    return frame_to_cond_channel(out, frame_rgb.contiguous(), 0)


def state_to_rgb8(state, scale=2.0, out=None):
    """uint8 [B,H,W,3] frame of a state: clip(scale * state[:, :3], -1, 1) -> (v + 1) / 2 -> uint8(clip(v, 0, 1) * 255)
    (dynca.py:130-131, video_utils.py:78-82,20-27)."""
    _need_cuda(state)
    if not state.is_contiguous():
        raise NcaError("state_to_rgb8 needs a contiguous state")
    B, Cc, H, W = state.shape
    if out is None:
        out = torch.empty(B, H, W, 3, device=state.device, dtype=torch.uint8)
    with torch.cuda.device(state.device):
        check(load_library().nca_state_to_rgb8(B, Cc, H, W, _ptr(state), float(scale), _ptr(out), _stream()))
    return out


class FrameStylizer:
    """The loop of `save_video` (video_utils.py:65-82) without the file writer.

    model: a drop-in DyNCA (EC flavour: the frame's grayscale is the last state channel; CD flavour with
    conditioning='edges': it is passed as `cond_img`; any other model is run unconditioned).  `push(frame)` advances the
    automaton `step_n` steps on that frame and returns the uint8 image on the device; `run(frames)` does it for a whole clip
    and returns a pinned host array, copies overlapped with compute."""

    def __init__(self, model, size, step_n=8, steps_per_frame=1, batch=1, update_rate=0.5, seed=None, graph=False):
        self.model, self.step_n, self.steps_per_frame, self.rate = model, int(step_n), int(steps_per_frame), float(update_rate)
        # graph=True: the per-frame sequence (conditioning, mask draw, step_n steps, rgb8, step counter) is captured once into a
        # CUDA graph and replayed per frame - one launch from the host instead of step_n + 4.  Small frames (the reference's
        # 256x256) are host bound without it.  The weights are read at replay time (the weight repack is inside the graph).
        self.use_graph, self._graph = bool(graph), None
        H, W = (size, size) if isinstance(size, int) else size
        self.B, self.H, self.W = batch, H, W
        self.dev = model.w1.weight.device
        self.flavour = "cd" if hasattr(model, "conditioning") else "ec"
        self.C = model.c_in
        self.seed = Fn.new_seed() if seed is None else int(seed)
        self.t0 = 0
        self.slots = torch.zeros(2, batch, self.C, H, W, device=self.dev, dtype=torch.float32)
        self.cur = 0
        self.gray = torch.empty(batch, 1, H, W, device=self.dev) if self.flavour == "cd" and model.conditioning == 'edges' else None
        self._host = self._dev_out = self._dev_in = self._side = self._side_up = None
        self._tdev = torch.zeros(1, device=self.dev, dtype=torch.int32)      # Philox step counter of the graph mode
        self.reset()

    def reset(self):
        # h = nca_model.seed(1, size=...) (video_utils.py:66); the EC flavour seeds c_in - 1 channels, the last one is the frame
        h = self.model.seed(self.B, size=(self.W, self.H))
        self.slots.zero_()
        self.slots[0, :, :h.shape[1]].copy_(h)
        self.cur, self.t0 = 0, 0
        self._tdev.zero_()

    @property
    def state(self):
        return self.slots[self.cur]

    def _cfg_cond(self, frame):
        m = self.model
        if self.flavour == "ec":
            frame_to_cond_channel(self.slots[self.cur], frame, self.C - 1)       # video_utils.py:72
            kind, cc = m._kind()
            return m._cfg(kind, cc), None
        if m.conditioning == 'edges':
            frame_to_cond_channel(self.gray, frame, 0)                           # CD video_utils.py:72: cond_img = gray frame
            return m._cfg(NCA_COND_TENSOR, 3), m.cond_layer(self.gray)
        return m._cfg(NCA_COND_CPE if m.conditioning == 'pos_emb' else NCA_COND_NONE, 2 if m.conditioning == 'pos_emb' else 0), None

    def _steps(self, cfg, cond, masks, t0):
        """step_n steps from slot 0 into slot step_n & 1 (masks: supplied [T,B,1,H,W] or None = in-kernel Philox from t0)."""
        lib = load_library()
        d = cfg.desc(self.B, self.H, self.W, self.rate, masks is not None)
        w = [t.detach() for t in self.model._w()]
        with torch.cuda.device(self.dev):
            nbytes = lib.nca_dynca_workspace_bytes(C.byref(d), 0)
            if getattr(self, "_ws", None) is None or self._ws.numel() < nbytes:
                self._ws = torch.empty(max(nbytes, 16), device=self.dev, dtype=torch.uint8)
            wst = Fn._weights_struct(*w)
            check(lib.nca_dynca_forward(C.byref(d), C.byref(wst), _ptr(cond), _ptr(masks), C.c_uint64(self.seed), t0, self.step_n,
                                        0, _ptr(self.slots), None, None, _ptr(self._ws), nbytes, _stream()))

    def _capture(self):
        lib = load_library()
        self._gframe = torch.zeros(self.B, 3, self.H, self.W, device=self.dev)
        self._gout8 = torch.empty(self.B, self.H, self.W, 3, device=self.dev, dtype=torch.uint8)
        self._gmasks = torch.empty(self.step_n, self.B, 1, self.H, self.W, device=self.dev)
        keep = (self.slots.clone(), self._tdev.clone())

        def sequence():
            cfg, cond = self._cfg_cond(self._gframe)
            with torch.cuda.device(self.dev):
                check(lib.nca_philox_mask_at(self.B, self.H, self.W, self.rate, 0, C.c_uint64(self.seed), 0, _ptr(self._tdev),
                                             self.step_n, _ptr(self._gmasks), _stream()))
            self._steps(cfg, cond, self._gmasks, 0)
            state_to_rgb8(self.slots[self.step_n & 1], 2.0, self._gout8)
            if self.step_n & 1:                 # every replay starts from slot 0
                self.slots[0].copy_(self.slots[1])
            self._tdev.add_(self.step_n)

        # warm-up on a side stream (allocations, function attributes), then capture; the state is restored afterwards
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            sequence()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            sequence()
        self.slots.copy_(keep[0])
        self._tdev.copy_(keep[1])

    @torch.no_grad()
    def push(self, frame_rgb, out=None):
        """frame_rgb [B,3,H,W] float in [-1,1] on the device -> uint8 [B,H,W,3] after `step_n` steps."""
        if self.use_graph:
            if self._graph is None:
                self._capture()
            self._gframe.copy_(frame_rgb)
            self._graph.replay()
            self.t0 += self.step_n
            if out is None:
                return self._gout8
            out.copy_(self._gout8)
            return out
        if self.cur == 1:                      # an odd step_n left the state in slot 1: the C ABI reads its input from slot 0
            self.slots[0].copy_(self.slots[1])
            self.cur = 0
        cfg, cond = self._cfg_cond(frame_rgb.contiguous())
        self._steps(cfg, cond, None, self.t0)
        self.t0 += self.step_n
        self.cur = self.step_n & 1
        return state_to_rgb8(self.slots[self.cur], 2.0, out)

    @torch.no_grad()
    def run(self, frames, out=None):
        """frames [F,3,H,W] (device or pinned host; one stream, batch 1) or [F,B,3,H,W] -> pinned uint8 [F*steps_per_frame,B,H,W,3]
        (`out`, or an internal pinned buffer that the next run() of the same length overwrites)."""
        if frames.dim() == 4:
            frames = frames.unsqueeze(1)
        F = frames.shape[0]
        n_out = F * self.steps_per_frame
        if out is None:
            # pinning is a cudaHostAlloc (~0.2 ms / MB): the buffer is kept and reused by later calls of the same length, so the
            # caller must consume (or copy) a result before the next run()
            if self._host is None or self._host.shape[0] != n_out:
                self._host = torch.empty(n_out, self.B, self.H, self.W, 3, dtype=torch.uint8).pin_memory()
            out = self._host
        elif tuple(out.shape) != (n_out, self.B, self.H, self.W, 3) or out.dtype != torch.uint8 or out.is_cuda:
            raise NcaError(f"out must be a host uint8 tensor of shape {(n_out, self.B, self.H, self.W, 3)}")
        host = out
        if self._dev_out is None:
            self._dev_out = [torch.empty(self.B, self.H, self.W, 3, device=self.dev, dtype=torch.uint8) for _ in range(2)]
            self._side = torch.cuda.Stream(self.dev)        # downloads
            self._side_up = torch.cuda.Stream(self.dev)     # uploads: their own stream, so that both directions of the link are busy
        dev_out = self._dev_out
        dev_in = None
        if not frames.is_cuda:
            if self._dev_in is None:
                self._dev_in = [torch.empty(self.B, 3, self.H, self.W, device=self.dev) for _ in range(2)]
            dev_in = self._dev_in
        main, side, side_up = torch.cuda.current_stream(self.dev), self._side, self._side_up
        side.wait_stream(main)
        side_up.wait_stream(main)
        done = [None, None]      # D2H of the frame that used dev_out[k] has finished
        up = [None, None]
        if dev_in is not None:
            with torch.cuda.stream(side_up):
                dev_in[0].copy_(frames[0], non_blocking=True)
                up[0] = torch.cuda.Event(); up[0].record(side_up)
        j = 0
        for f in range(F):
            if dev_in is not None:
                if f + 1 < F:
                    free = torch.cuda.Event(); free.record(main)          # compute that read dev_in[(f+1)&1] (frame f-1) is queued before
                    with torch.cuda.stream(side_up):
                        side_up.wait_event(free)
                        dev_in[(f + 1) & 1].copy_(frames[f + 1], non_blocking=True)
                        up[(f + 1) & 1] = torch.cuda.Event(); up[(f + 1) & 1].record(side_up)
                main.wait_event(up[f & 1])
                frame = dev_in[f & 1]
            else:
                frame = frames[f]
            for _ in range(self.steps_per_frame):
                k = j & 1
                if done[k] is not None:
                    main.wait_event(done[k])
                self.push(frame, dev_out[k])
                ready = torch.cuda.Event(); ready.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(ready)
                    host[j].copy_(dev_out[k], non_blocking=True)
                    done[k] = torch.cuda.Event(); done[k].record(side)
                j += 1
        main.wait_stream(side)
        main.wait_stream(side_up)
        side.synchronize()
        return host
