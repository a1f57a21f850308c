#!/usr/bin/env python
"""bench.py — NCA cell-updates/s of the NCA-step hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: a T-step rollout (forward, state history kept) plus BPTT through it.
Headline workload = BASELINE.json configs[1] ("c2", fit_video_motion): DyNCA 256x256, C=16, fc=128, CPE, perception
scales [0,1], replicate padding, batch 8 per GPU, T=128, gradients injected at the final state and at two rgb taps,
synthetic random state / reference-init weights.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|f16x3|fp32] [--only-main]

Precisions (--precision selects the headline's; the others are reported under "precisions"):
  bf16   update MLP and every GEMM of the BPTT on tcgen05 with BF16 operands, fp32 accumulation in TMEM; state, perception and
         gradients fp32 (north_star: 1e-2 per-step tolerance)                                                    [default]
  f16x3  the same GEMMs with split-precision FP16 hi + lo operands (three MMAs per product): fp32-grade parity (1e-5 / 1e-4)
  fp32   CUDA-core FFMA kernels (the parity reference of the library)

One JSON line:
  value        fwd+BPTT cell-updates/s of c2, whole job (all ranks), inputs resident in HBM, CUDA events, max over ranks
  fwd_value    forward-only (no_grad, no history) cell-updates/s, same way
  e2e          the same fwd+BPTT step through the drop-in nn.Module API with HOST buffers (pinned host -> device copy of the batch
               state, rollout, backward, device -> host copy of the final state and the weight gradients, all in the timed region)
  roofline     dominant kernel (the BPTT step kernel) against the slower of the HBM / tensor rooflines;
               "step" = the whole fwd+BPTT step against the same
  configs      the other BASELINE.json configurations (c1, c3, c4, c5) measured the same way, each with its roofline fractions;
               c3 and c4 are the STRONG-scaling shapes (global batch 64 / 256 split over the ranks)
  precisions   c2 in the other precisions
  cpu_baseline the reference's own DyNCA module (unmodified, loaded by file path from baseline/_ref or /root/reference; the oracle's
               ATen phrasing when neither is there) timed on this box's host cores on a bounded sample
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

# BASELINE.json configs: B = batch per rank for weak shapes, global batch for strong ones
CONFIGS = {
    "c1": dict(kind="dynca", flavour="ec", B=4, C=12, fc=96, H=128, W=128, T=64, scales=[0], pad="replicate", cond="cpe", taps=(),
               scaling="weak", desc="c1 vector-field motion fit: DyNCA 128x128 C=12 fc=96 CPE replicate, batch 4 per GPU, T=64"),
    "c2": dict(kind="dynca", flavour="ec", B=8, C=16, fc=128, H=256, W=256, T=128, scales=[0, 1], pad="replicate", cond="cpe", taps=(1, 65),
               scaling="weak", desc="c2 fit_video_motion: DyNCA 256x256 C=16 fc=128 CPE scales[0,1] replicate pad, batch 8 per GPU, "
                                    "T=128 rollout + BPTT, taps at t=1,65 + final"),
    "c3": dict(kind="dynca", flavour="cd", B=64, C=12, fc=96, H=256, W=256, T=80, scales=[0], pad="circular", cond="edges", taps=(),
               scaling="strong", desc="c3 conditioned stylization: DyNCA (edge conditioning) 256x256 C=12 fc=96 circular, GLOBAL batch 64 split "
                                      "over the ranks, T=80"),
    "c4": dict(kind="enc", B=256, C=20, H=64, W=64, T=72, rollouts=2, scaling="strong",
               desc="c4 encoder-conditioned NCA: ConditionedNCA 64x64 C=20 (60->64->64->20), GLOBAL batch 256 split over the ranks, "
                    "2 rollouts of T=72 per iteration, encoder forward/backward included"),
    "c5": dict(kind="dynca", flavour="ec", B=1, C=13, fc=96, H=1080, W=1920, T=256, scales=[0], pad="circular", cond=None, taps=(),
               scaling="weak", nograd=True, desc="c5 inference rollout: one 1920x1080 frame stream per GPU, C=13 (12 + gray), 256 steps per frame, no grad"),
}


def constants(c):
    """algorithmic flops / bytes per cell-update"""
    C = c["C"]
    if c["kind"] == "enc":
        f_mlp, f_perc = 2 * (3 * C * 64 + 64 * 64 + 64 * C), 54 * C
        b_fwd, b_both = 8 * C + 4 * 16, 20 * C + 12 * 16
    else:
        cc = {"cpe": 2, "edges": 3, None: 0}[c["cond"]]
        f_mlp = 2 * c["fc"] * (4 * C + cc + C)
        f_perc = 42 * C * sum(4.0 ** -s for s in c["scales"])
        ext = 3 if c["cond"] == "edges" else 0
        b_fwd, b_both = 8 * C + 4 * ext, 20 * C + 8 * ext
    return dict(f_mlp=f_mlp, f_perc=f_perc, flops_fwd=f_mlp + f_perc, flops_both=4 * f_mlp + 3 * f_perc,
                flops_bwd_kernel=3 * f_mlp + 2 * f_perc, bytes_fwd=b_fwd, bytes_both=b_both, bytes_bwd_kernel=12 * C)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops_sustained"], "measured"
    return 6650.0, 1400.0, "fallback"


def fractions(cups, flops, nbytes):
    """cell-updates/s -> fraction of the HBM roofline, of the tensor roofline, and of the slower (= binding) one"""
    hbm, tens, _ = peaks()
    hf, tf = cups * nbytes / (hbm * 1e9), cups * flops / (tens * 1e12)
    bound = "tensor" if flops / (tens * 1e12) >= nbytes / (hbm * 1e9) else "hbm"
    return {"bound": bound, "frac": tf if bound == "tensor" else hf, "hbm_frac": hf, "tensor_frac": tf}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.th.join(timeout=2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def ref_init(C, fc, cc, seed=0):
    """reference initialisation (dynca.py:56-61), generated on the CPU so every arm and rank sees the same numbers"""
    g = torch.Generator().manual_seed(seed)
    P = 4 * C + cc
    w1 = torch.randn(fc, P, generator=g) * (0.2 * (2.0 / (fc + P)) ** 0.5)
    b1 = (torch.rand(fc, generator=g) - 0.5) * (2.0 / P ** 0.5)
    w2 = torch.randn(C, fc, generator=g) * (0.1 * (2.0 / (fc + C)) ** 0.5)
    return w1, b1, w2, torch.zeros(C)


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own module when its source is reachable, else the oracle's ATen phrasing of the same step
# ---------------------------------------------------------------------------------------------------------------------
def load_reference_dynca():
    for root in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        p = os.path.join(root, "ExtraChannels", "models", "dynca.py")
        if os.path.exists(p):
            spec = importlib.util.spec_from_file_location("ref_ec_dynca", p)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod.DyNCA, p
    return None, None


def run_cpu_sample(steps=1, warmup=0, B=2, T=8):
    """fwd+BPTT of the c2 model on the host cores.  Returns (cell-updates/s, s per step, sample text, kind, warm-up steps run)."""
    c = CONFIGS["c2"]
    torch.set_num_threads(os.cpu_count())
    w1, b1, w2, b2 = ref_init(c["C"], c["fc"], 2)
    x0 = torch.rand(B, c["C"], c["H"], c["W"], generator=torch.Generator().manual_seed(42)) - 0.5
    DyNCA, path = load_reference_dynca()
    if DyNCA is not None:
        kind = "reference"
        model = DyNCA(c_in=c["C"], c_out=3, fc_dim=c["fc"], padding_mode=c["pad"], pos_emb="CPE", perception_scales=list(c["scales"]),
                      device=torch.device("cpu"))
        with torch.no_grad():
            model.w1.weight.copy_(w1.reshape(model.w1.weight.shape)); model.w1.bias.copy_(b1)
            model.w2.weight.copy_(w2.reshape(model.w2.weight.shape)); model.w2.bias.copy_(b2)
        params = list(model.parameters())

        def step():
            state, _, mids = model.forward_nsteps(x0, T, update_rate=0.5, return_middle_feature=True)     # its own torch.rand masks
            loss = state.square().mean() + mids[0].square().mean()
            torch.autograd.grad(loss, params)
        what = f"unmodified reference module ({os.path.relpath(path, ROOT) if path.startswith(ROOT) else path})"
    else:
        kind = "port"
        from oracle import nca_oracle as O
        masks = (torch.rand(T, B, 1, c["H"], c["W"], generator=torch.Generator().manual_seed(424)) + 0.5).floor()
        cond = O.cpe2d(B, c["H"], c["W"])
        params = [p.clone().requires_grad_(True) for p in (w1, b1, w2, b2)]

        def step():
            final = O.dynca_rollout_aten(x0, *params, masks, c["scales"], c["pad"], cond)
            torch.autograd.grad(final.square().mean(), params)
        what = "oracle restatement in the ATen ops the reference dispatches (reference source not present on this box)"
    times = []
    for _ in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    dt = sum(times[warmup:]) / steps
    sample = f"c2 model, B={B}, T={T}, 256x256, fwd + autograd backward, fp32, {what}"
    return B * c["H"] * c["W"] * T / dt, dt, sample, kind, warmup


def reference_arm(args, rank, world):
    if rank != 0:
        return
    warm = min(args.warmup, 2)
    v, dt, sample, kind, warm = run_cpu_sample(steps=max(1, min(args.steps, 5)), warmup=warm)
    line = {"impl": "reference", "metric": "NCA cell-updates/s (fwd+BPTT)", "value": v, "unit": "cell-updates/s",
            "n_gpus": args.gpus, "steps": max(1, min(args.steps, 5)), "warmup": warm, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIGS["c2"]["desc"], "sample": sample},
            "cpu_baseline": {"value": v, "unit": "cell-updates/s", "cores": os.cpu_count(), "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
class Workload:
    """model + resident inputs + train_step / fwd_only closures of one configuration on this rank"""

    def __init__(self, name, precision, dev, rank, world, dist):
        import nca_b200
        c = CONFIGS[name]
        self.c, self.name, self.dev, self.world, self.dist = c, name, dev, world, dist
        self.B = c["B"] // world if c["scaling"] == "strong" else c["B"]
        if self.B < 1:
            raise SystemExit(f"{name}: global batch {c['B']} cannot be split over {world} ranks")
        B, C, H, W, T = self.B, c["C"], c["H"], c["W"], c["T"]
        self.T = T
        self.cells = B * H * W * T * c.get("rollouts", 1)           # cell-updates per rank per step
        gen = torch.Generator().manual_seed(42 + rank)
        if c["kind"] == "enc":
            torch.manual_seed(0)
            self.model = nca_b200.ConditionedNCA(target_shape=(3, H, W), num_hidden_channels=C - 4, living_channel_dim=3, precision=precision).to(dev)
            with torch.no_grad():
                for p in self.model.update_net.parameters():
                    p.mul_(0.5)
            self.x0 = (self.model.generate_seed(B).to(dev) + 0.2 * torch.rand(B, C, H, W, generator=gen).to(dev))
            self.goal = torch.rand(B, 3, H, W, generator=gen).to(dev)
            self.extra = {}
        else:
            cc = {"cpe": 2, "edges": 3, None: 0}[c["cond"]]
            kw = dict(fc_dim=c["fc"], padding_mode=c["pad"], perception_scales=c["scales"], device=dev, precision=precision)
            if c["flavour"] == "ec":
                self.model = nca_b200.DyNCA_EC(C, 3, pos_emb="CPE" if c["cond"] == "cpe" else None, **kw)
            else:
                self.model = nca_b200.DyNCA_CD(C, 3, conditioning="edges", edge_transform="None", **kw)
            w1, b1, w2, b2 = ref_init(C, c["fc"], cc)
            with torch.no_grad():
                self.model.w1.weight.copy_(w1.reshape(self.model.w1.weight.shape)); self.model.w1.bias.copy_(b1)
                self.model.w2.weight.copy_(w2.reshape(self.model.w2.weight.shape)); self.model.w2.bias.copy_(b2)
            self.x0 = (torch.rand(B, C, H, W, generator=gen) - 0.5).to(dev)
            self.extra = {"cond_img": (torch.rand(B, 1, H, W, generator=gen) * 2 - 1).to(dev)} if c["cond"] == "edges" else {}
            n = B * C * H * W
            self.g_final = (torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(5)) / n).to(dev)
            self.g_tap = [(torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(6 + i)) / (B * 3 * H * W)).to(dev)
                          for i in range(len(c["taps"]))]
        if world > 1:       # same weights everywhere; rank-local batch and Philox stream
            for p in self.model.parameters():
                dist.broadcast(p.data, 0)
        self.params = [p for p in self.model.parameters() if p.requires_grad]

    def reduce(self, grads):
        """the path's one exchange step: one all-reduce of the weight gradients (24-52 KB).  The BPTT writes them back to back
        into one buffer, so the collective runs on that buffer in place (no flatten launch)."""
        if self.world > 1:
            from nca_b200 import parallel as par
            flat = par.flat_view(grads)
            if flat is None:
                flat = torch.cat([g.reshape(-1) for g in grads])
            self.dist.all_reduce(flat)
            return flat
        return None

    def train_step(self, x_in, seed):
        c = self.c
        if c["kind"] == "enc":
            flat, state = None, x_in
            for r in range(c["rollouts"]):
                state = self.model.grow(x_in, self.T, self.goal, seed=seed * 4 + r)
                grads = torch.autograd.grad(state.square().mean(), self.params)
                flat = self.reduce(grads)
            return state, grads
        state, _, mids = self.model.forward_nsteps(x_in, self.T, return_middle_feature=True, seed=seed, **self.extra)
        loss = (state * self.g_final).sum()
        for i, t in enumerate(c["taps"]):
            loss = loss + (mids[t - 1] * self.g_tap[i]).sum()
        grads = torch.autograd.grad(loss, self.params)
        self.reduce(grads)
        return state, grads

    def fwd_only(self, x_in, seed):
        with torch.no_grad():
            if self.c["kind"] == "enc":
                for r in range(self.c["rollouts"]):
                    out = self.model.grow(x_in, self.T, self.goal, seed=seed * 4 + r)
                return out
            return self.model.forward_nsteps(x_in, self.T, seed=seed, **self.extra)[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16", "f16x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only-main", action="store_true", help="skip the other configurations / precisions")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import nca_b200
    from nca_b200 import functional as Fn, _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the NCA step has no CPU fallback")
    lib = nca_b200.load_library()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
            ms = float(t.item())
        return ms / steps

    c2 = CONFIGS["c2"]
    K2 = constants(c2)
    wl = Workload("c2", args.precision, dev, rank, world, dist)
    model, x0, params, g_final = wl.model, wl.x0, wl.params, wl.g_final
    B, C, H, W, T = wl.B, c2["C"], c2["H"], c2["W"], c2["T"]
    cells = wl.cells
    n_grad = sum(p.numel() for p in params)

    with ClockSampler(local) as clocks:
        ms_train = timed(lambda i: wl.train_step(x0, 1000 + i), args.steps, args.warmup)
    lib.nca_launch_count_reset()
    wl.train_step(x0, 999)
    torch.cuda.synchronize()
    launches = lib.nca_launch_count() * args.steps       # kernels of libnca_b200.so inside the timed region
    ms_fwd = timed(lambda i: wl.fwd_only(x0, 1000 + i), args.steps, args.warmup)

    # ---- dominant kernel: the BPTT step kernel, timed through the C ABI call that launches it T times ----
    cfg = model._cfg(_lib.NCA_COND_CPE, 2)
    hist, coarse, (ops, _t1) = Fn._dynca_forward_raw(cfg, x0, *[p.detach() for p in (model.w1.weight, model.w1.bias, model.w2.weight, model.w2.bias)],
                                              None, None, 77, T, 0.5, True, want_ops=True)
    import ctypes as Ct
    d = cfg.desc(B, H, W, 0.5, False)
    nbytes = lib.nca_dynca_workspace_bytes(Ct.byref(d), 1)
    ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    wst = Fn._weights_struct(model.w1.weight, model.w1.bias, model.w2.weight, model.w2.bias)
    gouts = [torch.empty_like(p) for p in (model.w1.weight, model.w1.bias, model.w2.weight, model.w2.bias)]
    gst = Fn._weights_struct(*gouts)
    gx0 = torch.empty_like(x0)
    stream = Ct.c_void_p(torch.cuda.current_stream().cuda_stream)

    def bwd_call(i):
        Fn.check(lib.nca_dynca_backward(Ct.byref(d), Ct.byref(wst), None, None, Ct.c_uint64(77), 0, T, hist.data_ptr(),
                                        coarse.data_ptr() if coarse is not None else None, ops.data_ptr() if ops is not None else None, g_final.data_ptr(), (Ct.c_void_p * 1)(), (Ct.c_int32 * 1)(), 0, 1, 2.0,
                                        gx0.data_ptr(), Ct.byref(gst), ws.data_ptr(), nbytes, stream))
    ms_bwd_call = timed(bwd_call, max(2, args.steps // 2), 1)
    del hist, coarse, ops, ws
    ms_kernel = ms_bwd_call / T
    hbm, tens, which = peaks()
    cu_kernel = B * H * W / (ms_kernel * 1e-3)          # cell-updates/s of one BPTT launch
    fr = fractions(cu_kernel, K2["flops_bwd_kernel"], K2["bytes_bwd_kernel"])
    ach_tf, ach_gb = cu_kernel * K2["flops_bwd_kernel"] / 1e12, cu_kernel * K2["bytes_bwd_kernel"] / 1e9
    if fr["bound"] == "tensor":
        roof = {"bound": "tensor", "achieved": ach_tf, "peak": tens, "unit": "TFLOP/s", "frac": fr["frac"]}
    else:
        roof = {"bound": "hbm", "achieved": ach_gb, "peak": hbm, "unit": "GB/s", "frac": fr["frac"]}
    variant = Fn.dynca_kernel_variant(cfg, B, H, W, backward=True)
    kname = {0: "dynca_bwd_f32_kernel<2>", 1: "dynca_bwd_bf16_kernel<2, false>", 2: "dynca_bwd_tc2_kernel<2>",
             3: "dynca_bwd_bf16_kernel<2, true>"}[variant]
    traffic = None      # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(kname)
    step_cups = cells / (ms_train * 1e-3)               # per rank
    roof.update({"traffic": traffic, "kernel": kname, "kernel_ms": ms_kernel,
                 "peaks": which, "hbm_frac": fr["hbm_frac"], "tensor_frac": fr["tensor_frac"],
                 "step": fractions(step_cups, K2["flops_both"], K2["bytes_both"]),
                 "fwd": fractions(cells / (ms_fwd * 1e-3), K2["flops_fwd"], K2["bytes_fwd"]),
                 "note": "algorithmic flops 3*F_mlp+2*F_perc = %d, bytes 12C = %d per cell-update, %d cells per launch; "
                         "kernel time = nca_dynca_backward call / T, CUDA events on the launching stream; 'step' = whole fwd+BPTT step "
                         "(%d flop, %d B per cell-update), 'fwd' = forward-only rollout (%d flop, %d B), both per GPU"
                         % (K2["flops_bwd_kernel"], K2["bytes_bwd_kernel"], B * H * W, K2["flops_both"], K2["bytes_both"],
                            K2["flops_fwd"], K2["bytes_fwd"])})

    # ---- e2e: public module API with HOST buffers.  Every step's input comes from pinned host memory and every step's result
    #      (final state + flat weight gradients) goes back to pinned host memory, all inside the timed region.  The copies run on
    #      a second stream, double buffered: the input of step i+1 is uploaded and the result of step i-1 is downloaded (and waited
    #      for by the host) while step i computes - what a training loop with a prefetching loader does. ----
    from nca_b200 import parallel as par
    x_host = x0.cpu().pin_memory()
    out_host = [torch.empty_like(x_host).pin_memory() for _ in range(2)]
    gflat_host = [torch.empty(n_grad).pin_memory() for _ in range(2)]
    x_dev = [torch.empty_like(x0) for _ in range(2)]
    copy_s = torch.cuda.Stream(device=dev)
    main_s = torch.cuda.current_stream()

    def e2e_run(steps, base):
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [None, None]          # compute that last read x_dev[k] has finished
        ev_out = [None, None]           # result of the step that used slot k is on the host
        keep = [None, None]

        def upload(i):
            k = i & 1
            with torch.cuda.stream(copy_s):
                if ev_free[k] is not None:
                    copy_s.wait_event(ev_free[k])
                x_dev[k].copy_(x_host, non_blocking=True)
                ev_in[k].record(copy_s)

        upload(0)
        for i in range(steps):
            k = i & 1
            if i + 1 < steps:
                upload(i + 1)
            main_s.wait_event(ev_in[k])
            state, grads = wl.train_step(x_dev[k], base + i)
            gflat = par.flat_view(grads)
            if gflat is None:
                gflat = torch.cat([g.reshape(-1) for g in grads])
            done = torch.cuda.Event(); done.record(main_s)
            ev_free[k] = done
            keep[k] = (state, gflat)
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(done)
                out_host[k].copy_(state.detach(), non_blocking=True)
                gflat_host[k].copy_(gflat, non_blocking=True)
                ev_out[k] = torch.cuda.Event(); ev_out[k].record(copy_s)
            if i > 0:
                ev_out[1 - k].synchronize()        # the caller reads the previous step's result while this one computes
        ev_out[(steps - 1) & 1].synchronize()
        main_s.wait_stream(copy_s)

    e2e_run(args.warmup, 2000)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps, 3000)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms_e2e = float(t.item())
    del wl, model, x_dev, x0
    torch.cuda.empty_cache()

    # ---- the other configurations and precisions: same timing rules, fewer steps (each is a whole rollout of its own T) ----
    others, precs = {}, {}
    if not args.only_main:
        ks, kw = max(2, min(args.steps, 3)), max(3, min(args.warmup, 3))
        for name in ("c1", "c3", "c4", "c5"):
            c = CONFIGS[name]
            K = constants(c)
            w = Workload(name, args.precision if c["kind"] == "dynca" else ("bf16" if args.precision != "fp32" else "fp32"), dev, rank, world, dist)
            r = {"workload": c["desc"], "scaling": c["scaling"], "batch_per_gpu": w.B, "T": c["T"] * c.get("rollouts", 1),
                 "cell_updates_per_step": world * w.cells}
            msf = timed(lambda i: w.fwd_only(w.x0, 100 + i), ks, kw)
            r["fwd_value"] = world * w.cells / (msf * 1e-3)
            r["ms_per_step_fwd"] = msf
            r["roofline_fwd"] = fractions(w.cells / (msf * 1e-3), K["flops_fwd"], K["bytes_fwd"])
            if not c.get("nograd"):
                mst = timed(lambda i: w.train_step(w.x0, 100 + i), ks, kw)
                r["value"] = world * w.cells / (mst * 1e-3)
                r["ms_per_step"] = mst
                r["roofline_step"] = fractions(w.cells / (mst * 1e-3), K["flops_both"], K["bytes_both"])
            others[name] = r
            del w
            torch.cuda.empty_cache()
        for prec in ("bf16", "f16x3", "fp32"):
            if prec == args.precision:
                continue
            w = Workload("c2", prec, dev, rank, world, dist)
            mst = timed(lambda i: w.train_step(w.x0, 100 + i), 2, 3 if prec != "fp32" else 1)
            msf = timed(lambda i: w.fwd_only(w.x0, 100 + i), 2, 3 if prec != "fp32" else 1)
            precs[prec] = {"value": world * w.cells / (mst * 1e-3), "fwd_value": world * w.cells / (msf * 1e-3), "ms_per_step": mst,
                           "ms_per_step_fwd": msf, "warmup": 3 if prec != "fp32" else 1, "steps": 2,
                           "roofline_step": fractions(w.cells / (mst * 1e-3), K2["flops_both"], K2["bytes_both"])}
            del w
            torch.cuda.empty_cache()

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            v, dt, sample, kind, _ = run_cpu_sample(steps=2, warmup=1)
            cpu = {"value": v, "unit": "cell-updates/s", "cores": os.cpu_count(), "kind": kind, "sample": sample}
        line = {
            "metric": "NCA cell-updates/s (fwd+BPTT)", "value": world * cells / (ms_train * 1e-3), "unit": "cell-updates/s",
            "fwd_value": world * cells / (ms_fwd * 1e-3), "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_train, "ms_per_step_fwd": ms_fwd, "higher_is_better": True, "scaling": c2["scaling"],
            "vs_baseline": None, "dtype": {"fp32": "f32", "bf16": "bf16 MLP / f32 state", "f16x3": "f16 hi+lo (x3) MLP / f32 state"}[args.precision],
            "data": "synthetic",
            "config": {"workload": c2["desc"], "cells_per_step": world * cells,
                       "l2": "state history 4.3 GB per step >> 126 MB L2 (no flush needed)",
                       "mask": "in-kernel Philox", "parallelism": f"dp{world} (batch sharded, weight-grad all-reduce)"},
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": world * cells / (ms_e2e * 1e-3), "unit": "cell-updates/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": (out_host[0].numel() + gflat_host[0].numel()) * 4,
                    "copies": "second stream, double buffered (upload of step i+1 / download of step i-1 under step i)"},
            "gpu_launches": int(launches), "clocks": clocks.summary(),
            "configs": others, "precisions": precs,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
