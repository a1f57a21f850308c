#!/usr/bin/env python
"""bench.py — NCA cell-updates/s of the DyNCA hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: a T-step DyNCA rollout (forward, state history kept) plus
BPTT through it with gradients injected at the final state and at two rgb taps (the fit_video_motion.py
pattern).  Workload = BASELINE.json configs[1] ("c2"): 256x256, C=16, fc=128, CPE, perception
scales [0,1], replicate padding, batch 8 per GPU, T=128, synthetic random state / reference-init weights.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]

Default precision is bf16: the update MLP and every GEMM of the BPTT step run on tcgen05 with BF16 operands and fp32
accumulation in TMEM, state / perception / gradients stay fp32 (north_star: 1e-2 per-step tolerance in this mode);
--precision fp32 times the CUDA-core parity path.

value      = fwd+BPTT cell-updates/s, whole job (all ranks), inputs resident in HBM, CUDA-event timed, max over ranks
fwd_value  = forward-only (no_grad, no history) cell-updates/s, measured the same way
e2e        = the same fwd+BPTT step through the drop-in nn.Module API with HOST buffers: pinned host -> device
             copy of the batch state, rollout, backward, device -> host copy of the final state and weight grads
roofline   = dominant kernel (the BPTT step kernel) vs the slower of the HBM / tensor rooflines
cpu_baseline = the oracle's ATen phrasing of the same step (== what the reference dispatches on CPU) timed on
             this box's host cores on a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(B=8, C=16, fc=128, H=256, W=256, T=128, scales=[0, 1], pad="replicate", cond="cpe", taps=(1, 65))
F_MLP = 2 * CFG["fc"] * (4 * CFG["C"] + 2 + CFG["C"])        # 20 992 flop / cell-update
F_PERC = 42 * CFG["C"] * 1.25                                 # 840
FLOPS_FWD = F_MLP + F_PERC
FLOPS_BWD_KERNEL = 3 * F_MLP + 2 * F_PERC                     # recompute + dgrad + wgrad
FLOPS_FWD_BPTT = 4 * F_MLP + 3 * F_PERC                       # 86 488
BYTES_FWD = 8 * CFG["C"]                                      # 128 B
BYTES_BWD_KERNEL = 12 * CFG["C"]                              # read x_t, read g_{t+1}, write g_t
BYTES_FWD_BPTT = 20 * CFG["C"]                                # 320 B


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops_sustained"], "measured"
    return 6650.0, 1400.0, "fallback"


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


def cpu_inputs(B, T, seed=0):
    g = torch.Generator().manual_seed(seed)
    C, fc, H, W = CFG["C"], CFG["fc"], CFG["H"], CFG["W"]
    P = 4 * C + 2
    w1 = torch.randn(fc, P, generator=g) * (0.2 * (2.0 / (fc + P)) ** 0.5)     # xavier_normal gain 0.2 (dynca.py:57)
    b1 = (torch.rand(fc, generator=g) - 0.5) * (2.0 / P ** 0.5)
    w2 = torch.randn(C, fc, generator=g) * (0.1 * (2.0 / (fc + C)) ** 0.5)     # gain 0.1, zero bias (dynca.py:60-61)
    b2 = torch.zeros(C)
    x0 = torch.rand(B, C, H, W, generator=torch.Generator().manual_seed(42)) - 0.5
    return x0, w1, b1, w2, b2


def run_cpu_sample(steps=1, warmup=0, B=2, T=6):
    """fwd+BPTT of the c2 architecture on the host cores via the oracle's ATen phrasing. Returns cell-updates/s."""
    from oracle import nca_oracle as O
    torch.set_num_threads(os.cpu_count())
    x0, w1, b1, w2, b2 = cpu_inputs(B, T)
    masks = (torch.rand(T, B, 1, CFG["H"], CFG["W"], generator=torch.Generator().manual_seed(424)) + 0.5).floor()
    cond = O.cpe2d(B, CFG["H"], CFG["W"])
    params = [p.clone().requires_grad_(True) for p in (w1, b1, w2, b2)]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        final = O.dynca_rollout_aten(x0, *params, masks, CFG["scales"], CFG["pad"], cond)
        final.square().mean().backward()
        times.append(time.perf_counter() - t0)
    dt = sum(times[warmup:]) / steps
    return B * CFG["H"] * CFG["W"] * T / dt, dt, f"c2 architecture, B={B}, T={T}, 256x256, fwd+backward (autograd), fp32"


def reference_arm(args, rank, world):
    if rank != 0:
        return
    v, dt, sample = run_cpu_sample(steps=args.steps, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": "NCA cell-updates/s (fwd+BPTT)", "value": v, "unit": "cell-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "c2: DyNCA 256x256 C=16 fc=128 CPE scales[0,1] replicate, fwd+BPTT", "sample": sample},
            "cpu_baseline": {"value": v, "unit": "cell-updates/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    B, C, fc, H, W, T = (CFG[k] for k in ("B", "C", "fc", "H", "W", "T"))
    model = nca_b200.DyNCA_EC(C, 3, fc_dim=fc, padding_mode=CFG["pad"], pos_emb="CPE", perception_scales=CFG["scales"],
                              device=dev, precision=args.precision)
    x0c, w1, b1, w2, b2 = cpu_inputs(B, T, seed=0)
    with torch.no_grad():
        model.w1.weight.copy_(w1.reshape(model.w1.weight.shape)); model.w1.bias.copy_(b1)
        model.w2.weight.copy_(w2.reshape(model.w2.weight.shape)); model.w2.bias.copy_(b2)
    if world > 1:   # same weights everywhere, rank-local batch and Philox stream
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    x0 = (torch.rand(B, C, H, W, generator=torch.Generator().manual_seed(42 + rank)) - 0.5).to(dev)
    g_final = (torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(5)) / (B * C * H * W)).to(dev)
    g_tap = [(torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(6 + i)) / (B * 3 * H * W)).to(dev)
             for i in range(len(CFG["taps"]))]
    params = list(model.parameters())
    flat_grads = torch.zeros(sum(p.numel() for p in params), device=dev)

    def train_step(x_in, seed):
        """rollout + BPTT (+ gradient all-reduce when sharded); returns final state"""
        state, _, mids = model.forward_nsteps(x_in, T, return_middle_feature=True, seed=seed)
        loss = (state * g_final).sum()
        for i, t in enumerate(CFG["taps"]):
            loss = loss + (mids[t - 1] * g_tap[i]).sum()
        grads = torch.autograd.grad(loss, params)
        if world > 1:
            torch.cat([g.reshape(-1) for g in grads], out=flat_grads)
            dist.all_reduce(flat_grads)      # the path's one exchange step: ~42 KB of weight gradients
        return state, grads

    def fwd_only(x_in, seed):
        with torch.no_grad():
            return model.forward_nsteps(x_in, T, seed=seed)[0]

    def timed(fn, steps, warmup, sampler=None):
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

    cells = B * H * W * T                 # per rank per step
    with ClockSampler(local) as clocks:
        ms_train = timed(lambda i: train_step(x0, 1000 + i), args.steps, args.warmup)
    lib.nca_launch_count_reset()
    train_step(x0, 999)
    torch.cuda.synchronize()
    launches = lib.nca_launch_count() * args.steps       # kernels of libnca_b200.so inside the timed region
    ms_fwd = timed(lambda i: fwd_only(x0, 1000 + i), args.steps, args.warmup)

    # ---- dominant kernel: the BPTT step kernel, timed through the C ABI call that launches it T times ----
    cfg = model._cfg(_lib.NCA_COND_CPE, 2)
    hist, coarse, ops = Fn._dynca_forward_raw(cfg, x0, *[p.detach() for p in (model.w1.weight, model.w1.bias, model.w2.weight, model.w2.bias)],
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
    del hist, coarse, ops
    ms_kernel = ms_bwd_call / T
    hbm, tens, which = peaks()
    cu_kernel = B * H * W / (ms_kernel * 1e-3)          # cell-updates/s of one BPTT launch
    ach_tf = cu_kernel * FLOPS_BWD_KERNEL / 1e12
    ach_gb = cu_kernel * BYTES_BWD_KERNEL / 1e9
    # the bound is the slower of the two rooflines for this kernel (north_star)
    t_hbm, t_tens = BYTES_BWD_KERNEL / (hbm * 1e9), FLOPS_BWD_KERNEL / (tens * 1e12)
    if t_tens >= t_hbm:
        roof = {"bound": "tensor", "achieved": ach_tf, "peak": tens, "unit": "TFLOP/s", "frac": ach_tf / tens}
    else:
        roof = {"bound": "hbm", "achieved": ach_gb, "peak": hbm, "unit": "GB/s", "frac": ach_gb / hbm}
    variant = Fn.dynca_kernel_variant(cfg, B, H, W, backward=True)
    kname = {0: "dynca_bwd_f32_kernel<2>", 1: "dynca_bwd_bf16_kernel<2>", 2: "dynca_bwd_tc2_kernel<2, true>"}[variant]      # <two scales, operand history>
    traffic = None      # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(kname)
    roof.update({"traffic": traffic, "kernel": kname, "kernel_ms": ms_kernel,
                 "peaks": which, "hbm_frac": ach_gb / hbm, "tensor_frac": ach_tf / tens,
                 "note": "algorithmic flops 3*F_mlp+2*F_perc = %d, bytes 12C = %d per cell-update, %d cells per launch; "
                         "kernel time = nca_dynca_backward call / T, CUDA events on the launching stream"
                         % (FLOPS_BWD_KERNEL, BYTES_BWD_KERNEL, B * H * W)})

    # ---- e2e: public module API with HOST buffers.  Every step's input comes from pinned host memory and every step's result
    #      (final state + flat weight gradients) goes back to pinned host memory, all inside the timed region.  The copies run on
    #      a second stream, double buffered: the input of step i+1 is uploaded and the result of step i-1 is downloaded (and waited
    #      for by the host) while step i computes - what a training loop with a prefetching loader does. ----
    x_host = x0.cpu().pin_memory()
    out_host = [torch.empty_like(x_host).pin_memory() for _ in range(2)]
    gflat_host = [torch.empty(flat_grads.numel()).pin_memory() for _ in range(2)]
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
            state, grads = train_step(x_dev[k], base + i)
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

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            v, dt, sample = run_cpu_sample(steps=2, warmup=1)
            cpu = {"value": v, "unit": "cell-updates/s", "cores": os.cpu_count(), "kind": "port", "sample": sample}
        line = {
            "metric": "NCA cell-updates/s (fwd+BPTT)", "value": world * cells / (ms_train * 1e-3), "unit": "cell-updates/s",
            "fwd_value": world * cells / (ms_fwd * 1e-3), "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_train, "ms_per_step_fwd": ms_fwd, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16 MLP / f32 state", "data": "synthetic",
            "config": {"workload": "c2 fit_video_motion: DyNCA 256x256 C=16 fc=128 CPE scales[0,1] replicate pad, "
                                   "batch 8 per GPU, T=128 rollout + BPTT, taps at t=1,65 + final",
                       "cells_per_step": world * cells, "l2": "state history 4.3 GB per step >> 126 MB L2 (no flush needed)",
                       "mask": "in-kernel Philox", "parallelism": f"dp{world} (batch sharded, weight-grad all-reduce)"},
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": world * cells / (ms_e2e * 1e-3), "unit": "cell-updates/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": (out_host[0].numel() + gflat_host[0].numel()) * 4,
                    "copies": "second stream, double buffered (upload of step i+1 / download of step i-1 under step i)"},
            "gpu_launches": int(launches), "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
