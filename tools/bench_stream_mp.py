#!/usr/bin/env python
"""BASELINE.json config 5 at N GPUs: independent 1920x1080 frame streams sharded across ranks (inference shards by
streams, no collective on the data path), each rank a FrameStylizer (EC flavour C = 13, fc = 96, bf16 MLP, 256 steps per frame,
frames from pinned host memory, uint8 frames back on the host).  Timed on the device with CUDA events, max over ranks.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/bench_stream_mp.py
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nca_b200  # noqa: E402
from nca_b200 import video as V, parallel as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames-per-rank", type=int, default=4)
    ap.add_argument("--step-n", type=int, default=256)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    H, W, C = 1080, 1920, 13
    torch.manual_seed(0)
    m = nca_b200.DyNCA_EC(C, 3, fc_dim=96, padding_mode="circular", pos_emb=None, perception_scales=[0], device=dev, precision="bf16")
    with torch.no_grad():
        m.w2.weight.mul_(0.1)          # random-init weights: keep the 256-step rollout bounded (a trained model is)
    F = a.frames_per_rank
    lo, hi = P.shard_range(F * world, rank, world)                      # this rank's frames of the job
    clip = (torch.rand(hi - lo, 3, H, W, generator=torch.Generator().manual_seed(100 + rank)) * 2 - 1).pin_memory()
    st = V.FrameStylizer(m, (H, W), step_n=a.step_n, seed=P.rank_seed(1, rank))
    st.run(clip)                                                         # warm-up (allocations, pinned buffers)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        out = st.run(clip)
    e1.record()
    torch.cuda.synchronize()
    ms = P.max_over_ranks_ms(e0.elapsed_time(e1), dev)
    if rank == 0:
        frames = F * world * a.reps
        print(json.dumps({"metric": "NCA cell-updates/s (fwd, frame stream c5)", "value": frames * a.step_n * H * W / (ms * 1e-3),
                          "unit": "cell-updates/s", "frames_per_s": frames / (ms * 1e-3), "n_gpus": world, "scaling": "weak",
                          "config": {"workload": f"c5: 1920x1080, C=13, fc=96, bf16 MLP, {a.step_n} steps per frame, {F} frames per rank x {a.reps} reps, "
                                                 "frames from pinned host memory, uint8 frames to pinned host memory",
                                     "parallelism": f"{world} independent streams (no collective on the data path)"},
                          "ms": ms, "checksum": int(out.to(torch.int64).sum()), "finite_state": bool(torch.isfinite(st.state).all())}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
