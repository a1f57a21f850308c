"""Host-side logic of the N>1 path on CPU: world_size-2 gloo.  Each rank owns a shard of the batch, computes
weight gradients of its shard (here with the oracle standing in for the kernels — the point is the plumbing),
and one all-reduce must reproduce the 1-rank gradients of the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nca_b200 import parallel as par


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 64, 257):
        for world in (1, 2, 3, 8):
            spans = [par.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        par.shard_range(4, 2, 2)
    assert len({par.rank_seed(5, r) for r in range(8)}) == 8


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, root)
        from oracle import nca_oracle as O
        torch.manual_seed(0)
        B, C, fc, H, W, T = 4, 12, 32, 8, 8, 3
        w1 = (torch.randn(fc, 4 * C) * 0.2).requires_grad_(True)
        b1 = (torch.randn(fc) * 0.1).requires_grad_(True)
        w2 = (torch.randn(C, fc) * 0.2).requires_grad_(True)
        b2 = (torch.randn(C) * 0.1).requires_grad_(True)
        params = [w1, b1, w2, b2]
        x0 = torch.rand(B, C, H, W) - 0.5
        masks = (torch.rand(T, B, 1, H, W) + 0.5).floor()
        lo, hi = par.shard_range(B, rank, world)
        final = O.dynca_rollout(x0[lo:hi], *params, masks[:, lo:hi], (0,), "circular", None)
        final.square().sum().backward()
        flat = par.allreduce_grads(params)
        ms = par.max_over_ranks_ms(10.0 + rank, torch.device("cpu"))
        # 1-rank run of the whole batch
        ref = [p.detach().clone().requires_grad_(True) for p in params]
        O.dynca_rollout(x0, *ref, masks, (0,), "circular", None).square().sum().backward()
        err = max(float((p.grad - r.grad).abs().max() / (r.grad.abs().max() + 1e-30)) for p, r in zip(params, ref))
        par.normalize_grads_(params)
        norms = [float(p.grad.norm()) for p in params]
        q.put((rank, err, ms, flat.numel(), norms))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_single_rank():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, ms, n, norms in res:
        assert err < 1e-5, (rank, err)
        assert ms == 11.0                       # max over ranks of (10, 11)
        assert n == 32 * 48 + 32 + 12 * 32 + 12
        assert all(abs(v - 1.0) < 1e-5 for v in norms)
