"""Parity of the callers' kernels through the C ABI / host mirror against the CPU oracle and the
golden fixture made from the reference (tests/golden/callers.npz).

Bars: pool gather / scatter and the uint8 frame are byte / bit exact; the grayscale channel is bit exact against the reference's
CPU result (sum in channel order, then one division); the overflow loss within 1e-6 relative (fp32 reduction order) with a bit
exact gradient; the fused normalise + Adam step within 2e-6 relative per step of torch.optim.Adam on CPU (fp32 reduction order of
the norm, FMA contraction) and within 1e-5 after the fixture's 6 steps."""
import os

import numpy as np
import pytest
import torch

import nca_b200
from nca_b200 import trainer as Tr, video as V, _lib
from oracle import callers_oracle as K
from helpers import GOLDEN, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("shape", [(16, 12, 8, 8, 4, 1), (9, 12, 7, 5, 3, 0), (32, 19, 16, 12, 8, 2), (256, 12, 64, 64, 8, 1)])
@pytest.mark.parametrize("inject", [0, 1, 2])
def test_pool_gather_scatter_exact(shape, inject):
    N, Cp, H, W, B, Cx = shape
    g = torch.Generator().manual_seed(N + inject)
    pool = torch.randn(N, Cp, H, W, generator=g)
    extra = torch.randn(B, Cx, H, W, generator=g) if Cx else None
    seed_state = torch.randn(Cp, H, W, generator=g) if inject == 2 else None
    idx = np.random.RandomState(0).choice(N, B, replace=False)
    got = Tr.pool_gather(pool.to(DEV), idx, None if extra is None else extra.to(DEV), None if seed_state is None else seed_state.to(DEV), inject)
    assert torch.equal(got.cpu(), K.pool_gather(pool, idx, extra, seed_state, inject))
    after = torch.randn(B, Cp + Cx, H, W, generator=g)
    pd = pool.to(DEV)
    Tr.pool_scatter(pd, torch.as_tensor(idx), after.to(DEV))
    assert torch.equal(pd.cpu(), K.pool_scatter(pool, idx, after))


def test_pool_bad_arguments():
    pool = torch.zeros(4, 3, 8, 8, device=DEV)
    with pytest.raises(nca_b200.NcaError):
        Tr.pool_gather(pool, [0, 1], extra=torch.zeros(3, 1, 8, 8, device=DEV))
    with pytest.raises(nca_b200.NcaError):
        Tr.pool_gather(torch.zeros(4, 3, 8, 8), [0, 1])          # CPU tensor: no fallback
    with pytest.raises(nca_b200.NcaError):
        Tr.pool_scatter(pool, [0, 1], torch.zeros(2, 2, 8, 8, device=DEV))
    # out-of-range slots: zeros on gather, skipped on scatter (memory safe)
    got = Tr.pool_gather(pool + 1.0, [7, 1])
    assert float(got[0].abs().max()) == 0.0 and float(got[1].min()) == 1.0


def test_enc_sample_batch_reseeds_dead_and_unwritten_slots():
    """TensorSamplePool.sample_batch == ConditionedNCATrainer.sample_batch + `batch[:2] = seed` on the list-backed pool."""
    g = torch.Generator().manual_seed(8)
    Np, C, H, W, B, d = 12, 20, 16, 16, 6, 3
    seed_state = torch.zeros(C, H, W)
    seed_state[3:, H // 2, W // 2] = 1.0                       # generate_seed: centre cell on (nca.py:113-121)
    pool_list = [None] * Np
    tp = nca_b200.TensorSamplePool(Np, (C, H, W), DEV)
    assert len(tp) == Np and tp[0] is None
    vals = {}
    for slot in (1, 4, 5, 7, 9):
        x = torch.randn(C, H, W, generator=g) * 0.3
        if slot in (4, 9):
            x[d] = x[d].clamp(max=0.1)                         # dead: nothing above the threshold (0.1 itself is not alive)
        if slot == 7:
            x[d] = -1.0
            x[d, 0, W - 1] = 0.1000001                         # a single living corner cell
        vals[slot] = x
    slots = sorted(vals)
    tp[slots] = torch.stack([vals[s_] for s_ in slots]).to(DEV)
    for s_ in slots:
        pool_list[s_] = vals[s_]
    assert tp[1] is not None and torch.equal(tp[1].cpu(), vals[1])
    idx = [7, 2, 4, 1, 9, 5]
    got = tp.sample_batch(idx, seed_state.to(DEV), d, 0.1, inject_n=2)
    want = K.enc_sample_batch(pool_list, idx, seed_state, d, 0.1, 2)
    assert torch.equal(got.cpu(), want)
    assert torch.equal(got[3].cpu(), vals[1]) and torch.equal(got[4].cpu(), seed_state)


def test_normalized_adam_golden_fixture():
    d = np.load(os.path.join(GOLDEN, "callers.npz"))
    ps = [torch.nn.Parameter(torch.from_numpy(d[f"p{i}"]).to(DEV)) for i in range(4)]
    opt = nca_b200.NormalizedAdam(ps, lr=1e-3)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, [2, 4], 0.5)
    lib = nca_b200.load_library()
    for it in range(int(d["n_steps"])):
        for i, p in enumerate(ps):
            p.grad = torch.from_numpy(d[f"g{it}_{i}"]).to(DEV)
        lib.nca_launch_count_reset()
        opt.step()
        assert lib.nca_launch_count() == 1           # one launch for the whole model
        # the normalised gradient is left in p.grad, as the reference's in-place division leaves it
        gn = torch.from_numpy(d[f"g{it}_0"])
        assert rel_err(ps[0].grad.cpu(), gn / (gn.norm() + 1e-8)) < 2e-6
        opt.zero_grad(); sched.step()
    for i in range(4):
        assert rel_err(ps[i].detach().cpu(), d[f"p_final{i}"]) < 1e-5, i
    sd = opt.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and sd["state"][0]["step"] == int(d["n_steps"])


@pytest.mark.parametrize("norm_eps,normalize", [(1e-8, True), (1e-10, True), (1e-8, False)])
def test_normalized_adam_vs_torch_adam_stepwise(norm_eps, normalize):
    """10 parameter tensors of the ConditionedNCA + encoder shapes (conditioned_trainer.py:61,134-137), each step compared."""
    g = torch.Generator().manual_seed(5)
    shapes = [(60, 1, 3, 3), (64, 60, 1, 1), (64,), (64, 64, 1, 1), (64,), (20, 64, 1, 1), (16, 6, 3, 3), (16,), (16, 16, 3, 3), (16,)]
    init = [torch.randn(s, generator=g) * 0.2 for s in shapes]
    ref = [torch.nn.Parameter(p.clone()) for p in init]
    mine = [torch.nn.Parameter(p.clone().to(DEV)) for p in init]
    o_ref = torch.optim.Adam(ref, lr=2e-3, betas=(0.9, 0.999), eps=1e-8)
    o_mine = nca_b200.NormalizedAdam(mine, lr=2e-3, norm_eps=norm_eps, normalize=normalize, zero_grads=True)
    for it in range(12):
        for a, b in zip(ref, mine):
            gr = torch.randn(a.shape, generator=g) * (10.0 ** ((it % 5) - 2))
            if it == 7:
                gr.zero_()                                  # an all-zero gradient: 0 / (0 + eps) must stay 0, not NaN
            a.grad = gr.clone()
            b.grad = gr.clone().to(DEV)
        if normalize:
            for p in ref:
                p.grad /= torch.norm(p.grad) + norm_eps
        o_ref.step()
        o_mine.step()
        for a, b in zip(ref, mine):
            assert rel_err(b.detach().cpu(), a.detach()) < 2e-6 * (it + 1), (it, tuple(a.shape))
            assert float(b.grad.abs().max()) == 0.0          # zero_grads folded in


def test_normalized_adam_more_than_16_tensors():
    g = torch.Generator().manual_seed(6)
    init = [torch.randn(7 + i, generator=g) for i in range(21)]
    ref = [torch.nn.Parameter(p.clone()) for p in init]
    mine = [torch.nn.Parameter(p.clone().to(DEV)) for p in init]
    o_ref, o_mine = torch.optim.Adam(ref, lr=1e-2), nca_b200.NormalizedAdam(mine, lr=1e-2)
    for a, b in zip(ref, mine):
        a.grad = torch.randn(a.shape, generator=g)
        b.grad = a.grad.clone().to(DEV)
        a.grad /= a.grad.norm() + 1e-8
    o_ref.step(); o_mine.step()
    for a, b in zip(ref, mine):
        assert rel_err(b.detach().cpu(), a.detach()) < 2e-6


@pytest.mark.parametrize("shape", [(2, 13, 24, 40), (1, 12, 7, 5), (8, 16, 256, 256), (3, 5, 1, 3)])
def test_overflow_loss_and_gradient(shape):
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g) * 1.2
    x.view(-1)[:3] = torch.tensor([1.0, -1.0, 1.0000001])        # the clamp boundaries
    xd = x.clone().to(DEV).requires_grad_(True)
    loss = nca_b200.overflow_loss(xd)
    (loss * 3.0).backward()
    assert rel_err(loss.detach().cpu(), K.overflow_loss(x)) < 1e-6
    assert torch.equal(xd.grad.cpu(), K.overflow_grad(x) * 3.0)
    # run-to-run reproducible (fixed reduction order)
    assert float(nca_b200.overflow_loss(xd.detach())) == float(loss)
    # accumulate into an existing g_final
    gf = torch.full(shape, 0.25, device=DEV)
    l2 = Tr.overflow_loss_into(xd.detach(), gf, weight=2.0)
    assert float(l2) == float(loss)
    assert torch.equal(gf.cpu(), 0.25 + K.overflow_grad(x, 2.0))


def test_overflow_golden_fixture():
    d = np.load(os.path.join(GOLDEN, "callers.npz"))
    x = torch.from_numpy(d["state"]).to(DEV).requires_grad_(True)
    loss = nca_b200.overflow_loss(x)
    loss.backward()
    assert abs(float(loss) - float(d["overflow"])) <= 1e-6 * float(d["overflow"])
    assert np.array_equal(x.grad.cpu().numpy(), d["overflow_grad"])


def test_frame_kernels_golden_fixture():
    d = np.load(os.path.join(GOLDEN, "callers.npz"))
    frame = torch.from_numpy(d["frame"]).to(DEV)
    state = torch.from_numpy(d["state"]).to(DEV)
    assert np.array_equal(V.rgb_to_grayscale(frame).cpu().numpy(), d["gray"])
    before = state.clone()
    V.frame_to_cond_channel(state, frame, -1)
    assert np.array_equal(state[:, -1:].cpu().numpy(), d["gray"]) and torch.equal(state[:, :-1], before[:, :-1])
    assert np.array_equal(V.state_to_rgb8(before).cpu().numpy(), d["rgb8"])


@pytest.mark.parametrize("shape", [(1, 13, 1080, 1920), (2, 12, 9, 7), (3, 3, 16, 12)])
def test_frame_kernels_vs_oracle(shape):
    B, C, H, W = shape
    g = torch.Generator().manual_seed(H)
    state = torch.randn(shape, generator=g) * 0.4
    # every float that lands near a uint8 boundary matters: add exact k/255 pre-images
    k = torch.arange(0, 256, dtype=torch.float32)
    pre = ((k / 255.0) * 2.0 - 1.0) / 2.0
    n = min(pre.numel(), H * W)
    state[0, 0].view(-1)[:n] = pre[:n]
    frame = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    got = V.state_to_rgb8(state.to(DEV)).cpu().numpy()
    assert np.array_equal(got, K.state_to_rgb8(state))
    sd = state.to(DEV)
    V.frame_to_cond_channel(sd, frame.to(DEV), C - 1)
    assert torch.equal(sd.cpu(), K.frame_to_cond_channel(state, frame, C - 1))


def _ec_model(C=13, fc=96, precision="fp32"):
    torch.manual_seed(0)
    m = nca_b200.DyNCA_EC(C, 3, fc_dim=fc, padding_mode="circular", pos_emb=None, perception_scales=[0], device=torch.device(DEV), precision=precision)
    with torch.no_grad():
        m.w2.weight.mul_(3.0)
    return m


@pytest.mark.parametrize("step_n", [8, 5])
def test_frame_stylizer_equals_reference_loop(step_n):
    """FrameStylizer.run == the loop of save_video (video_utils.py:65-82) written with the module API and host-side numpy,
    frame by frame, for the same Philox key (even and odd step counts: the odd one exercises the slot copy)."""
    H, W, F = 32, 48, 4
    m = _ec_model()
    frames = torch.rand(F, 3, H, W, generator=torch.Generator().manual_seed(1)) * 2 - 1
    st = V.FrameStylizer(m, (H, W), step_n=step_n, seed=1234)
    got = st.run(frames.pin_memory()).numpy()
    assert got.shape == (F, 1, H, W, 3)
    # reference loop (video_utils.py:66-82), fire masks materialised from the same Philox stream
    with torch.no_grad():
        h = m.seed(1, size=(W, H))
        t0 = 0
        for f in range(F):
            fr = frames[f].unsqueeze(0).to(DEV)
            h = torch.cat((h, torch.mean(fr, dim=1, keepdim=True)), 1)
            masks = nca_b200.functional.philox_mask(1, H, W, 0.5, 1234, step_n, t0=t0)
            nca_state, nca_feature = m.forward_nsteps(h, step_n, masks=masks)
            t0 += step_n
            h = nca_state[:, :-1, :, :]
            img = nca_feature.detach().cpu().numpy()[0].transpose(1, 2, 0)
            img = np.clip(img, -1.0, 1.0)
            img = (img + 1.0) / 2.0
            assert np.array_equal(got[f, 0], np.uint8(img.clip(0, 1) * 255)), f
    assert got.std() > 0


def test_frame_stylizer_cd_edges_and_device_frames():
    H, W, F = 32, 32, 3
    torch.manual_seed(0)
    m = nca_b200.DyNCA_CD(12, 3, fc_dim=96, padding_mode="circular", conditioning="edges", edge_transform="None",
                          perception_scales=[0], device=torch.device(DEV))
    with torch.no_grad():
        m.w2.weight.mul_(3.0)
    frames = (torch.rand(F, 3, H, W, generator=torch.Generator().manual_seed(2)) * 2 - 1).to(DEV)
    st = V.FrameStylizer(m, (H, W), step_n=6, seed=99)
    got = st.run(frames).numpy()
    with torch.no_grad():
        h = m.seed(1, size=(W, H))
        for f in range(F):
            gray = torch.mean(frames[f:f + 1], dim=1, keepdim=True)
            masks = nca_b200.functional.philox_mask(1, H, W, 0.5, 99, 6, t0=6 * f)
            h, feat = m.forward_nsteps(h, 6, cond_img=gray, masks=masks)
            assert np.array_equal(got[f, 0], K.state_to_rgb8(h)[0]), f


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_frame_sequence_is_graph_capturable(precision):
    """include/nca_b200.h promises that every entry point only enqueues work on the caller's stream: the per-frame sequence of the
    inference stream (conditioning channel, step_n steps with programmatic dependent launches and TMA tensor maps, rgb8 packing)
    is captured into one CUDA graph and replayed on new inputs."""
    H, W, T = 64, 64, 6
    m = _ec_model(precision=precision)
    g0 = torch.Generator().manual_seed(3)
    xa = (torch.rand(2, 13, H, W, generator=g0) - 0.5).to(DEV)
    xb = (torch.rand(2, 13, H, W, generator=g0) - 0.5).to(DEV)
    fr = (torch.rand(2, 3, H, W, generator=g0) * 2 - 1).to(DEV)

    def seq(xin, out8):
        V.frame_to_cond_channel(xin, fr, -1)
        s, _ = m.forward_nsteps(xin, T, seed=5)
        V.state_to_rgb8(s, 2.0, out8)
        return s

    with torch.no_grad():
        refs = []
        for x in (xa, xb):
            o8 = torch.empty(2, H, W, 3, device=DEV, dtype=torch.uint8)
            refs.append((seq(x.clone(), o8).clone(), o8))
        torch.cuda.synchronize()
        xs = xa.clone()
        out8 = torch.empty(2, H, W, 3, device=DEV, dtype=torch.uint8)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            got = seq(xs, out8)
        for x, (ref, ref8) in zip((xb, xa), (refs[1], refs[0])):
            xs.copy_(x)
            graph.replay()
            torch.cuda.synchronize()
            assert torch.equal(got, ref) and torch.equal(out8, ref8)


@pytest.mark.parametrize("flavour,step_n,precision", [("ec", 8, "fp32"), ("ec", 5, "bf16"), ("cd", 6, "bf16")])
def test_frame_stylizer_graph_mode_equals_eager(flavour, step_n, precision):
    """graph=True (mask draw from a device-side step counter + rollout with supplied masks + rgb8, one CUDA graph per frame)
    produces the frames of the eager stream bit for bit, across two clips and a reset."""
    H, W, F = 32, 64, 5
    if flavour == "ec":
        m = _ec_model(precision=precision)
    else:
        torch.manual_seed(0)
        m = nca_b200.DyNCA_CD(12, 3, fc_dim=96, padding_mode="circular", conditioning="edges", edge_transform="tanh",
                              perception_scales=[0], device=torch.device(DEV), precision=precision)
        with torch.no_grad():
            m.w2.weight.mul_(3.0)
    frames = (torch.rand(F, 3, H, W, generator=torch.Generator().manual_seed(7)) * 2 - 1)
    eager = V.FrameStylizer(m, (H, W), step_n=step_n, seed=77)
    graph = V.FrameStylizer(m, (H, W), step_n=step_n, seed=77, graph=True)
    a1 = eager.run(frames.pin_memory()).clone().numpy()
    b1 = graph.run(frames.to(DEV)).clone().numpy()
    assert np.array_equal(a1, b1) and a1.std() > 0
    assert torch.equal(eager.state, graph.state)
    # the stream continues (step counter on the device keeps advancing), then starts over after reset()
    a2 = eager.run(frames.flip(0).pin_memory()).clone().numpy()
    b2 = graph.run(frames.flip(0).pin_memory()).clone().numpy()
    assert np.array_equal(a2, b2) and not np.array_equal(a1, a2)
    eager.reset(); graph.reset()
    assert np.array_equal(graph.run(frames.to(DEV)).numpy(), a1)


def test_pool_round_trip_at_c3_size():
    """BASELINE.json config 3 at full size (pool 256 x 12 x 256 x 256, batch 64, + the EC conditioning channel): gather ->
    scatter into a second pool reproduces exactly the sampled slots and touches nothing else; the appended channel is `extra`."""
    N, Cp, H, W, B = 256, 12, 256, 256, 64
    g = torch.Generator(device=DEV).manual_seed(0)
    pool = torch.rand(N, Cp, H, W, device=DEV, generator=g) - 0.5
    extra = torch.rand(B, 1, H, W, device=DEV, generator=g)
    idx = torch.as_tensor(np.random.RandomState(1).choice(N, B, replace=False)).to(DEV)
    batch = Tr.pool_gather(pool, idx, extra)
    assert torch.equal(batch[:, :Cp], pool[idx]) and torch.equal(batch[:, Cp:], extra)
    other = torch.zeros_like(pool)
    Tr.pool_scatter(other, idx, batch)
    assert torch.equal(other[idx], pool[idx])
    untouched = torch.ones(N, dtype=torch.bool, device=DEV)
    untouched[idx] = False
    assert float(other[untouched].abs().max()) == 0.0
    # seed injection replaces exactly the first sample
    inj = Tr.pool_gather(pool, idx, extra, None, 1)
    assert float(inj[0, :Cp].abs().max()) == 0.0 and torch.equal(inj[1:], batch[1:]) and torch.equal(inj[0, Cp:], extra[0])


def test_frame_stylizer_steps_per_frame_and_batch():
    """steps_per_frame > 1 emits one image per block of step_n steps on the same target frame (video_utils.py:70-82); a batch of
    independent streams [F,B,3,H,W] equals the streams run one by one with the same key (the Philox counter includes b)."""
    H, W, F = 32, 32, 3
    m = _ec_model()
    frames = (torch.rand(F, 2, 3, H, W, generator=torch.Generator().manual_seed(9)) * 2 - 1).to(DEV)
    st = V.FrameStylizer(m, (H, W), step_n=4, steps_per_frame=2, batch=2, seed=5)
    got = st.run(frames).clone()
    assert tuple(got.shape) == (2 * F, 2, H, W, 3)
    ref = V.FrameStylizer(m, (H, W), step_n=4, steps_per_frame=1, batch=2, seed=5)
    j = 0
    for f in range(F):
        for _ in range(2):
            assert torch.equal(ref.push(frames[f]).cpu(), got[j]), (f, j)
            j += 1
    assert torch.equal(ref.state, st.state)
    # stream 0 of the batch alone: same key, same b = 0 -> identical frames
    one = V.FrameStylizer(m, (H, W), step_n=4, steps_per_frame=2, batch=1, seed=5)
    assert torch.equal(one.run(frames[:, :1])[:, 0], got[:, 0])
