"""GPU parity of the teacher-forced faceformer_vert training step (BASELINE configs[4]) through the drop-in
`FaceformerVert.forward(...) -> loss; loss.backward(); FlatAdam.step()` against
  (1) tests/golden/train.npz - the reference's OWN forward_switch_frame + loss.backward() + torch.optim.Adam (oracle/make_golden.py),
  (2) oracle/train_oracle.py (torch autograd over the CPU oracle) at the config-5 clip shape.
Tolerances: fp32 mode - loss rel 1e-4, every gradient tensor max|err| <= 5e-4 * max|ref|; bf16 GEMM mode - loss rel 1e-2 and
gradient relative L2 error <= 5e-2 per tensor (gradients through 12 bf16 layers; the north-star 1e-2 is for forward coefficients).
"""
import numpy as np
import pytest
import torch
import torch.nn as nn

from avi_talking_b200 import synth, train
from oracle.make_golden import GRAD_STRIDE, train_inputs
from helpers import build_flame, build_wav2vec

pytestmark = pytest.mark.gpu


def build_vert(precision, fd, seed):
    from avi_talking_b200.faceformer import FaceformerVert, make_args
    w2v = build_wav2vec(precision, device="cpu")
    flame = build_flame(100, device="cuda", mediapipe=True, precision="fp32")
    m = FaceformerVert(make_args(feature_dim=fd), audio_encoder=w2v, flame=flame)
    sd = synth.faceformer_state(fd=fd, seed=seed, variant="vert")
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected
    m.precision = precision
    m.audio_encoder.precision = precision
    return m.cuda()


def sub(t, stride=GRAD_STRIDE):
    t = t.detach().reshape(-1)
    return (t[::stride] if t.numel() > 4096 else t).float().cpu().numpy()


def exactly_zero(ref, got):
    """Gradients that are mathematically zero (softmax is invariant to the key bias): both sides hold rounding noise only."""
    return np.abs(ref).max() < 1e-10 and np.abs(got).max() < 1e-10


def run_case(precision, fd, B, n_samples, T):
    m = build_vert(precision, fd, 200 + fd)
    coeff, pose, shape, mean, std = train_inputs(B, T, seed=90 + fd)
    m.coeff_mean, m.coeff_std = mean.cuda(), std.cuda()
    audio = synth.audio(B, n_samples, seed=4321).cuda()
    opt = train.FlatAdam(m, lr=1e-4)
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    opt.zero_grad()
    loss = m(audio, coeff.cuda(), pose.cuda(), shape.cuda(), criterion=nn.MSELoss(reduction="none"), teacher_forcing=True)
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
    opt.step()
    torch.cuda.synchronize()
    after = {n: p.detach().clone() for n, p in m.named_parameters()}
    return float(loss.detach()), grads, before, after


@pytest.mark.parametrize("tag,fd,B,n,T", [("a", 64, 2, 16000, 24), ("b", 128, 1, 16000, 20)])
def test_train_step_fp32_vs_reference_golden(golden, tag, fd, B, n, T):
    g = golden("train")
    loss, grads, before, after = run_case("fp32", fd, B, n, T)
    np.testing.assert_allclose(loss, g[f"{tag}_loss"][0], rtol=1e-4)
    names = [str(x) for x in g[f"{tag}_names"]]
    worst = 0.0
    for nme in names:
        ref = g[f"{tag}_g/{nme}"]
        scale = max(np.abs(ref).max(), 1e-12)
        if nme not in grads:                                   # parameters the step never touches: the reference leaves zeros / None
            assert np.abs(ref).max() == 0.0, nme
            continue
        got = sub(grads[nme])
        if exactly_zero(ref, got):
            continue
        err = np.abs(got - ref).max() / scale
        worst = max(worst, err)
        assert err <= 5e-4, (nme, err, scale)
        dp = sub(after[nme] - before[nme])
        ok = np.abs(ref) > 1e-4 * scale + 1e-7                  # Adam's first step is lr * g / (|g| + eps): skip the eps floor
        np.testing.assert_allclose(dp[ok], g[f"{tag}_dp/{nme}"][ok], atol=3e-7, rtol=0, err_msg=nme)
    print(f"fp32 train step {tag}: loss {loss:.6e}, worst gradient max-err / max|ref| = {worst:.2e}")


def test_train_step_bf16_vs_reference_golden(golden):
    g = golden("train")
    tag, fd, B, n, T = "a", 64, 2, 16000, 24
    loss, grads, before, after = run_case("bf16", fd, B, n, T)
    np.testing.assert_allclose(loss, g[f"{tag}_loss"][0], rtol=1e-2)
    worst = ("", 0.0)
    for nme in (str(x) for x in g[f"{tag}_names"]):
        if nme not in grads:
            continue
        ref = g[f"{tag}_g/{nme}"]
        got = sub(grads[nme])
        if exactly_zero(ref, got):
            continue
        rel = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-20)
        if rel > worst[1]:
            worst = (nme, rel)
        assert rel <= 5e-2, (nme, rel)
    print(f"bf16 train step: loss {loss:.6e}, worst gradient rel-L2 {worst[1]:.2e} ({worst[0]})")


def test_train_step_config5_shape_vs_oracle():
    """BASELINE configs[4] clip: 4 s of audio, 120 frames at 30 fps (frame_num = 120), batch 1, fd 64; ground-truth vertices given
    directly (the VOCASET flavour). Checked live against torch autograd over the CPU oracle."""
    from oracle import train_oracle as to
    fd, B, n, T = 64, 1, 64000, 120
    sd_w2v = synth.wav2vec2_state(0)
    sd_ff = synth.faceformer_state(fd=fd, seed=264, variant="vert")
    template = synth.flame_buffers()["v_template"].reshape(1, 1, 15069)
    gt = template + 1e-3 * torch.from_numpy(np.random.default_rng(5).normal(size=(B, T, 15069)).astype(np.float32))
    audio = synth.audio(B, n, seed=99)
    losses, grads, _ = to.train_step(sd_ff, sd_w2v, template, audio, gt, lr=1e-4)
    m = build_vert("fp32", fd, 264)
    opt = train.FlatAdam(m, lr=1e-4)
    opt.zero_grad()
    loss = m.training_loss(audio.cuda(), gt.cuda())
    loss.backward()
    np.testing.assert_allclose(float(loss), losses[0], rtol=1e-4)
    named = dict(m.named_parameters())
    for nme, ref in grads.items():
        if nme == "audio_encoder.masked_spec_embed" or named[nme].grad is None:
            continue
        got = named[nme].grad.cpu()
        if exactly_zero(ref.numpy(), got.numpy()):
            continue
        err = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)
        assert err <= 1e-3, (nme, err)


def test_three_adam_steps_track_the_oracle_and_refresh_inference_packs():
    """Three fused-Adam steps (fp32 mode) follow torch.optim.Adam over the CPU oracle step for step (moment state, bias correction),
    and predict() afterwards sees the updated weights (operand-pack invalidation through ops.WEIGHT_EPOCH)."""
    from oracle import train_oracle as to
    fd, B, n, T = 64, 2, 16000, 24
    sd_w2v = synth.wav2vec2_state(0)
    sd_ff = synth.faceformer_state(fd=fd, seed=264, variant="vert")
    template = synth.flame_buffers()["v_template"].reshape(1, 1, 15069)
    gt = template + 1e-3 * torch.from_numpy(np.random.default_rng(6).normal(size=(B, T, 15069)).astype(np.float32))
    audio = synth.audio(B, n, seed=98)
    ref_losses, _, ref_after = to.train_step(sd_ff, sd_w2v, template, audio, gt, lr=1e-4, steps=3)
    m = build_vert("fp32", fd, 264)
    audio, gt = audio.cuda(), gt.cuda()
    v0 = m.predict_from_embeddings(audio).clone()
    opt = train.FlatAdam(m, lr=1e-4)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        loss = m.training_loss(audio, gt)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    np.testing.assert_allclose(losses, ref_losses, rtol=2e-3)
    named = dict(m.named_parameters())
    for nme in ("vertice_map_r.weight", "audio_feature_map.weight", "audio_encoder.encoder.layers.5.feed_forward.output_dense.weight",
                "transformer_decoder.layers.0.linear1.weight"):
        d = (named[nme].detach().cpu() - ref_after[nme]).abs().max().item()
        assert d <= 1.5e-4, (nme, d)            # 3 steps of lr 1e-4: sign flips of near-zero gradients cost at most ~1 lr each
    v1 = m.predict_from_embeddings(audio)
    assert (v1 - v0).abs().max().item() > 0.0


def test_graphed_step_matches_eager():
    """The CUDA-graph replay of forward + backward gives the eager step's loss and gradients (atomics reorder sums only)."""
    fd, B, n, T = 64, 2, 16000, 24
    template = synth.flame_buffers()["v_template"].reshape(1, 1, 15069)
    gt = (template + 1e-3 * torch.from_numpy(np.random.default_rng(6).normal(size=(B, T, 15069)).astype(np.float32))).cuda()
    audio = synth.audio(B, n, seed=98).cuda()
    res = []
    for graphed in (False, True):
        m = build_vert("bf16", fd, 264)
        opt = train.FlatAdam(m, lr=1e-4)
        gstep = train.GraphedTrainStep(m, audio.shape, gt.shape) if graphed else None
        losses = []
        for _ in range(3):
            if graphed:
                loss = gstep(audio, gt)
            else:
                opt.zero_grad()
                loss = m.training_loss(audio, gt)
                loss.backward()
            losses.append(float(loss.detach()))
            opt.step()
        torch.cuda.synchronize()
        res.append((losses, m._flat_params.clone()))
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=2e-3)
    assert (res[0][1] - res[1][1]).abs().max().item() <= 2.5e-4      # 3 Adam steps of lr 1e-4: sign noise of ~zero gradients only


# ------------------------------------------------------------------------------------------------ TRAIN mode (regularisers as inputs)
def run_reg_case(precision, fd=64, B=2, n_samples=16000, T=24):
    m = build_vert(precision, fd, 200 + fd)
    coeff, pose, shape, mean, std = train_inputs(B, T, seed=90 + fd)
    m.coeff_mean, m.coeff_std = mean.cuda(), std.cuda()
    audio = synth.audio(B, n_samples, seed=4321).cuda()
    reg = synth.train_regularisers(B, T, fd, seed=300)
    with torch.no_grad():                                                        # faceformer_vert.py:405-412
        gt = (m.convert_coeff2verts(coeff.cuda()[:, :, :53].reshape(-1, 53), pose.cuda().reshape(-1, 6),
                                    torch.zeros(B * T, shape.shape[-1], device="cuda")) * m.vertice_scale).reshape(B, T, -1)
    opt = train.FlatAdam(m, lr=1e-4)
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    opt.zero_grad()
    loss = m.training_loss(audio, gt, reg=reg)
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
    opt.step()
    torch.cuda.synchronize()
    after = {n: p.detach().clone() for n, p in m.named_parameters()}
    return float(loss.detach()), grads, before, after, reg


def test_train_mode_step_fp32_vs_reference_golden(golden):
    """Dropout (wav2vec2 hidden / attention / activation, PPE, the five sites of nn.TransformerDecoderLayer), SpecAugment and LayerDrop
    ACTIVE, every draw an input: tests/golden/train_reg.npz is the reference's own forward_switch_frame in .train() mode with the same
    draws injected (oracle/make_golden.golden_train_reg). Loss rel 1e-4, every gradient max|err| <= 5e-4 max|ref|, the dropped layers'
    gradients exactly zero, masked_spec_embed's gradient present, Adam update within 3e-7."""
    g = golden("train_reg")
    loss, grads, before, after, reg = run_reg_case("fp32")
    np.testing.assert_allclose(loss, g["loss"][0], rtol=1e-4)
    worst = ("", 0.0)
    for nme in (str(x) for x in g["names"]):
        ref = g[f"g/{nme}"]
        scale = max(np.abs(ref).max(), 1e-12)
        if nme not in grads:
            assert np.abs(ref).max() == 0.0, nme
            continue
        got = sub(grads[nme])
        if exactly_zero(ref, got):
            continue
        err = np.abs(got - ref).max() / scale
        if err > worst[1]:
            worst = (nme, err)
        assert err <= 5e-4, (nme, err, scale)
        dp = sub(after[nme] - before[nme])
        ok = np.abs(ref) > 1e-4 * scale + 1e-7
        np.testing.assert_allclose(dp[ok], g[f"dp/{nme}"][ok], atol=3e-7, rtol=0, err_msg=nme)
    dropped = [l for l, k in enumerate(reg["layer_keep"]) if not k]
    assert dropped
    for l in dropped:
        for nme, t in grads.items():
            if nme.startswith(f"audio_encoder.encoder.layers.{l}."):
                assert not torch.any(t), nme
    assert grads["audio_encoder.masked_spec_embed"].abs().max().item() > 0
    print(f"fp32 TRAIN-mode step: loss {loss:.6e}, worst gradient max-err / max|ref| = {worst[1]:.2e} ({worst[0]})")


def test_train_mode_step_bf16_vs_reference_golden(golden):
    g = golden("train_reg")
    loss, grads, _, _, _ = run_reg_case("bf16")
    np.testing.assert_allclose(loss, g["loss"][0], rtol=1e-2)
    worst = ("", 0.0)
    for nme in (str(x) for x in g["names"]):
        if nme not in grads:
            continue
        ref = g[f"g/{nme}"]
        got = sub(grads[nme])
        if exactly_zero(ref, got):
            continue
        rel = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-20)
        if rel > worst[1]:
            worst = (nme, rel)
        assert rel <= 5e-2, (nme, rel)
    print(f"bf16 TRAIN-mode step: loss {loss:.6e}, worst gradient rel-L2 {worst[1]:.2e} ({worst[0]})")


def test_graphed_train_mode_step_takes_new_draws_without_recapture():
    """GraphedTrainStep with regularisers given as tensors: the draws live in static buffers and LayerDrop is a 0 / 1 blend inside the
    graph; different draws over the same set of sites replay ONE graph, and each replay equals the eager step (which skips dropped
    layers) on the same draw."""
    fd, B, n, T = 64, 2, 16000, 24
    template = synth.flame_buffers()["v_template"].reshape(1, 1, 15069)
    gt = (template + 1e-3 * torch.from_numpy(np.random.default_rng(6).normal(size=(B, T, 15069)).astype(np.float32))).cuda()
    audio = synth.audio(B, n, seed=98).cuda()
    m = build_vert("fp32", fd, 264)
    train.flatten_parameters(m)
    gstep = train.GraphedTrainStep(m, audio.shape, gt.shape)
    regs = [synth.train_regularisers(B, T, fd, seed=s) for s in (300, 301)] + [synth.train_regularisers(B, T, fd, seed=302, drop_layers=(1,))]
    for i, reg in enumerate(regs):
        lg = float(gstep(audio, gt, reg=reg))
        gg = m._flat_grad.clone()
        loss = m.training_loss(audio, gt, reg=reg)
        loss.backward()
        assert abs(lg - float(loss)) <= 1e-5 * abs(float(loss)), (i, lg, float(loss))
        scale = m._flat_grad.abs().max().item()
        assert (gg - m._flat_grad).abs().max().item() <= 1e-5 * scale, i
        # synth.train_regularisers has no draws for the layers it drops, so the third dict activates a different set of sites: one more
        # graph (DeviceDraws, which draws every site every step, always replays one)
        assert len(gstep.graphs) == (1 if i < 2 else 2)
        for l, k in enumerate(reg["layer_keep"]):
            if not k:
                p = f"audio_encoder.encoder.layers.{l}.feed_forward.output_dense.weight"
                assert not torch.any(m._flat_layout.view(gg, p)), p


def test_drawn_regularisers_follow_the_config_and_the_step_runs():
    """train.draw_regularisers on the device: keep rates ~ 1 - p, masks pre-scaled, SpecAugment spans of mask_time_length, and
    `model.regularisers = "draw"` makes .train() mode stochastic (two calls differ) while .eval() stays deterministic."""
    fd, B, n, T = 64, 2, 64000, 100
    m = build_vert("bf16", fd, 264)
    cfg = m.audio_encoder.config
    gen = torch.Generator(device="cuda").manual_seed(5)
    reg = train.draw_regularisers(B, T, fd, cfg, torch.device("cuda"), generator=gen)
    assert set(reg["masks"]) == set(train.regulariser_shapes(B, T, fd, cfg))
    mk = reg["masks"]["l0.act"]
    assert abs((mk > 0).float().mean().item() - (1 - cfg.activation_dropout)) < 0.01
    assert abs(mk.max().item() - 1 / (1 - cfg.activation_dropout)) < 1e-6
    assert reg["spec_mask"].shape == (B, T) and reg["spec_mask"].sum(1).min().item() >= cfg.mask_time_length
    assert len(reg["layer_keep"]) == cfg.num_hidden_layers
    template = synth.flame_buffers()["v_template"].reshape(1, 1, 15069)
    gt = (template + 1e-3 * torch.from_numpy(np.random.default_rng(6).normal(size=(B, T, 15069)).astype(np.float32))).cuda()
    audio = synth.audio(B, n, seed=98).cuda()
    m.regularisers = "draw"
    m.train()
    l1, l2 = float(m.training_loss(audio, gt)), float(m.training_loss(audio, gt))
    assert np.isfinite(l1) and np.isfinite(l2) and l1 != l2
    m.eval()
    e1, e2 = float(m.training_loss(audio, gt)), float(m.training_loss(audio, gt))
    assert e1 == e2


def test_device_draws_bit_exact_vs_philox_oracle():
    """csrc/train_draw.cu against oracle/philox_oracle.py: every dropout mask element, the LayerDrop decisions / blend rows and the
    SpecAugment rows of two consecutive steps, bit for bit (integer generator + one fp32 compare)."""
    from transformers import Wav2Vec2Config
    from oracle import philox_oracle as po
    cfg = Wav2Vec2Config()
    cfg.layerdrop = 0.3                                                       # so that both outcomes occur among 12 layers
    B, T, fd, seed = 2, 24, 64, (9 << 32) + 1234
    d = train.DeviceDraws(B, T, fd, cfg, "cuda", seed=seed)
    for step in range(2):
        reg = d.draw()
        torch.cuda.synchronize()
        assert d.state.tolist()[2] == step + 1
        for gi, (p, start, n) in enumerate(d.groups):
            want = po.dropout_masks(n, p, seed, step, stream_id=gi)
            assert np.array_equal(d.flat[start:start + n].cpu().numpy(), want), (step, gi)
        keep = po.layer_keep(cfg.num_hidden_layers, cfg.layerdrop, seed, step)
        assert np.array_equal(d.keep_flags.cpu().numpy() > 0, keep) and 0 < keep.sum() < len(keep)
        bl = d.blend.cpu().numpy()
        assert np.array_equal(bl[0], np.repeat(keep[:, None], d.rows, 1).astype(np.float32)) and np.array_equal(bl[1], 1 - bl[0])
        want_spec = po.spec_mask(B, T, d.span_len, d.span_rate, d.min_spans, seed, step)
        assert np.array_equal(d.spec.cpu().numpy().reshape(B, T), want_spec) and want_spec.any()
        ca = reg["masks"]["dec.ca"]
        assert torch.equal(reg["ca_diag"], torch.diagonal(ca, dim1=2, dim2=3).permute(0, 2, 1))


def test_graph_with_device_draws_equals_eager_on_the_same_stream_of_draws():
    """The draw launches are captured inside the step's graph: every replay sees the next step of the Philox stream, and the losses /
    gradients equal the eager step fed by a second DeviceDraws with the same seed (which also checks the blend form of LayerDrop
    against itself across launch modes)."""
    fd, B, n, T = 64, 2, 16000, 24
    template = synth.flame_buffers()["v_template"].reshape(1, 1, 15069)
    gt = (template + 1e-3 * torch.from_numpy(np.random.default_rng(6).normal(size=(B, T, 15069)).astype(np.float32))).cuda()
    audio = synth.audio(B, n, seed=98).cuda()
    m = build_vert("fp32", fd, 264)
    train.flatten_parameters(m)
    cfg = m.audio_encoder.config
    d_graph, d_eager = (train.DeviceDraws(B, T, fd, cfg, "cuda", seed=42) for _ in range(2))
    m.training_loss(audio, gt, reg=train.DeviceDraws(B, T, fd, cfg, "cuda", seed=1).draw()).backward()   # one-time kernel attributes
    gstep = train.GraphedTrainStep(m, audio.shape, gt.shape, warmup=0)        # no warm-up draws: replay i is step i of the stream
    losses = []
    for i in range(3):
        lg = float(gstep(audio, gt, reg=d_graph))
        gg = m._flat_grad.clone()
        if i == 0:                    # the capture itself consumed step 0 without executing; the first replay runs it
            assert d_graph.state.tolist()[2] == 1
        loss = m.training_loss(audio, gt, reg=d_eager.draw())
        loss.backward()
        assert abs(lg - float(loss)) <= 1e-5 * abs(float(loss)), (i, lg, float(loss))
        assert (gg - m._flat_grad).abs().max().item() <= 1e-5 * m._flat_grad.abs().max().item(), i
        losses.append(lg)
    assert len(set(losses)) == 3 and len(gstep.graphs) == 1
