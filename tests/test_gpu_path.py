"""End-to-end parity of the CUDA hot path against the oracle and the committed golden vectors
(outputs of the reference's own classes, tests/golden/, see oracle/make_golden.py).

Tolerances (BASELINE.json north_star): fp32 mode <= 1e-5 (max-abs vertex error in metres, coefficient relative error);
bf16 mode: max-abs vertex error <= 1e-4 m, coefficient (hidden-state) relative error <= 1e-2."""
import numpy as np
import pytest
import torch

from avi_talking_b200 import synth
from oracle import faceformer_oracle as ffo
from oracle import flame_oracle as fo
from oracle import wav2vec2_oracle as wo
from oracle.make_golden import COL_STRIDE

from helpers import build_faceformer, build_flame, build_wav2vec

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return ((a - b).norm() / b.norm()).item()


# ------------------------------------------------------------------------------------------------ FLAME
@pytest.mark.parametrize("n_shape,tag", [(100, "a"), (300, "b")])
def test_flame_matches_reference_golden(golden, n_shape, tag):
    g = golden("flame")
    m = build_flame(n_shape, mediapipe=(n_shape == 100))
    p = {k: v.cuda() for k, v in synth.flame_params(4, n_shape=n_shape, seed=3).items()}
    res = m(p["shape"], p["exp"], p["pose"], p["eye"])
    assert np.abs(res[0].cpu().numpy() - g[f"verts_{tag}"]).max() < 1e-6          # metres
    assert np.abs(res[1].cpu().numpy() - g[f"lmk2d_{tag}"]).max() < 1e-6
    assert np.abs(res[2].cpu().numpy() - g[f"lmk3d_{tag}"]).max() < 1e-6
    if n_shape == 100:
        assert np.abs(res[3].cpu().numpy() - g["lmkmp_a"]).max() < 1e-6
    pose = p["pose"].clone()
    pose[:, :3] = 0
    v = m(p["shape"], p["exp"], pose)[0]
    assert np.abs(v.cpu().numpy() - g[f"verts_jawonly_{tag}"]).max() < 1e-6


def test_flame_lbs_function_and_identities(golden, monkeypatch):
    from avi_talking_b200.flame import lbs
    monkeypatch.setenv("AVI_B200_PRECISION", "fp32")   # the free function has no module to carry the mode
    g = golden("flame")
    buf = {k: v.cuda() for k, v in synth.flame_buffers(100, 50).items()}
    p = synth.flame_params(2, seed=5)
    betas = torch.cat([p["shape"], p["exp"]], 1).cuda()
    full_pose = torch.cat([p["pose"][:, :3], torch.zeros(2, 3), p["pose"][:, 3:], p["eye"]], 1).cuda()
    v, J = lbs(betas, full_pose, buf["v_template"][None].expand(2, -1, -1), buf["shapedirs"], buf["posedirs"],
               buf["J_regressor"], buf["parents"], buf["lbs_weights"])
    assert np.abs(v.cpu().numpy() - g["gdl_lbs_verts"]).max() < 1e-6
    assert np.abs(J.cpu().numpy() - g["gdl_lbs_joints"]).max() < 1e-6
    # zero betas + zero pose => template ; zero pose => pure blendshape
    z = torch.zeros(3, 150, device="cuda")
    v, _ = lbs(z, torch.zeros(3, 15, device="cuda"), buf["v_template"], buf["shapedirs"], buf["posedirs"], buf["J_regressor"],
               buf["parents"], buf["lbs_weights"])
    assert (v - buf["v_template"][None]).abs().max().item() < 3e-7
    v, _ = lbs(betas, torch.zeros(2, 15, device="cuda"), buf["v_template"], buf["shapedirs"], buf["posedirs"],
               buf["J_regressor"], buf["parents"], buf["lbs_weights"])
    want = buf["v_template"] + torch.einsum("bl,mkl->bmk", betas, buf["shapedirs"])
    assert (v - want).abs().max().item() < 1e-6
    monkeypatch.setenv("AVI_B200_PRECISION", "bf16")   # tcgen05 blend (fp16 operands): 5e-5 m
    v, J = lbs(betas, full_pose, buf["v_template"], buf["shapedirs"], buf["posedirs"], buf["J_regressor"], buf["parents"],
               buf["lbs_weights"])
    assert np.abs(v.cpu().numpy() - g["gdl_lbs_verts"]).max() < 5e-5
    assert np.abs(J.cpu().numpy() - g["gdl_lbs_joints"]).max() < 1e-6


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-6), ("bf16", 5e-5)])
def test_flame_many_frames_and_template_mutation(precision, tol):
    """Ragged frame/vertex tiles (F=333 is not a multiple of 64) and the in-place v_template mutation callers perform.
    bf16 mode = tcgen05 blend with fp16 operands (tolerance 5e-5 m, inside the 1e-4 m budget); fp32 mode = CUDA-core blend."""
    m = build_flame(100, mediapipe=False, precision=precision)
    buf = synth.flame_buffers(100, 50)
    p = synth.flame_params(333, seed=11)
    v = m.vertices_only(p["shape"].cuda(), p["exp"].cuda(), p["pose"].cuda(), p["eye"].cuda())
    ref = fo.flame_forward(buf, p["shape"], p["exp"], p["pose"], p["eye"])[0]
    err = (v.cpu() - ref).abs().max().item()
    print(f"FLAME {precision} F=333 max abs vertex error {err:.3e} m")
    assert err < tol
    if precision == "bf16":   # a multi-chunk case: > 16 frame tiles per vertex tile, > 148 work items
        p2 = synth.flame_params(2500, seed=12)
        v2 = m.vertices_only(p2["shape"].cuda(), p2["exp"].cuda(), p2["pose"].cuda(), p2["eye"].cuda())
        ref2 = fo.flame_forward(buf, p2["shape"], p2["exp"], p2["pose"], p2["eye"])[0]
        assert (v2.cpu() - ref2).abs().max().item() < tol
    m.v_template.add_(0.01)    # TalkingHeadWrapper.py:140-158 style mutation must invalidate the packed cache
    buf["v_template"] = buf["v_template"] + 0.01
    v = m.vertices_only(p["shape"][:5].cuda(), p["exp"][:5].cuda(), p["pose"][:5].cuda(), p["eye"][:5].cuda())
    ref = fo.flame_forward(buf, p["shape"][:5], p["exp"][:5], p["pose"][:5], p["eye"][:5])[0]
    assert (v.cpu() - ref).abs().max().item() < tol


# ------------------------------------------------------------------------------------------------ wav2vec2
def test_wav2vec2_fp32_matches_reference_golden(golden):
    g = golden("w2v")
    m = build_wav2vec("fp32")
    a1 = synth.audio(2, 16000, seed=1234).cuda()
    hs = m(a1, "vocaset").last_hidden_state
    assert hs.shape == (2, 24, 768)
    ref = torch.from_numpy(g["hs_1s"])
    assert relerr(hs.cpu(), ref) < 1e-5 and (hs.cpu() - ref).abs().max().item() < 1e-4
    hs = m(a1, "vocaset", frame_num=20).last_hidden_state
    assert relerr(hs.cpu(), torch.from_numpy(g["hs_1s_frame20"])) < 1e-5
    a4 = synth.audio(1, 64000, seed=1234).cuda()
    hs = m(a4, "vocaset").last_hidden_state
    assert hs.shape == (1, 99, 768)
    assert relerr(hs.cpu(), torch.from_numpy(g["hs_4s"])) < 1e-5


def test_wav2vec2_bf16_matches_reference_golden(golden):
    g = golden("w2v")
    m = build_wav2vec("bf16")
    a1 = synth.audio(2, 16000, seed=1234).cuda()
    hs = m(a1, "vocaset").last_hidden_state
    e1 = relerr(hs.cpu(), torch.from_numpy(g["hs_1s"]))
    a4 = synth.audio(1, 64000, seed=1234).cuda()
    e4 = relerr(m(a4, "vocaset").last_hidden_state.cpu(), torch.from_numpy(g["hs_4s"]))
    print("bf16 wav2vec2 relative error:", e1, e4)
    assert e1 < 1e-2 and e4 < 1e-2


def test_wav2vec2_feature_extractor_stage(golden):
    g = golden("w2v")
    for prec, tol in (("fp32", 1e-5), ("bf16", 1e-2)):
        m = build_wav2vec(prec)
        a1 = synth.audio(2, 16000, seed=1234).cuda()
        feats, T50, La = m._feature_extractor(a1, m._pack())
        got = feats[:, :T50].float().cpu().transpose(1, 2)
        assert T50 == 49
        assert relerr(got, torch.from_numpy(g["feats_1s"])) < tol


# ------------------------------------------------------------------------------------------------ decoder
@pytest.mark.parametrize("fd", [64, 128, 256])
def test_decoder_ar_matches_oracle(fd):
    """forward_ff autoregressive branch on given hidden states, batched over 3 clips, against the literal O(T^2) oracle."""
    m = build_faceformer("fp32", fd=fd, seed=10 + fd)
    sd = synth.faceformer_state(fd=fd, seed=10 + fd)
    rng = np.random.default_rng(5)
    B, T = 3, 40
    hs = torch.from_numpy(rng.normal(size=(B, T, 36 + fd)).astype(np.float32))
    obj = torch.from_numpy((0.3 * rng.normal(size=(B, fd))).astype(np.float32))
    template = synth.flame_buffers()["v_template"].reshape(1, 1, 15069)
    ref = ffo.forward_ff(sd, template, hs, obj, T, teacher_forcing=False)
    got = m.forward_ff(None, hs.cuda(), obj.cuda(), T, teacher_forcing=False)
    assert (got.cpu() - ref).abs().max().item() < 1e-5      # metres
    disp_ref = ref - template
    assert relerr(got.cpu() - template, disp_ref) < 1e-4


@pytest.mark.parametrize("fd", [64, 128, 256])
def test_decoder_teacher_forced_matches_oracle(fd):
    m = build_faceformer("fp32", fd=fd, seed=10 + fd)
    sd = synth.faceformer_state(fd=fd, seed=10 + fd)
    rng = np.random.default_rng(6)
    B, T = 2, 33
    hs = torch.from_numpy(rng.normal(size=(B, T, 36 + fd)).astype(np.float32))
    obj = torch.from_numpy((0.3 * rng.normal(size=(B, fd))).astype(np.float32))
    template = synth.flame_buffers()["v_template"].reshape(1, 1, 15069)
    gt = template + 1e-3 * torch.from_numpy(rng.normal(size=(B, T, 15069)).astype(np.float32))
    ref = ffo.forward_ff(sd, template, hs, obj, T, teacher_forcing=True, gt_verts=gt)
    got = m.forward_ff(gt.cuda(), hs.cuda(), obj.cuda(), T, teacher_forcing=True)
    assert (got.cpu() - ref).abs().max().item() < 1e-5


# ------------------------------------------------------------------------------------------------ predict end to end
@pytest.mark.parametrize("fd", [64, 128])
def test_predict_fp32_matches_reference_golden(golden, fd):
    g = golden("faceformer")
    m = build_faceformer("fp32", fd=fd, seed=10 + fd)
    a = synth.audio(1, 16000, seed=1234).cuda()
    emo = synth.fan_embeddings(24, seed=20)["emo"][None].cuda()
    v = m.predict_from_embeddings(a, emo)
    assert v.shape == (1, 24, 15069)
    ref = torch.from_numpy(g[f"predict_fd{fd}_sub"])
    err = (v[0, :, ::COL_STRIDE].cpu() - ref).abs().max().item()
    print("fp32 predict max abs vertex error (m):", err)
    assert err < 1e-5


def test_predict_c1_fp32_and_bf16(golden):
    """BASELINE config 1 (one 4 s clip, T=99) against the reference's own predict()."""
    g = golden("faceformer")
    ref = torch.from_numpy(g["predict_c1_sub"])
    a = synth.audio(1, 64000, seed=1234).cuda()
    emo = synth.fan_embeddings(99, seed=20)["emo"][None].cuda()
    template = synth.flame_buffers()["v_template"].reshape(-1)[::COL_STRIDE]
    for prec, tol_v, tol_rel in (("fp32", 1e-5, 1e-4), ("bf16", 1e-4, 1e-2)):
        m = build_faceformer(prec, fd=64, seed=74)
        v = m.predict_from_embeddings(a, emo)
        got = v[0, :, ::COL_STRIDE].cpu()
        err = (got - ref).abs().max().item()
        rel = relerr(got - template, ref - template)
        print(f"{prec} C1 predict: max abs vertex error {err:.3e} m, displacement relative error {rel:.3e}")
        assert err < tol_v and rel < tol_rel


def test_predict_api_with_fan_stub(golden):
    """The reference predict(audio, head_img, eye_img, emotion_img) signature, FanEncoder replaced by a stub."""
    g = golden("faceformer")
    emb = synth.fan_embeddings(24, seed=20)

    class Fan(torch.nn.Module):
        """A per-image function (what FanEncoder is in eval mode): the frame index travels in pixel [0,0,0]; the emotion embedding
        is scaled by (1 + sum of the mouth rows), which the reference's mask_lip (:119-133, :791) zeroes before the encoder."""
        calls = 0

        def forward(self, img):
            Fan.calls += 1
            i = img.reshape(img.shape[0], -1)[:, 0].round().long().cpu()
            lower = 1.0 + img[:, :, int(100. / 224. * img.shape[2]):, :].sum(dim=(1, 2, 3))
            return emb["head"][i].cuda(), emb["eye"][i].cuda(), emb["emo"][i].cuda() * lower[:, None], None

    m = build_faceformer("fp32", fd=64, seed=74)
    m.fan_net = Fan().eval()
    frames = torch.zeros(24, 3, 4, 4, device="cuda")
    frames[:, 0, 0, 0] = torch.arange(24).float()
    frames[:, :, 1:, :] = 0.37                                        # a non-zero mouth region: mask_lip must remove it
    a = synth.audio(1, 16000, seed=1234).cuda()
    v = m.predict(a, frames, frames, frames)
    v2 = m.predict_from_embeddings(a, emb["emo"][None].cuda())
    assert torch.equal(v, v2) and Fan.calls == 1                     # ONE batched encoder call
    # ... and it is the reference's own predict() output on the same frames (golden minted with the real mask_lip)
    # (same weights seed 10 + 64 = 74, audio and embeddings as oracle/make_golden.golden_faceformer's fd = 64 case)
    assert np.abs(v[0, :, ::7].cpu().numpy() - g["predict_fd64_sub"]).max() < 1e-5
    assert frames[:, :, 1:, :].min().item() == float(np.float32(0.37))                  # the caller's frames are not modified (masked copy)
    # a 5-frame emotion clip played ping-pong over the 24 output frames (loop_utils.loopback_frames, golden index pattern):
    # batched-unique path (eval) == upstream's frame-by-frame path (taken for a train-mode provider)
    idx = torch.from_numpy(g["loop_idx_5_17"]).long()
    from avi_talking_b200.loop_utils import calc_loop_idx
    assert [calc_loop_idx(i, 5) for i in range(17)] == idx.tolist()
    Fan.calls = 0
    v_eval = m.predict(a, frames[:5], frames[:5], frames[:5])
    assert Fan.calls == 1
    m.fan_net.train()
    v_train = m.predict(a, frames[:5], frames[:5], frames[:5])
    assert Fan.calls == 1 + 24 and torch.equal(v_eval, v_train)
    want = emb["emo"][torch.tensor([calc_loop_idx(i, 5) for i in range(24)])][None].cuda()
    assert torch.equal(v_eval, m.predict_from_embeddings(a, want))


def test_batched_predict_equals_per_clip():
    """Batching over clips is new behaviour (the reference loops clip by clip): it must not change any clip's result."""
    m = build_faceformer("bf16", fd=64, seed=74)
    a = synth.audio(3, 16000, seed=99).cuda()
    emo = torch.from_numpy(np.random.default_rng(3).normal(size=(3, 24, 30)).astype(np.float32)).cuda()
    vb = m.predict_from_embeddings(a, emo)
    for c in range(3):
        vc = m.predict_from_embeddings(a[c:c + 1], emo[c:c + 1])
        assert (vb[c] - vc[0]).abs().max().item() < 2e-6


@pytest.mark.gpu
def test_predict_and_convert_overlapped_equals_sequential():
    """The side-stream FLAME launch (sized to the SMs the AR decoder leaves idle) changes scheduling only."""
    from avi_talking_b200.smoke import build_models
    m = build_models("bf16")
    B, n = 3, 16000
    a = synth.audio(B, n, seed=5).cuda()
    T = 24
    emo = torch.stack([synth.fan_embeddings(T, seed=30 + c)["emo"] for c in range(B)]).cuda()
    rng = np.random.default_rng(3)
    coeff = torch.from_numpy(rng.standard_normal((B * T, 53)).astype(np.float32)).cuda()
    pose = torch.from_numpy((0.1 * rng.standard_normal((B * T, 6))).astype(np.float32)).cuda()
    shape = torch.from_numpy(rng.standard_normal((B * T, 100)).astype(np.float32)).cuda()
    v0 = m.predict_from_embeddings(a, emo)
    f0 = m.convert_coeff2verts(coeff, pose.clone(), shape)
    for _ in range(3):
        v1, f1 = m.predict_and_convert(a, emo, coeff, pose.clone(), shape)
    torch.cuda.synchronize()
    assert torch.equal(v0, v1) and torch.equal(f0, f1)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-6), ("bf16", 5e-5)])
def test_convert_coeff2verts_matches_oracle(precision, tol):
    """Faceformer.convert_coeff2verts (faceformer_disentangle.py:425-433: de-normalise the 53 coefficients, zero the global pose in
    place, FLAME with exp[:50] and the jaw) against the CPU oracle restatement; per-frame shapes and the hoisted per-clip shape."""
    from oracle import flame_oracle as fo

    from avi_talking_b200.smoke import build_models
    m = build_models(precision)
    rng = np.random.default_rng(11)
    F_ = 77
    coeff = torch.from_numpy(rng.standard_normal((F_, 53)).astype(np.float32))
    pose = torch.from_numpy(np.concatenate([0.3 * rng.standard_normal((F_, 3)), 0.1 * rng.standard_normal((F_, 3))], 1).astype(np.float32))
    shape = torch.from_numpy(rng.standard_normal((F_, 100)).astype(np.float32))
    buf = synth.flame_buffers(100, 50)
    want = fo.convert_coeff2verts(buf, m.coeff_mean.reshape(-1).cpu(), m.coeff_std.reshape(-1).cpu(), coeff, pose.clone(), shape)
    pose_dev = pose.cuda()
    got = m.convert_coeff2verts(coeff.cuda(), pose_dev, shape.cuda())
    assert float(pose_dev[:, :3].abs().max()) == 0.0 and torch.equal(pose_dev[:, 3:].cpu(), pose[:, 3:])   # in place, as upstream (:429)
    err = (got.reshape(F_, -1).cpu() - want.reshape(F_, -1)).abs().max().item()
    print(f"convert_coeff2verts {precision}: max abs vertex error {err:.2e} m")
    assert err < tol


@pytest.mark.gpu
def test_linear_interpolation_function():
    """models/lib/wav2vec.py:67-73 as a standalone function (inside forward it is fused with the LayerNorm)."""
    import torch.nn.functional as F

    from avi_talking_b200.wav2vec import linear_interpolation
    x = torch.from_numpy(np.random.default_rng(5).normal(size=(3, 199, 512)).astype(np.float32))
    for out_len in (None, 97, 250):
        got = linear_interpolation(x.cuda(), 50, 25, output_len=out_len)
        n = out_len if out_len is not None else int(199 / 50.0 * 25)
        want = F.interpolate(x.transpose(1, 2), size=n, align_corners=True, mode="linear").transpose(1, 2)
        assert got.shape == want.shape and (got.cpu() - want).abs().max().item() < 1e-5


@pytest.mark.gpu
def test_graphed_predict_and_convert_matches_eager():
    """CUDA-graph replay of predict_and_convert (avi_talking_b200/graphs.py) returns what the eager call returns, also after the
    inputs change and after the weights change (re-capture keyed on parameter versions)."""
    from avi_talking_b200.smoke import build_models
    m = build_models("bf16")
    B, n, T = 3, 16000, 24
    g = torch.Generator().manual_seed(5)
    for trial in range(3):
        audio = synth.audio(B, n, seed=300 + trial).cuda()
        emo = torch.randn(B, T, 30, generator=g).cuda()
        coeff = torch.randn(B * T, 53, generator=g).cuda()
        pose = (0.1 * torch.randn(B * T, 6, generator=g)).cuda()
        shape = torch.randn(B * T, 100, generator=g).cuda()
        if trial == 2:
            with torch.no_grad():
                m.audio_feature_map.bias.add_(0.01)          # bumps _version: the graph must be re-captured with new packs
        v0, f0 = m.predict_and_convert(audio, emo, coeff, pose.clone(), shape)
        v1, f1 = m.graphed_predict_and_convert(audio, emo, coeff, pose.clone(), shape)
        assert torch.equal(v0, v1) and torch.equal(f0, f1)


@pytest.mark.gpu
def test_two_batches_in_flight_on_two_graph_instances():
    """graphed_predict_and_convert(slot=k): independent graph instances (own activation pools, own static outputs) replayed
    concurrently on two streams - what bench.py --inflight 2 times - each return their own batch's eager result."""
    from avi_talking_b200.smoke import build_models
    m = build_models("bf16")
    B, n, T = 4, 16000, 24
    g = torch.Generator().manual_seed(8)
    batches, want = [], []
    for k in range(2):
        audio = synth.audio(B, n, seed=400 + k).cuda()
        emo = torch.randn(B, T, 30, generator=g).cuda()
        coeff = torch.randn(B * T, 53, generator=g).cuda()
        pose = (0.1 * torch.randn(B * T, 6, generator=g)).cuda()
        shape = torch.randn(B * T, 100, generator=g).cuda()
        batches.append((audio, emo, coeff, pose, shape))
        v, f = m.predict_and_convert(audio, emo, coeff, pose.clone(), shape)
        want.append((v.clone(), f.clone()))
    lanes = [torch.cuda.Stream() for _ in range(2)]
    for rep in range(3):                                   # first pass captures, the next ones replay concurrently
        got = []
        for k in range(2):
            lanes[k].wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(lanes[k]):
                a, e, c, p, s_ = batches[k]
                got.append(m.graphed_predict_and_convert(a, e, c, p.clone(), s_, slot=k))
        for k in range(2):
            torch.cuda.current_stream().wait_stream(lanes[k])
        torch.cuda.synchronize()
        for k in range(2):
            assert torch.equal(got[k][0], want[k][0]) and torch.equal(got[k][1], want[k][1]), (rep, k)


@pytest.mark.gpu
def test_long_and_ragged_clips_take_the_general_kernels():
    """Clips beyond the fast paths' limits: 12 s of audio -> T = 299 frames (> 256: generic AR decoder and generic attention), a
    batch of 3 such clips, and the shortest clip the conv stack admits one frame for. fp32 mode against the CPU oracle."""
    n = 192000
    a = synth.audio(3, n, seed=77)
    T = 299
    emo = torch.stack([synth.fan_embeddings(T, seed=30 + c)["emo"] for c in range(3)])
    sd_w2v, sd_ff = synth.wav2vec2_state(0), synth.faceformer_state(fd=64, seed=74)
    template = synth.flame_buffers()["v_template"].reshape(1, 1, 15069)
    ref = ffo.predict(sd_ff, sd_w2v, template, a[:1], emo[:1], cached=True)
    assert ref.shape[1] == T
    m = build_faceformer("fp32", fd=64, seed=74)
    v = m.predict_from_embeddings(a.cuda(), emo.cuda())
    assert tuple(v.shape) == (3, T, 15069)
    assert (v[:1].cpu() - ref).abs().max().item() < 1e-5
    m16 = build_faceformer("bf16", fd=64, seed=74)
    v16 = m16.predict_from_embeddings(a.cuda(), emo.cuda())
    assert (v16[:1].cpu() - ref).abs().max().item() < 1.5e-4      # 299 sequential decoder steps on bf16 audio features
    assert (v16[1:] - v[1:]).abs().max().item() < 1.5e-4
    # one second and a bit: T = 4 frames; and a clip so short that the encoder yields a single frame
    for n_s, T_s in ((3000, 4), (1000, 1)):
        a_s = synth.audio(2, n_s, seed=78)
        emo_s = torch.zeros(2, max(T_s, 1), 30)
        ref_s = ffo.predict(sd_ff, sd_w2v, template, a_s, emo_s, cached=True)
        assert ref_s.shape[1] == T_s
        v_s = m.predict_from_embeddings(a_s.cuda(), emo_s.cuda())
        assert (v_s.cpu() - ref_s).abs().max().item() < 1e-5


@pytest.mark.gpu
def test_full_size_configs1_properties():
    """BASELINE configs[1] at full size (64 clips x 10 s = 15 936 frames, bf16 mode), through size-independent properties:
    (1) clips are independent: rows of the batched result equal the same clips run alone (first, middle, last);
    (2) FLAME linearity in the blend: with zero pose the skinning is the identity, so verts(shape, exp) - template is additive in
        (shape, exp) and the zero-coefficient mesh is the template;
    (3) every output is finite and the vertex head adds the template exactly once (mean displacement << template scale);
    (4) one clip checked against the CPU oracle end to end (1e-4 m)."""
    from avi_talking_b200.smoke import build_models
    from bench import make_inputs, n_frames
    m = build_models("bf16")
    B, n = 64, 160000
    T = n_frames(n)
    assert T == 249
    inp = {k: v.cuda() for k, v in make_inputs(B, n, T, seed=1000).items()}
    v, fv = m.predict_and_convert(inp["audio"], inp["emo"], inp["coeff"], inp["pose"].clone(), inp["shape"])
    assert tuple(v.shape) == (B, T, 15069) and tuple(fv.shape)[0] == B * T
    assert torch.isfinite(v).all() and torch.isfinite(fv).all()
    for c in (0, 31, 63):                                                                                   # (1)
        vc = m.predict_from_embeddings(inp["audio"][c:c + 1], inp["emo"][c:c + 1])
        assert (v[c] - vc[0]).abs().max().item() < 2e-6, c
    template = m.template.reshape(-1).to(v.device)
    disp = v - template
    assert disp.abs().mean().item() < 0.05 and disp.abs().max().item() < 1.0                                # (3)
    flame = m.flame
    F_ = 512                                                                                                 # (2)
    g = torch.Generator().manual_seed(2)
    s1, e1 = torch.randn(F_, 100, generator=g).cuda(), torch.randn(F_, 50, generator=g).cuda()
    s2, e2 = torch.randn(F_, 100, generator=g).cuda(), torch.randn(F_, 50, generator=g).cuda()
    zp = torch.zeros(F_, 6, device="cuda")
    tpl = flame.v_template.reshape(1, -1)
    d = lambda s, e: flame.vertices_only(shape_params=s, expression_params=e, pose_params=zp).reshape(F_, -1) - tpl  # noqa: E731
    z = torch.zeros_like(s1), torch.zeros_like(e1)
    assert d(*z).abs().max().item() < 1e-6
    assert (d(s1 + s2, e1 + e2) - d(s1, e1) - d(s2, e2)).abs().max().item() < 2e-4       # fp16 operand rounding of the blend, 3 terms
    sd_w2v, sd_ff = synth.wav2vec2_state(0), synth.faceformer_state(fd=64, seed=74)                         # (4)
    ref = ffo.predict(sd_ff, sd_w2v, synth.flame_buffers()["v_template"].reshape(1, 1, 15069), inp["audio"][5:6].cpu(),
                      inp["emo"][5:6].cpu(), cached=True)
    assert (v[5:6].cpu() - ref).abs().max().item() < 1e-4


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-6), ("bf16", 5e-5)])
def test_lbs_pose2rot_false_matches_reference_golden(golden, monkeypatch, precision, tol):
    """lbs(pose2rot=False): the pose argument is the stack of rotation matrices (lbs.py:205-209); golden from both reference copies
    (tests/golden/lbs_rotmat.npz), every joint rotated."""
    from avi_talking_b200.flame import lbs
    monkeypatch.setenv("AVI_B200_PRECISION", precision)
    g = golden("lbs_rotmat")
    buf = {k: v.cuda() for k, v in synth.flame_buffers(100, 50).items()}
    p = synth.flame_params(3, seed=7)
    betas = torch.cat([p["shape"], p["exp"]], 1).cuda()
    rot = torch.from_numpy(g["rot"]).cuda()
    v, J = lbs(betas, rot, buf["v_template"], buf["shapedirs"], buf["posedirs"], buf["J_regressor"], buf["parents"], buf["lbs_weights"],
               pose2rot=False)
    assert np.abs(v.cpu().numpy() - g["gdl_verts"]).max() < tol
    assert np.abs(J.cpu().numpy() - g["gdl_joints"]).max() < 1e-6
    with pytest.raises(ValueError):
        lbs(betas, rot[:, :4], buf["v_template"], buf["shapedirs"], buf["posedirs"], buf["J_regressor"], buf["parents"],
            buf["lbs_weights"], pose2rot=False)
