"""The CPU oracle against the golden vectors minted from the reference's own classes
(oracle/make_golden.py).  This is what pins the oracle (prompt section 3)."""
import numpy as np
import torch

from oracle import faceformer_oracle as ffo
from oracle import flame_oracle as fo
from oracle import synth
from oracle import wav2vec2_oracle as wo
from oracle.make_golden import COL_STRIDE, checksum


def test_flame_matches_reference(golden):
    g = golden("flame")
    for n_shape, tag in ((100, "a"), (300, "b")):
        buf = synth.flame_buffers(n_shape, 50)
        p = synth.flame_params(4, n_shape=n_shape, seed=3)
        res = fo.flame_forward(buf, p["shape"], p["exp"], p["pose"], p["eye"], mediapipe=(n_shape == 100))
        np.testing.assert_allclose(res[0].numpy(), g[f"verts_{tag}"], atol=1e-6, rtol=0)
        np.testing.assert_allclose(res[1].numpy(), g[f"lmk2d_{tag}"], atol=1e-6, rtol=0)
        np.testing.assert_allclose(res[2].numpy(), g[f"lmk3d_{tag}"], atol=1e-6, rtol=0)
        if n_shape == 100:
            np.testing.assert_allclose(res[3].numpy(), g["lmkmp_a"], atol=1e-6, rtol=0)
        pose = p["pose"].clone()
        pose[:, :3] = 0
        v = fo.flame_forward(buf, p["shape"], p["exp"], pose)[0]
        np.testing.assert_allclose(v.numpy(), g[f"verts_jawonly_{tag}"], atol=1e-6, rtol=0)


def test_gdl_lbs_copy_matches(golden):
    g = golden("flame")
    buf = synth.flame_buffers(100, 50)
    p = synth.flame_params(2, seed=5)
    betas = torch.cat([p["shape"], p["exp"]], 1)
    full_pose = torch.cat([p["pose"][:, :3], torch.zeros(2, 3), p["pose"][:, 3:], p["eye"]], 1)
    v, J = fo.lbs(betas, full_pose, buf["v_template"], buf["shapedirs"], buf["posedirs"], buf["J_regressor"],
                  buf["parents"], buf["lbs_weights"])
    np.testing.assert_allclose(v.numpy(), g["gdl_lbs_verts"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(J.numpy(), g["gdl_lbs_joints"], atol=1e-6, rtol=0)


def test_flame_identities():
    """Known-answer cases: zero everything => template; zero pose => pure blendshape."""
    buf = synth.flame_buffers(100, 50)
    z = torch.zeros(2, 100)
    v = fo.flame_forward(buf, z, torch.zeros(2, 50), torch.zeros(2, 6))[0]
    np.testing.assert_allclose(v.numpy(), buf["v_template"][None].expand(2, -1, -1).numpy(), atol=2e-7)
    p = synth.flame_params(2, seed=9)
    v = fo.flame_forward(buf, p["shape"], p["exp"], torch.zeros(2, 6))[0]
    want = buf["v_template"] + torch.einsum("bl,mkl->bmk", torch.cat([p["shape"], p["exp"]], 1), buf["shapedirs"])
    np.testing.assert_allclose(v.numpy(), want.numpy(), atol=5e-7)


def test_wav2vec2_matches_reference(golden):
    g = golden("w2v")
    sd = synth.wav2vec2_state(0)
    a1 = synth.audio(2, 16000, seed=1234)
    np.testing.assert_allclose(wo.feature_extractor(sd, a1).numpy(), g["feats_1s"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(wo.wav2vec2_forward(sd, a1).numpy(), g["hs_1s"], atol=3e-5, rtol=0)
    np.testing.assert_allclose(wo.wav2vec2_forward(sd, a1, frame_num=20).numpy(), g["hs_1s_frame20"], atol=3e-5, rtol=0)
    a4 = synth.audio(1, 64000, seed=1234)
    assert wo.output_frames(64000) == 99 and wo.output_frames(160000) == 249
    np.testing.assert_allclose(wo.wav2vec2_forward(sd, a4).numpy(), g["hs_4s"], atol=3e-5, rtol=0)


def test_masks_and_ppe_match_reference(golden):
    g = golden("faceformer")
    m = ffo.init_biased_mask(4, 600, 30)
    np.testing.assert_array_equal(m[:, :64, :64].numpy(), g["biased_mask_p30"])
    s = torch.nan_to_num(ffo.init_biased_mask(4, 600, 25), neginf=0.0).double().sum().item()
    assert s == g["biased_mask_p25_full_sum"][0]
    np.testing.assert_array_equal(ffo.enc_dec_mask("vocaset", 5, 7).numpy(), g["enc_dec_mask_voca"])
    np.testing.assert_array_equal(ffo.ppe_table(64, 30)[0, :70].numpy(), g["ppe_fd64_p30"])
    # closed form used by the CUDA kernels: mask[h,i,j] = -slope_h * floor((i-j)/period), j<=i
    i = torch.arange(600)[:, None]
    j = torch.arange(600)[None]
    slopes = torch.tensor(ffo.get_slopes(4))
    closed = torch.where(j <= i, -slopes[:, None, None] * ((i - j) // 30).float()[None], torch.tensor(float("-inf")))
    assert torch.equal(closed, m)
    assert ffo.get_slopes(4) == [2 ** -2, 2 ** -4, 2 ** -6, 2 ** -8]


def _ff_inputs(fd, seed, n_samples, T):
    sd_w2v = synth.wav2vec2_state(0)
    sd_ff = synth.faceformer_state(fd=fd, seed=seed)
    template = synth.flame_buffers()["v_template"].reshape(1, 1, 15069)
    a = synth.audio(1, n_samples, seed=1234)
    emb = synth.fan_embeddings(T, seed=20)
    return sd_w2v, sd_ff, template, a, emb


def test_predict_matches_reference(golden):
    g = golden("faceformer")
    for fd in (64, 128):
        sd_w2v, sd_ff, template, a, emb = _ff_inputs(fd, 10 + fd, 16000, 24)
        v = ffo.predict(sd_ff, sd_w2v, template, a, emb["emo"][None])
        np.testing.assert_allclose(v[0, :, ::COL_STRIDE].numpy(), g[f"predict_fd{fd}_sub"], atol=2e-6, rtol=0)
        np.testing.assert_allclose(checksum(v), g[f"predict_fd{fd}_chk"], rtol=1e-5)
        # KV-cached O(T) restatement == literal O(T^2) loop
        vc = ffo.predict(sd_ff, sd_w2v, template, a, emb["emo"][None], cached=True)
        np.testing.assert_allclose(vc.numpy(), v.numpy(), atol=2e-6, rtol=0)


def test_teacher_forced_matches_reference(golden):
    g = golden("faceformer")
    for fd in (64, 128):
        sd_w2v, sd_ff, template, a, emb = _ff_inputs(fd, 10 + fd, 16000, 24)
        T = 24
        ha = wo.wav2vec2_forward(sd_w2v, a)
        ha = torch.nn.functional.linear(ha, sd_ff["audio_feature_map.weight"], sd_ff["audio_feature_map.bias"])
        hs = torch.cat([sd_ff["learnable_eye_embed"].expand(1, T, -1), emb["emo"][None], ha], -1)
        obj = torch.nn.functional.linear(torch.eye(8)[:1], sd_ff["obj_vector.weight"])
        gt = template + 1e-3 * torch.from_numpy(
            np.random.default_rng(77).normal(size=(1, T, 15069)).astype(np.float32))
        vt = ffo.forward_ff(sd_ff, template, hs, obj, T, teacher_forcing=True, gt_verts=gt)
        np.testing.assert_allclose(vt[0, :, ::COL_STRIDE].numpy(), g[f"tf_fd{fd}_sub"], atol=2e-6, rtol=0)


def test_predict_c1_matches_reference(golden):
    """BASELINE config 1: one 4 s clip, batch 1, fd=64 (T=99), via the O(T) restatement."""
    g = golden("faceformer")
    sd_w2v, sd_ff, template, a, emb = _ff_inputs(64, 74, 64000, 99)
    v = ffo.predict(sd_ff, sd_w2v, template, a, emb["emo"][None], cached=True)
    np.testing.assert_allclose(v[0, :, ::COL_STRIDE].numpy(), g["predict_c1_sub"], atol=3e-6, rtol=0)


def test_loopback_frames(golden):
    from avi_talking_b200.loop_utils import calc_loop_idx
    g = golden("faceformer")
    assert [calc_loop_idx(i, 5) for i in range(17)] == list(g["loop_idx_5_17"])


def _train_case(tag, fd, B, n_samples, T):
    from oracle.make_golden import train_inputs
    sd_w2v = synth.wav2vec2_state(0)
    sd_ff = synth.faceformer_state(fd=fd, seed=200 + fd, variant="vert")
    coeff, pose, shape, mean, std = train_inputs(B, T, seed=90 + fd)
    buf = synth.flame_buffers(100, 50)
    gt = fo.convert_coeff2verts(buf, mean.reshape(-1), std.reshape(-1), coeff.reshape(-1, 53), pose.reshape(-1, 6).clone(),
                                torch.zeros(B * T, 100)).reshape(B, T, 15069)               # faceformer_vert.py:408-412
    audio = synth.audio(B, n_samples, seed=4321)
    return sd_ff, sd_w2v, buf["v_template"].reshape(1, 1, 15069), audio, gt


def sub(t, stride):
    t = t.reshape(-1)
    return (t[::stride] if t.numel() > 4096 else t).numpy()


def test_train_step_matches_reference(golden):
    """oracle/train_oracle.py against the reference's own forward_switch_frame + backward + Adam (tests/golden/train.npz)."""
    from oracle import train_oracle as to
    from oracle.make_golden import GRAD_STRIDE
    g = golden("train")
    for tag, fd, B, n, T in (("a", 64, 2, 16000, 24), ("b", 128, 1, 16000, 20)):
        sd_ff, sd_w2v, template, audio, gt = _train_case(tag, fd, B, n, T)
        before = to.trainable(sd_ff, sd_w2v)
        losses, grads, after = to.train_step(sd_ff, sd_w2v, template, audio, gt, lr=1e-4)
        np.testing.assert_allclose(losses[0], g[f"{tag}_loss"][0], rtol=2e-5)
        names = [str(x) for x in g[f"{tag}_names"]]
        assert set(names) - {"audio_encoder.masked_spec_embed"} == set(grads)
        for nme in grads:
            ref = g[f"{tag}_g/{nme}"]
            got = sub(grads[nme], GRAD_STRIDE)
            scale = max(np.abs(ref).max(), 1e-12)
            assert np.abs(got - ref).max() <= 2e-4 * scale + 1e-9, (nme, np.abs(got - ref).max(), scale)
            dp = sub(after[nme] - before[nme], GRAD_STRIDE)
            # Adam's first step is lr * g / (|g| + eps): compare where the reference's |g| is not at the eps floor
            ok = np.abs(ref) > 1e-6 * scale + 1e-7
            np.testing.assert_allclose(dp[ok], g[f"{tag}_dp/{nme}"][ok], atol=2e-7, rtol=0, err_msg=nme)


def test_train_step_in_train_mode_matches_reference(golden):
    """The same step with dropout / SpecAugment / LayerDrop ACTIVE (tests/golden/train_reg.npz: the reference in .train() mode with every
    draw injected from synth.train_regularisers): loss, every gradient - masked_spec_embed included, dropped layers exactly zero -
    and the Adam update."""
    from oracle import train_oracle as to
    from oracle.make_golden import GRAD_STRIDE
    g = golden("train_reg")
    fd, B, n, T = 64, 2, 16000, 24
    sd_ff, sd_w2v, template, audio, gt = _train_case("a", fd, B, n, T)
    reg = synth.train_regularisers(B, T, fd, seed=300)
    assert [str(x) for x in g["sites_in_call_order"]][:len(reg["order"]) - 7] == reg["order"][:-7]     # encoder sites, then per clip
    before = to.trainable(sd_ff, sd_w2v, spec_augment=True)
    losses, grads, after = to.train_step(sd_ff, sd_w2v, template, audio, gt, lr=1e-4, reg=reg)
    np.testing.assert_allclose(losses[0], g["loss"][0], rtol=2e-5)
    assert {str(x) for x in g["names"]} == set(grads)
    for nme in grads:
        ref = g[f"g/{nme}"]
        got = sub(grads[nme], GRAD_STRIDE)
        scale = max(np.abs(ref).max(), 1e-12)
        assert np.abs(got - ref).max() <= 2e-4 * scale + 1e-9, (nme, np.abs(got - ref).max(), scale)
        ok = np.abs(ref) > 1e-6 * scale + 1e-7
        np.testing.assert_allclose(sub(after[nme] - before[nme], GRAD_STRIDE)[ok], g[f"dp/{nme}"][ok], atol=2e-7, rtol=0, err_msg=nme)
    dropped = [l for l, k in enumerate(reg["layer_keep"]) if not k]
    assert dropped and all(not np.any(g[f"g/audio_encoder.encoder.layers.{l}.attention.out_proj.weight"]) for l in dropped)
    assert np.abs(g["g/audio_encoder.masked_spec_embed"]).max() > 0


def test_frontend_matches_reference_bit_exact(golden):
    """avi_talking_b200.frontend (host-side framing of the audio, integer work) against the reference's own process_audio /
    create_base_sample compiled from source (oracle/make_golden.golden_frontend): bit-exact, including upstream's np.pad quirk."""
    from avi_talking_b200 import frontend as fe
    from oracle.make_golden import frontend_wav
    g = golden("frontend")
    cases = [(16000 * 4 + 123, {}), (16000 * 2, dict(smallest_unit=8)), (9999, dict(silent_frames_start=3, silent_frames_end=2)),
             (640 * 7, dict(smallest_unit=4, silence_all=True)), (100, {})]
    for i, (n, kw) in enumerate(cases):
        wav = frontend_wav(n, 700 + i)
        pa = fe.process_audio(wav, 16000, 25)
        assert pa["raw_audio"].dtype == np.int16 and np.array_equal(pa["raw_audio"], g[f"pa_{i}"])
        s = fe.create_base_sample(wav, **kw)
        assert np.array_equal(s["raw_audio"], g[f"cbs_{i}_raw"])
        for k in ("gt_exp", "gt_shape", "gt_jaw", "gt_tex"):
            assert tuple(s[k].shape) == tuple(g[f"cbs_{i}_{k}_shape"]) and not s[k].any()


def test_clip_text_oracle_matches_transformers(golden):
    """oracle/clip_oracle.py against transformers.CLIPTextModel (the class models/diffusion_prior.py:37 instantiates)."""
    from oracle import clip_oracle as co
    g = golden("clip_text")
    for tag, layers, B in (("l12", 12, 3), ("l2", 2, 2)):
        sd = synth.clip_text_state(60, layers)
        ids = synth.clip_tokens(B, seed=61)
        with torch.no_grad():
            last = co.clip_text_forward(sd, ids, layers)
        np.testing.assert_allclose(last[:, ::4, ::3].numpy(), g[f"{tag}_last_sub"], atol=2e-5, rtol=0)
        np.testing.assert_allclose(last.mean(dim=1).numpy(), g[f"{tag}_voxel"], atol=1e-5, rtol=0)


def test_get_subject_labels_matches_reference():
    """TalkingHeadWrapper.get_subject_labels (host logic) against lists minted from the reference's own method compiled from source
    (tests/golden/subject_labels.json, oracle/make_golden.golden_subject_labels; TalkingHeadWrapper.py:168-236)."""
    import json
    import os
    import types
    from avi_talking_b200.talking_head import TalkingHeadWrapper, emote_cfg
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "subject_labels.json")))
    for key, want in ref.items():
        split, which = key.split("/")
        cfg = emote_cfg()
        cfg.data.split = split
        assert TalkingHeadWrapper.get_subject_labels(types.SimpleNamespace(cfg=cfg), which) == want


def test_fan_encoder_oracle_matches_reference(golden):
    """oracle/fan_oracle.py against the reference's own FanEncoder (tests/golden/fan.npz, oracle/make_golden.golden_fan)."""
    from oracle import fan_oracle as fo2
    g = golden("fan")
    out = fo2.fan_encoder_forward(synth.fan_state(80), synth.fan_images(3, seed=81))
    for k, t in zip(("head", "eye", "emo", "mouth"), out):
        np.testing.assert_allclose(t.numpy(), g[k], atol=2e-4, rtol=1e-5)


def test_philox_known_answers():
    """oracle/philox_oracle.py (the checker of the device-side dropout / LayerDrop / SpecAugment draws) against the known-answer vectors
    of Philox4x32-10 from the Random123 distribution (kat_vectors: zero, all-ones and the pi-digits counter / key)."""
    from oracle import philox_oracle as po
    kat = [((0, 0, 0, 0), (0, 0), "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
           ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, "408f276d 41c83b0e a20bc7c6 6d5451fd"),
           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), "d16cfe09 94fdcceb 5001e420 24126ea1")]
    for ctr, key, want in kat:
        assert " ".join(f"{int(x):08x}" for x in po.philox4x32_10(ctr, key)) == want
    m = po.dropout_masks(1 << 16, 0.1, seed=77, step=3)
    assert set(np.unique(m)) == {np.float32(0.0), np.float32(1.0) / (np.float32(1.0) - np.float32(0.1))}
    assert abs((m > 0).mean() - 0.9) < 5e-3
    assert not np.array_equal(m, po.dropout_masks(1 << 16, 0.1, seed=77, step=4))            # a new step is a new draw
    assert not np.array_equal(m, po.dropout_masks(1 << 16, 0.1, seed=77, step=3, stream_id=1))
    sp = po.spec_mask(4, 120, 10, 0.6, 2, seed=77, step=0)
    assert sp.shape == (4, 120) and sp.sum(1).min() >= 10 and sp.sum(1).max() <= 20


def test_device_draws_layout():
    """train.DeviceDraws (host logic only, CPU tensors): every dropout site of the step is a view of ONE flat buffer, 16-byte aligned,
    disjoint, grouped by probability; the site list is the reference's (synth.train_regularisers minus LayerDrop's omissions)."""
    from transformers import Wav2Vec2Config
    from avi_talking_b200 import train
    cfg = Wav2Vec2Config()
    B, T, fd = 2, 24, 64
    d = train.DeviceDraws(B, T, fd, cfg, "cpu", seed=(5 << 32) + 7)
    ref = synth.train_regularisers(B, T, fd, drop_layers=())
    assert set(d.masks) == set(ref["masks"]) and all(tuple(d.masks[k].shape) == tuple(ref["masks"][k].shape) for k in ref["masks"])
    spans = sorted((v.data_ptr() - d.flat.data_ptr(), v.numel() * 4) for v in d.masks.values())
    assert all(a % 16 == 0 for a, _ in spans)
    assert all(spans[i][0] + spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))
    assert spans[-1][0] + spans[-1][1] <= d.flat.numel() * 4
    assert len(d.groups) == 1 and abs(d.groups[0][0] - 0.1) < 1e-9          # HF defaults: every site at p = 0.1
    assert d.state.tolist()[:3] == [7, 5, 0]
    assert d.blend.shape == (2, cfg.num_hidden_layers, B * T * cfg.hidden_size) and d.spec.numel() == B * T


def test_lbs_with_rotation_matrices_matches_reference(golden):
    """lbs(pose2rot=False) (lbs.py:205-209: the pose argument is the stack of rotation matrices): oracle vs both reference copies."""
    g = golden("lbs_rotmat")
    buf = synth.flame_buffers(100, 50)
    p = synth.flame_params(3, seed=7)
    betas = torch.cat([p["shape"], p["exp"]], 1)
    v, J = fo.lbs(betas, torch.from_numpy(g["rot"]), buf["v_template"], buf["shapedirs"], buf["posedirs"], buf["J_regressor"],
                  buf["parents"], buf["lbs_weights"], pose2rot=False)
    assert np.abs(v.numpy() - g["gdl_verts"]).max() < 1e-6 and np.abs(J.numpy() - g["gdl_joints"]).max() < 1e-6
    assert np.array_equal(g["gdl_verts"], g["inferno_verts"])


def test_mask_builders_match_reference(golden):
    """init_biased_mask / enc_dec_mask (BIWI and vocaset) of the oracle AND of the drop-in module (host-side closed forms, kept for
    API compatibility; the kernels build the same bias on the fly) against the reference's own builders (tests/golden/masks.npz)."""
    from avi_talking_b200 import faceformer as ff
    g = golden("masks")
    for impl in (ffo, ff):
        biwi = impl.enc_dec_mask("BIWI", 5, 10) if impl is ffo else impl.enc_dec_mask("cpu", "BIWI", 5, 10)
        voca = impl.enc_dec_mask("vocaset", 7, 7) if impl is ffo else impl.enc_dec_mask("cpu", "vocaset", 7, 7)
        assert np.array_equal(biwi.numpy(), g["edm_biwi_5_10"]) and np.array_equal(voca.numpy(), g["edm_vocaset_7_7"])
        for heads, period, L in ((4, 30, 64), (4, 25, 60)):
            m = impl.init_biased_mask(heads, L, period).numpy()
            ref = g[f"bias_h{heads}_p{period}_L{L}"]
            assert np.array_equal(np.isinf(m), np.isinf(ref))
            fin = ~np.isinf(ref)
            assert np.abs(m[fin] - ref[fin]).max() == 0.0


def test_host_helpers_match_reference(golden):
    """Pure host-side helpers of the drop-in (CPU-executable data movement / index math) against the reference's own functions
    (tests/golden/host_helpers.npz): the ping-pong frame indexer, mask_lip on a square and a non-square frame, the output length of
    linear_interpolation."""
    from avi_talking_b200 import loop_utils
    from avi_talking_b200.faceformer import mask_lip
    from avi_talking_b200.wav2vec import linear_interpolation_length
    g = golden("host_helpers")
    for n, frames in ((5, 24), (1, 7), (3, 3), (4, 33)):
        assert [loop_utils.calc_loop_idx(i, n) for i in range(frames)] == g[f"loop_{n}_{frames}"].tolist()
        img = torch.arange(n, dtype=torch.float32)[:, None].repeat(1, 2)
        assert np.array_equal(loop_utils.loopback_frames(img, frames).numpy(), g[f"loopback_{n}_{frames}"])
    for tag in ("sq", "nonsq"):
        x = torch.from_numpy(g[f"masklip_in_{tag}"])
        assert np.array_equal(mask_lip(x).numpy(), g[f"masklip_out_{tag}"])
        assert np.array_equal(x.numpy(), g[f"masklip_in_{tag}"])                       # the input is not modified
    for t50, l25, l30, l20 in g["lerp_lengths"].tolist():
        assert linear_interpolation_length(t50, 50, 25) == l25 and linear_interpolation_length(t50, 50, 30) == l30
        assert linear_interpolation_length(t50, 50, 25, output_len=20) == l20
