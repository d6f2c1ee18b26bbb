"""Mint golden fixtures from the reference's OWN Python classes (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python -m oracle.make_golden            # writes tests/golden/*.npz

Each fixture stores only OUTPUTS (inputs/weights are regenerated from oracle/synth.py seeds on
both sides) plus the seeds used.  What is imported from the reference, unmodified:
  * inferno.models.DecaFLAME.FLAME / FLAME_mediapipe and inferno.utils.lbs (third_party/inferno)
  * gdl.utils.lbs.lbs (BlendshapeVisualizer/EMOCA copy)
  * models.lib.wav2vec.Wav2Vec2Model (subclass of the installed transformers Wav2Vec2Model)
  * models.faceformer_disentangle: init_biased_mask, enc_dec_mask, PeriodicPositionalEncoding and
    Faceformer.predict / forward_ff, reached through ``Faceformer.__new__`` because __init__ needs
    network access, licensed FLAME assets and private paths.  Modules the import pulls in
    that are absent here (easydict, omegaconf, pytorch3d-based visualisers, DECA, pirender) are stubbed
    with MagicMock; none of them is touched by predict/forward_ff.
  * loop_utils.loopback_frames
"""
from __future__ import annotations

import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
COL_STRIDE = 7  # vertex-coordinate subsampling for the big [T,15069] outputs


def _paths():
    for p in (REF, os.path.join(REF, "third_party", "inferno"), os.path.join(REF, "BlendshapeVisualizer", "EMOCA")):
        if p not in sys.path:
            sys.path.insert(0, p)


def checksum(t: torch.Tensor) -> np.ndarray:
    d = t.double()
    return np.array([d.sum().item(), (d * d).sum().item(), d.abs().max().item()])


def golden_flame():
    from inferno.models.DecaFLAME import FLAME, FLAME_mediapipe
    from . import synth
    out = {}
    for n_shape, tag in ((100, "a"), (300, "b")):
        cfg = synth.write_flame_assets("/tmp/avi_flame_assets")
        cfg.n_shape = n_shape
        m = FLAME_mediapipe(cfg) if n_shape == 100 else FLAME(cfg)
        p = synth.flame_params(4, n_shape=n_shape, seed=3)
        with torch.no_grad():
            res = m(p["shape"], p["exp"], p["pose"], p["eye"])
            # Path-A call convention: pose=[0,0,0,jaw], default eyes (faceformer_disentangle.py:425-433)
            pose_a = p["pose"].clone()
            pose_a[:, :3] = 0
            res_a = m(p["shape"], p["exp"], pose_a)
        out[f"verts_{tag}"] = res[0].numpy()
        out[f"lmk2d_{tag}"] = res[1].numpy()
        out[f"lmk3d_{tag}"] = res[2].numpy()
        if n_shape == 100:
            out["lmkmp_a"] = res[3].numpy()
        out[f"verts_jawonly_{tag}"] = res_a[0].numpy()
    # the gdl copy of lbs() called directly
    from gdl.utils.lbs import lbs as gdl_lbs
    buf = synth.flame_buffers(100, 50)
    p = synth.flame_params(2, seed=5)
    betas = torch.cat([p["shape"], p["exp"]], 1)
    full_pose = torch.cat([p["pose"][:, :3], torch.zeros(2, 3), p["pose"][:, 3:], p["eye"]], 1)
    v, J = gdl_lbs(betas, full_pose, buf["v_template"][None].expand(2, -1, -1), buf["shapedirs"], buf["posedirs"],
                   buf["J_regressor"], buf["parents"], buf["lbs_weights"], detach_pose_correctives=False)
    out["gdl_lbs_verts"] = v.numpy()
    out["gdl_lbs_joints"] = J.numpy()
    np.savez(os.path.join(GOLD, "flame.npz"), **out)
    print("flame.npz", {k: v.shape for k, v in out.items()})


def golden_lbs_rotmat():
    """lbs(pose2rot=False) of BOTH reference copies (gdl and inferno, lbs.py:205-209): the pose argument is the stack of rotation
    matrices. Separate small fixture (tests/golden/lbs_rotmat.npz) so that flame.npz stays byte-identical."""
    from gdl.utils.lbs import batch_rodrigues as gdl_rodrigues
    from gdl.utils.lbs import lbs as gdl_lbs
    from inferno.utils.lbs import lbs as inf_lbs
    from . import synth
    buf = synth.flame_buffers(100, 50)
    p = synth.flame_params(3, seed=7)
    betas = torch.cat([p["shape"], p["exp"]], 1)
    full_pose = torch.cat([p["pose"][:, :3], 0.1 * p["pose"][:, :3], p["pose"][:, 3:], p["eye"]], 1)      # every joint rotated
    rot = gdl_rodrigues(full_pose.view(-1, 3)).view(3, 5, 3, 3)
    out = {"rot": rot.numpy()}
    for tag, fn in (("gdl", gdl_lbs), ("inferno", inf_lbs)):
        v, J = fn(betas, rot, buf["v_template"][None].expand(3, -1, -1), buf["shapedirs"], buf["posedirs"], buf["J_regressor"],
                  buf["parents"], buf["lbs_weights"], pose2rot=False)
        out[f"{tag}_verts"], out[f"{tag}_joints"] = v.numpy(), J.numpy()
    assert np.array_equal(out["gdl_verts"], out["inferno_verts"])
    np.savez_compressed(os.path.join(GOLD, "lbs_rotmat.npz"), **out)
    print("lbs_rotmat.npz", {k: v.shape for k, v in out.items()})


def golden_masks():
    """The reference's own mask builders at sizes / datasets the other fixtures do not cover: enc_dec_mask for BIWI (two visible memory
    frames per query, faceformer_disentangle.py:83-85) and vocaset, init_biased_mask at two (heads, period) settings. Separate small
    fixture (tests/golden/masks.npz) so that faceformer.npz stays byte-identical."""
    ffd = _import_faceformer()
    out = {"edm_biwi_5_10": ffd.enc_dec_mask("cpu", "BIWI", 5, 10).numpy(), "edm_vocaset_7_7": ffd.enc_dec_mask("cpu", "vocaset", 7, 7).numpy()}
    for heads, period, L in ((4, 30, 64), (4, 25, 60)):
        out[f"bias_h{heads}_p{period}_L{L}"] = ffd.init_biased_mask(n_head=heads, max_seq_len=L, period=period).numpy()
    np.savez_compressed(os.path.join(GOLD, "masks.npz"), **out)
    print("masks.npz", {k: v.shape for k, v in out.items()})


def golden_host_helpers():
    """Pure host-side helpers of the path, from the reference's own functions: the ping-pong frame indexer (loop_utils.py:4-16),
    mask_lip on a square and a NON-square frame (faceformer_disentangle.py:119-133: upstream scales the row range by shape[3] and the
    column range by shape[2]), and the output length of linear_interpolation (models/lib/wav2vec.py:67-73)."""
    ffd = _import_faceformer()
    import loop_utils as ref_loop
    import models.lib.wav2vec as ref_w2v
    out = {}
    for n, frames in ((5, 24), (1, 7), (3, 3), (4, 33)):
        out[f"loop_{n}_{frames}"] = np.array([ref_loop.calc_loop_idx(i, n) for i in range(frames)])
        img = torch.arange(n, dtype=torch.float32)[:, None].repeat(1, 2)
        out[f"loopback_{n}_{frames}"] = ref_loop.loopback_frames(img, frames).numpy()
    g = torch.Generator().manual_seed(4)
    for tag, shp in (("sq", (2, 3, 56, 56)), ("nonsq", (1, 3, 48, 60))):
        x = torch.rand(shp, generator=g)
        out[f"masklip_in_{tag}"] = x.numpy()
        out[f"masklip_out_{tag}"] = ffd.mask_lip(x).numpy()
    lens = []
    for t50 in (49, 199, 499, 3, 2, 77):
        f = torch.zeros(1, t50, 4)
        lens.append([t50, ref_w2v.linear_interpolation(f, 50, 25).shape[1], ref_w2v.linear_interpolation(f, 50, 30).shape[1],
                     ref_w2v.linear_interpolation(f, 50, 25, output_len=20).shape[1]])
    out["lerp_lengths"] = np.array(lens)
    np.savez_compressed(os.path.join(GOLD, "host_helpers.npz"), **out)
    print("host_helpers.npz", {k: v.shape for k, v in out.items()})


def _ref_wav2vec2(sd):
    from transformers import Wav2Vec2Config
    from models.lib.wav2vec import Wav2Vec2Model
    m = Wav2Vec2Model(Wav2Vec2Config(attn_implementation="eager")).eval()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return m


def golden_wav2vec2():
    from . import synth
    sd = synth.wav2vec2_state(0)
    m = _ref_wav2vec2(sd)
    out = {}
    with torch.no_grad():
        a1 = synth.audio(2, 16000, seed=1234)
        out["hs_1s"] = m(a1, "vocaset").last_hidden_state.numpy()                 # [2,24,768]
        out["hs_1s_frame20"] = m(a1, "vocaset", frame_num=20).last_hidden_state.numpy()
        a4 = synth.audio(1, 64000, seed=1234)
        out["hs_4s"] = m(a4, "vocaset").last_hidden_state.numpy()                 # [1,99,768] (config C1)
        feats = m.feature_extractor(a1)                                            # [2,512,49]
        out["feats_1s"] = feats.numpy()
    np.savez(os.path.join(GOLD, "w2v.npz"), **out)
    print("w2v.npz", {k: v.shape for k, v in out.items()})


def _import_faceformer():
    _paths()
    import gdl.models.DecaFLAME  # noqa: F401  (real module, imported before the stubs go in)
    stubs = ["easydict", "omegaconf", "scripts", "scripts.meshio", "visualize", "visualize.flame_visualization",
             "gdl.layers", "gdl.layers.losses", "gdl.layers.losses.DecaLosses", "gdl.utils.DecaUtils",
             "gdl.models.DECA", "third_party", "third_party.pirender", "third_party.pirender.generators",
             "third_party.pirender.generators.face_model", "third_party.pirender.config",
             "third_party.pirender.loss", "third_party.pirender.loss.perceptual",
             "third_party.pirender.util", "third_party.pirender.util.meters"]
    for s in stubs:
        if s not in sys.modules:
            sys.modules[s] = MagicMock(name=s)
    import models.faceformer_disentangle as ffd
    return ffd


class _FanStub(nn.Module):
    """Same 4-tuple API as FanEncoder.forward (pd_fgc_inference encoder.py:116-126). The images carry the frame index in pixel
    [0,0,0,0] (row 0: never masked); embeddings come from oracle.synth.fan_embeddings and are SCALED by (1 + sum of the image rows
    below 100/224 of the height), so an emotion frame that reaches the encoder without the reference's mask_lip (:119-133,:791)
    changes the result."""

    def __init__(self, emb):
        super().__init__()
        self.emb = emb

    def forward(self, img):
        i = int(round(float(img.reshape(img.shape[0], -1)[0, 0])))
        lower = 1.0 + float(img[:, :, int(100. / 224. * img.shape[2]):, :].sum())
        return self.emb["head"][i:i + 1], self.emb["eye"][i:i + 1], self.emb["emo"][i:i + 1] * lower, None


def build_reference_faceformer(ffd, fd, sd_ff, sd_w2v, template, fan_emb, period=30):
    args = types.SimpleNamespace(dataset="vocaset", feature_dim=fd, vertice_dim=15069, period=period,
                                 train_subjects="a b c d e f g h", device="cpu", is_concat_mode=0, load_mld=0)
    m = ffd.Faceformer.__new__(ffd.Faceformer)
    nn.Module.__init__(m)
    m.args, m.dataset, m.device = args, "vocaset", "cpu"
    m.audio_encoder = _ref_wav2vec2(sd_w2v)
    m.audio_feature_map = nn.Linear(768, fd)
    m.vertice_map = nn.Linear(15069, fd)
    m.vertice_map_r = nn.Linear(fd, 15069)
    m.obj_vector = nn.Linear(8, fd, bias=False)
    m.PPE = ffd.PeriodicPositionalEncoding(fd, period=period)
    m.biased_mask = ffd.init_biased_mask(n_head=4, max_seq_len=600, period=period)
    layer = nn.TransformerDecoderLayer(d_model=fd, nhead=4, dim_feedforward=2 * fd, batch_first=True)
    m.transformer_decoder = nn.TransformerDecoder(layer, num_layers=1)
    m.v_merge2hidden = nn.Linear(36 + fd, fd)
    m.learnable_eye_embed = nn.Parameter(torch.zeros(1, 1, 6))
    m.template = template
    m.fan_net = _FanStub(fan_emb)
    own = {k: v for k, v in m.state_dict().items() if not k.startswith("audio_encoder.") and not k.startswith("PPE.")}
    assert set(own) == set(sd_ff), (set(own) ^ set(sd_ff))
    m.load_state_dict(sd_ff, strict=False)
    return m.eval()


def golden_frames(T):
    """[T,3,4,4] frames: the frame index in pixel [0,0,0] and a non-zero mouth region (rows >= int(100/224*4) = 1) that the
    reference's mask_lip must remove before the encoder sees the emotion frames."""
    frames = torch.zeros(T, 3, 4, 4)
    frames[:, 0, 0, 0] = torch.arange(T).float()
    frames[:, :, 1:, :] = 0.37
    return frames


def golden_faceformer():
    from . import synth
    ffd = _import_faceformer()
    out = {}
    # closed-form pieces
    out["biased_mask_p30"] = ffd.init_biased_mask(4, 600, 30)[:, :64, :64].numpy()
    out["biased_mask_p25_full_sum"] = np.array(
        [torch.nan_to_num(ffd.init_biased_mask(4, 600, 25), neginf=0.0).double().sum().item()])
    out["enc_dec_mask_voca"] = ffd.enc_dec_mask("cpu", "vocaset", 5, 7).numpy()
    out["ppe_fd64_p30"] = ffd.PeriodicPositionalEncoding(64, period=30).pe[0, :70].numpy()

    sd_w2v = synth.wav2vec2_state(0)
    template = synth.flame_buffers()["v_template"].reshape(1, 1, 15069)
    for fd in (64, 128):
        sd_ff = synth.faceformer_state(fd=fd, seed=10 + fd)
        a = synth.audio(1, 16000, seed=1234)                       # 1 s -> T=24
        T = 24
        emb = synth.fan_embeddings(T, seed=20)
        m = build_reference_faceformer(ffd, fd, sd_ff, sd_w2v, template, emb)
        frames = golden_frames(T)
        v = m.predict(a, frames, frames, frames)                   # [1,24,15069]
        out[f"predict_fd{fd}_sub"] = v[0, :, ::COL_STRIDE].numpy()
        out[f"predict_fd{fd}_chk"] = checksum(v)
        # teacher-forced branch of forward_ff on the same hidden states
        with torch.no_grad():
            ha = m.audio_feature_map(m.audio_encoder(a, "vocaset").last_hidden_state)
            hs = torch.cat([m.learnable_eye_embed.expand(1, T, -1), emb["emo"][None], ha], -1)
            obj = m.obj_vector(torch.eye(8)[:1])
            gt = template + 1e-3 * torch.from_numpy(
                np.random.default_rng(77).normal(size=(1, T, 15069)).astype(np.float32))
            vt = m.forward_ff(gt, hs, obj, T, teacher_forcing=True)
        out[f"tf_fd{fd}_sub"] = vt[0, :, ::COL_STRIDE].numpy()
        out[f"tf_fd{fd}_chk"] = checksum(vt)
    # config C1: 4 s clip, fd=64, full predict (T=99)
    sd_ff = synth.faceformer_state(fd=64, seed=74)
    a = synth.audio(1, 64000, seed=1234)
    T = 99
    emb = synth.fan_embeddings(T, seed=20)
    m = build_reference_faceformer(ffd, 64, sd_ff, sd_w2v, template, emb)
    frames = golden_frames(T)
    v = m.predict(a, frames, frames, frames)
    out["predict_c1_sub"] = v[0, :, ::COL_STRIDE].numpy()
    out["predict_c1_chk"] = checksum(v)
    # loop_utils.loopback_frames index pattern
    from loop_utils import calc_loop_idx
    out["loop_idx_5_17"] = np.array([calc_loop_idx(i, 5) for i in range(17)])
    np.savez(os.path.join(GOLD, "faceformer.npz"), **out)
    print("faceformer.npz", {k: v.shape for k, v in out.items()})


class _TorchProxy:
    """`torch` as seen by the reference's class bodies, with randn / randn_like replaced by a queue of injected draws
    (the reference draws its sampling noise from a CUDA torch.Generator, train_diffusion_prior.py:803-804)."""

    def __init__(self, queue):
        self._q = queue

    def __getattr__(self, name):
        return getattr(torch, name)

    def randn(self, *a, **k):   # constructor-time parameter inits run with an empty queue -> real torch.randn
        return self._q.pop(0) if self._q else torch.randn(*a, **{kk: v for kk, v in k.items() if kk != "generator"})

    def randn_like(self, x):
        return self._q.pop(0) if self._q else torch.randn_like(x)


def _exec_reference_prior_classes(queue):
    """Execute the reference's OWN class sources from models/diffusion_prior.py (BrainNetwork :58-117, FlaggedCausalTransformer
    :119-166, VersatileDiffusionPriorNetwork :169-313, InstructDiffusionPrior :315-456) in a namespace where the un-vendored
    dalle2_pytorch / rotary_embedding_torch names resolve to oracle/dalle2_standin.py. The file itself cannot be imported
    (clip, dalle2_pytorch, torchvision transforms of PIL...)."""
    import ast
    from functools import partial

    from tqdm.auto import tqdm

    from . import dalle2_standin as d2
    path = os.path.join(REF, "models", "diffusion_prior.py")
    with open(path) as fh:
        src = fh.read()
    tree = ast.parse(src)
    want = ("BrainNetwork", "FlaggedCausalTransformer", "VersatileDiffusionPriorNetwork", "InstructDiffusionPrior")
    ns = {k: getattr(d2, k) for k in ("DiffusionPrior", "l2norm", "default", "exists", "RotaryEmbedding", "CausalTransformer",
                                      "SinusoidalPosEmb", "MLP", "Rearrange", "repeat", "rearrange", "prob_mask_like", "LayerNorm",
                                      "RelPosBias", "Attention", "FeedForward")}
    ns.update(torch=_TorchProxy(queue), nn=nn, partial=partial, tqdm=tqdm, random=__import__("random"))
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in want:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns


def golden_prior():
    from . import prior_oracle as po
    from . import synth
    sd = synth.prior_state()
    queue = []
    ns = _exec_reference_prior_classes(queue)
    out = {}
    with torch.no_grad():
        # BrainNetwork: fully the reference's own code (plain torch)
        brain = ns["BrainNetwork"](in_dim=768, out_dim=128, clip_size=128, use_projector=True).eval()
        brain.load_state_dict({k[len("voxel2clip."):]: v for k, v in sd.items() if k.startswith("voxel2clip.")}, strict=True)
        inp = synth.prior_inputs(4, 100)
        x, proj = brain(inp["voxel"])
        out["brain_x"], out["brain_proj"] = x.numpy(), proj.numpy()
        # prior network + sampler as constructed at train_diffusion_prior.py:972-991
        net = ns["VersatileDiffusionPriorNetwork"](dim=128, depth=6, dim_head=64, heads=8, causal=False, num_tokens=1,
                                                   learned_query_mode="pos_emb")
        prior = ns["InstructDiffusionPrior"](net=net, image_embed_dim=128, condition_on_text_encodings=False, timesteps=100,
                                             cond_drop_prob=0.2, image_embed_scale=None, voxel2clip=brain).eval()
        own = {k: v for k, v in prior.state_dict().items() if not k.startswith("noise_scheduler.")}
        assert set(own) == set(sd), sorted(set(own) ^ set(sd))
        prior.load_state_dict(sd, strict=False)
        out["state_keys"] = np.array(sorted(own))
        text = x.view(4, -1, 128)
        t = torch.full((4,), 37, dtype=torch.long)
        out["net_t37"] = net(inp["image_embed"], t, text_embed=text).numpy()
        # DDPM-100: the reference's p_sample_loop_ddpm / p_sample with injected draws (noises[i] = draw of step i)
        queue.extend([inp["noises"][i] for i in reversed(range(100))])
        gen = object()  # any non-None value selects the generator branch of p_sample (:333-337); the proxy ignores it
        y = prior.p_sample_loop(text.shape, text_cond=dict(text_embed=text), cond_scale=1.0, timesteps=100, generator=gen,
                                image_embed=inp["image_embed"])
        assert not queue
        out["ddpm100"] = y.numpy()
        # DDIM-64: base-class sampler (stand-in) around the reference's network
        pairs = po.ddim_time_pairs(100, 64)
        y = prior.p_sample_loop(text.shape, text_cond=dict(text_embed=text), cond_scale=1.0, timesteps=64,
                                image_embed=inp["image_embed"], noises=[inp["noises"][k] for k in range(len(pairs))])
        out["ddim64"] = y.numpy()
        for name in ("betas", "alphas_cumprod_prev", "posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped"):
            out["sched_" + name] = getattr(prior.noise_scheduler, name).numpy()
    np.savez(os.path.join(GOLD, "prior.npz"), **out)
    print("prior.npz", {k: v.shape for k, v in out.items()})



def golden_prior_cfg():
    """Classifier-free guidance (cond_scale = 2.5) through the reference's own forward_with_cond_scale / p_sample /
    p_sample_loop_ddpm and the stand-in's DDIM loop: tests/golden/prior_cfg.npz."""
    from . import prior_oracle as po
    from . import synth
    sd = synth.prior_state()
    queue = []
    ns = _exec_reference_prior_classes(queue)
    out = {}
    with torch.no_grad():
        brain = ns["BrainNetwork"](in_dim=768, out_dim=128, clip_size=128, use_projector=True).eval()
        net = ns["VersatileDiffusionPriorNetwork"](dim=128, depth=6, dim_head=64, heads=8, causal=False, num_tokens=1,
                                                   learned_query_mode="pos_emb")
        prior = ns["InstructDiffusionPrior"](net=net, image_embed_dim=128, condition_on_text_encodings=False, timesteps=100,
                                             cond_drop_prob=0.2, image_embed_scale=None, voxel2clip=brain).eval()
        prior.load_state_dict(sd, strict=False)
        inp = synth.prior_inputs(4, 100)
        text = brain(inp["voxel"])[0].view(4, -1, 128)
        cs = 2.5
        t = torch.full((4,), 37, dtype=torch.long)
        out["net_t37_cfg"] = net.forward_with_cond_scale(inp["image_embed"], t, cond_scale=cs, text_embed=text).numpy()
        queue.extend([inp["noises"][i] for i in reversed(range(100))])
        y = prior.p_sample_loop(text.shape, text_cond=dict(text_embed=text), cond_scale=cs, timesteps=100, generator=object(),
                                image_embed=inp["image_embed"])
        assert not queue
        out["ddpm100_cfg"] = y.numpy()
        pairs = po.ddim_time_pairs(100, 64)
        y = prior.p_sample_loop(text.shape, text_cond=dict(text_embed=text), cond_scale=cs, timesteps=64,
                                image_embed=inp["image_embed"], noises=[inp["noises"][k] for k in range(len(pairs))])
        out["ddim64_cfg"] = y.numpy()
        out["cond_scale"] = np.array(cs)
    np.savez(os.path.join(GOLD, "prior_cfg.npz"), **out)
    print("prior_cfg.npz", {k: v.shape for k, v in out.items()})


def prior_train_inputs(B=6, seed=77):
    """Seeded inputs of one prior-training iteration (every stochastic draw of the reference as an explicit tensor)."""
    g = torch.Generator().manual_seed(seed)
    voxel = torch.randn(B, 768, generator=g)                                   # mean CLIP text embedding (:438-439)
    clip_target = torch.randn(B, 1, 128, generator=g) * 0.5                    # EMOTE style latent (:195)
    times = torch.tensor([3, 97, 40, 0, 62, 15][:B] + [int(x) for x in torch.randint(0, 100, (max(0, B - 6),), generator=g)], dtype=torch.long)
    noise = torch.randn(B, 1, 128, generator=g)
    keep_brain = torch.tensor(([True, False, True, True, False, True] * ((B + 5) // 6))[:B])
    keep_image = torch.tensor(([True, True, False, True, False, True] * ((B + 5) // 6))[:B])
    masks = [(torch.rand(B, 4096, generator=g) >= p).float() / (1 - p) for p in (0.5, 0.15, 0.15, 0.15, 0.15)]
    return dict(voxel=voxel, clip_target=clip_target, times=times, noise=noise, keep_brain=keep_brain, keep_image=keep_image, masks=masks)


def sample_tensor(t, n=2048):
    """Compact fingerprint of a (possibly huge) tensor: n evenly strided elements, the sum and the l2 norm."""
    f = t.detach().reshape(-1).double()
    step = max(1, f.numel() // n)
    return f[::step][:n].float().numpy(), np.array([float(f.sum()), float(f.norm())])


class _MaskDropout(nn.Module):
    def __init__(self, mask):
        super().__init__()
        self.mask = mask

    def forward(self, x):
        return x * self.mask


def golden_prior_train():
    """One iteration of train_diffusion_prior.py:434-486 executed on the reference's OWN classes (BrainNetwork,
    VersatileDiffusionPriorNetwork, InstructDiffusionPrior.forward / p_losses) over the dalle2 stand-in, its own soft_clip_loss,
    loss.backward() and torch.optim.AdamW with the groups of :996-1004. Draws are injected: timesteps (sample_random_times),
    noise (p_losses' noise argument), the two keep masks (prob_mask_like), BrainNetwork's Dropout masks."""
    from . import synth
    sd = synth.prior_state()
    ns = _exec_reference_prior_classes([])
    fns = _exec_reference_functions(os.path.join(REF, "train_diffusion_prior.py"), ("soft_clip_loss",), dict(torch=torch, nn=nn))
    inp = prior_train_inputs()
    for variant in ("eval", "dropout"):
        brain = ns["BrainNetwork"](in_dim=768, out_dim=128, clip_size=128, use_projector=True)
        net = ns["VersatileDiffusionPriorNetwork"](dim=128, depth=6, dim_head=64, heads=8, causal=False, num_tokens=1,
                                                   learned_query_mode="pos_emb")
        prior = ns["InstructDiffusionPrior"](net=net, image_embed_dim=128, condition_on_text_encodings=False, timesteps=100,
                                             cond_drop_prob=0.2, image_embed_scale=None, voxel2clip=brain)
        prior.load_state_dict(sd, strict=False)
        prior.train()
        if variant == "dropout":
            brain.lin0[3] = _MaskDropout(inp["masks"][0])
            for i in range(4):
                brain.mlp[i][3] = _MaskDropout(inp["masks"][i + 1])
        else:
            brain.eval()
        prior.noise_scheduler.sample_random_times = lambda b: inp["times"]
        keep = [inp["keep_brain"], inp["keep_image"]]
        ns["prob_mask_like"] = lambda shape, prob, device: keep.pop(0)
        no_decay = ["bias", "LayerNorm.bias", "LayerNorm.weight"]                                   # :997-1003, verbatim grouping
        groups = [
            {"params": [p for n, p in prior.net.named_parameters() if not any(nd in n for nd in no_decay)], "weight_decay": 1e-2},
            {"params": [p for n, p in prior.net.named_parameters() if any(nd in n for nd in no_decay)], "weight_decay": 0.0},
            {"params": [p for n, p in prior.voxel2clip.named_parameters() if not any(nd in n for nd in no_decay)], "weight_decay": 1e-2},
            {"params": [p for n, p in prior.voxel2clip.named_parameters() if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
        opt = torch.optim.AdamW(groups, lr=3e-4)
        opt.zero_grad()
        voxel, clip_target = inp["voxel"].clone().requires_grad_(True), inp["clip_target"].clone().requires_grad_(True)
        clip_voxels, clip_voxels_proj = prior.voxel2clip(voxel)                                     # :441
        clip_voxels = clip_voxels.view(len(voxel), -1, 128)                                         # :445
        loss_prior, aligned = prior(text_embed=clip_voxels, image_embed=clip_target, noise=inp["noise"])   # :449
        assert not keep
        clip_voxels_norm = nn.functional.normalize(clip_voxels_proj.flatten(1), dim=-1)             # :455-456
        clip_target_norm = nn.functional.normalize(clip_target.flatten(1), dim=-1)
        temp = 0.0045
        loss_nce = fns["soft_clip_loss"](clip_voxels_norm, clip_target_norm, temp=temp)             # :465-468
        loss = loss_nce + 30 * loss_prior                                                           # :474
        loss.backward()
        out = dict(loss_nce=np.array(float(loss_nce)), loss_prior=np.array(float(loss_prior)), pred=aligned.detach().numpy(),
                   temp=np.array(temp), names=np.array([n for n, p in prior.named_parameters() if p.requires_grad]))
        for n, p in prior.named_parameters():
            if p.requires_grad:
                assert p.grad is not None, n
                out["g:" + n], out["gs:" + n] = sample_tensor(p.grad)
        opt.step()
        for n, p in prior.named_parameters():
            if p.requires_grad:
                out["p:" + n], out["ps:" + n] = sample_tensor(p)
        np.savez_compressed(os.path.join(GOLD, f"prior_train_{variant}.npz"), **out)
        print(f"prior_train_{variant}.npz", float(loss_nce), float(loss_prior), len(out))


def build_reference_talking_head(sd, sd_w2v):
    """The reference's EMOTE inference graph assembled from its OWN classes (stub-imported, oracle/_inferno_import.py):
    TalkingHeadBase.forward over Wav2Vec2Encoder/Wav2Vec2ModelResampled, LinearSequenceEncoder, BertPriorDecoder (+ real
    LinearEmotionCondition, StackLinearSquash), MotionPrior.decoding_step over the real L2lDecoder and FlamePreprocessor/FLAME.
    Constructors that need checkpoints / cfg.yaml / the network are bypassed with __new__ and the attributes they would set."""
    from collections import OrderedDict

    from transformers import Wav2Vec2Config, Wav2Vec2FeatureExtractor

    from . import _inferno_import as ii
    from . import synth
    Munch = ii.Munch
    (ffdec, l2l, mp, pre, seqenc, audenc, thb) = ii.load(
        "inferno.models.talkinghead.FaceFormerDecoder", "inferno.models.temporal.motion_prior.L2lMotionPrior",
        "inferno.models.temporal.motion_prior.MotionPrior", "inferno.models.temporal.Preprocessors",
        "inferno.models.temporal.SequenceEncoders", "inferno.models.temporal.AudioEncoders", "inferno.models.talkinghead.TalkingHeadBase")
    from inferno.models.DecaFLAME import FLAME
    E = synth.EMOTE
    fcfg = synth.write_flame_assets("/tmp/avi_flame_assets")
    fcfg.n_shape, fcfg.n_exp = E.n_shape, E.n_exp
    flame = FLAME(fcfg)
    # FlamePreprocessor (Preprocessors.py:27-60)
    prep = pre.FlamePreprocessor.__new__(pre.FlamePreprocessor)      # a plain object, not an nn.Module (Bases.py:168)
    prep.cfg = Munch(flame=Munch(n_exp=E.n_exp, n_shape=E.n_shape), use_texture=False, test_time=True)
    prep.flame, prep.flame_tex = flame, None
    # L2lDecoder: real constructor (motion_prior_conf l2l_decoder.yaml / l2l_sizes.yaml)
    dcfg = Munch(feature_dim=E.bottleneck, nhead=8, intermediate_size=E.l2l_ff, activation="gelu", dropout=0.0, num_layers=1,
                 positional_encoding=Munch(type="none"), temporal_bias=Munch(type="alibi_future", max_len=600))
    l2l_dec = l2l.L2lDecoder(dcfg, Munch(quant_factor=E.quant_factor, sequence_length=32), E.n_out)
    mprior = mp.MotionPrior.__new__(mp.MotionPrior)
    nn.Module.__init__(mprior)
    mprior.cfg = Munch(model=Munch(sequence_components=OrderedDict([("exp", 50), ("jaw", "rot")]), rotation_representation="aa",
                                   sizes=Munch(quant_factor=E.quant_factor)))
    mprior.motion_encoder, mprior.motion_quantizer, mprior.motion_decoder = None, None, l2l_dec
    mprior.preprocessor = mprior.postprocessor = prep
    # BertPriorDecoder (FaceFormerDecoder.py:987-1075, bertprior_wild.yaml)
    scfg = Munch(use_video_expression=False, use_video_feature=False, gt_expression_label=True, gt_expression_intensity=True,
                 n_intensities=E.n_intensities, gt_expression_identity=True, n_identities=E.n_identities, use_expression=False,
                 n_expression=E.n_expression, use_valence=False, use_arousal=False, use_emotion_feature=False, use_shape=False, use_bias=True)
    dec = ffdec.BertPriorDecoder.__new__(ffdec.BertPriorDecoder)
    nn.Module.__init__(dec)
    dec.cfg = Munch(feature_dim=E.feature_dim, nhead=8, num_layers=1, post_bug_fix=True, motion_prior=Munch(trainable=False))
    dec.style_type, dec.style_op, dec.PE = "emotion_linear", "add", None
    dec.obj_vector = ffdec.LinearEmotionCondition(scfg, E.feature_dim)
    layer = nn.TransformerEncoderLayer(d_model=E.feature_dim, nhead=8, dim_feedforward=E.feature_dim, activation="gelu", dropout=0.25,
                                       batch_first=True)
    dec.bert_decoder = nn.TransformerEncoder(layer, num_layers=1, enable_nested_tensor=False)
    dec.post_bug_fix, dec.temporal_bias_type, dec.biased_mask = True, "none", None
    dec.motion_prior, dec.latent_frame_size, dec.flame = mprior, E.latent_frame, flame
    dec.squasher = None
    dec.decoder = nn.Linear(E.feature_dim, E.bottleneck)
    dec.squasher_2 = ffdec.StackLinearSquash(E.bottleneck, E.latent_frame, E.bottleneck)
    # audio model (AudioEncoders.py:130-166) - real Wav2Vec2ModelResampled, real HF feature extractor as the processor
    aud = audenc.Wav2Vec2Encoder.__new__(audenc.Wav2Vec2Encoder)
    nn.Module.__init__(aud)
    aud.input_processor = Wav2Vec2FeatureExtractor(feature_size=1, sampling_rate=16000, padding_value=0.0, do_normalize=True,
                                                   return_attention_mask=False)
    aud.model = audenc.Wav2Vec2ModelResampled(Wav2Vec2Config(attn_implementation="eager"))
    aud.model.load_state_dict(sd_w2v, strict=False)
    aud.resampling, aud.dropout, aud.trainable = True, None, False
    enc = seqenc.LinearSequenceEncoder(Munch(feature_dim=E.feature_dim, input_feature_dim=768))

    class _Base(thb.TalkingHeadBase):
        device = torch.device("cpu")

    model = _Base.__new__(_Base)
    nn.Module.__init__(model)
    model.cfg = Munch(data=Munch(), model=Munch())
    model.audio_model, model.sequence_encoder, model.sequence_decoder = aud, enc, dec
    model.preprocessor, model.renderer, model.neural_losses, model.shape_model = prep, None, {}, None
    own = {k: v for k, v in model.state_dict().items() if k.startswith("sequence_")
           and ".flame." not in k and "preprocessor" not in k and "postprocessor" not in k}
    assert set(own) == set(sd), sorted(set(own) ^ set(sd))
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    return model.eval(), sorted(own)


def golden_emote():
    from . import _inferno_import as ii
    from . import synth
    sd, sd_w2v = synth.emote_state(), synth.wav2vec2_state(0)
    model, keys = build_reference_talking_head(sd, sd_w2v)
    out = {"state_keys": np.array(keys)}
    # The reference pins torch 1.9.0 (requirements.txt:5), which has no fused "fast path" for nn.TransformerEncoderLayer. The
    # fast path of the torch installed here (2.11, eval + no_grad) mishandles the 3-D additive float mask the L2L decoder passes
    # ([B*heads, T, T], L2lMotionPrior.py:477-481): it disagrees with the layer's own reference (slow) path by O(1). The goldens
    # are therefore minted on the slow path = the arithmetic torch 1.9 performs.
    torch.backends.mha.set_fastpath_enabled(False)
    with torch.no_grad():
        # B = 1 (the only way the reference calls it, evaluation_functions.py:381-383), T = 27 (not a multiple of 8 -> padding path)
        for tag, T in (("t27", 27), ("t48", 48)):
            s = synth.emote_sample(1, T, seed=50)
            r = model(dict(s))
            for k in ("predicted_exp", "predicted_jaw", "prior_input_sequence"):
                out[f"{tag}_{k}"] = r[k].numpy()
            out[f"{tag}_gt_vertices_sub"] = r["gt_vertices"].numpy()[:, :, ::COL_STRIDE]
            out[f"{tag}_template"] = r["template"].numpy()[:, ::COL_STRIDE]
            out[f"{tag}_predicted_vertices_sub"] = r["predicted_vertices"].numpy()[:, :, ::COL_STRIDE]
            out[f"{tag}_predicted_vertices_chk"] = checksum(r["predicted_vertices"])
        # external style embedding [B,1,128] (what voxel2style_emb feeds, train_diffusion_prior.py:195,218)
        s = synth.emote_sample(1, 27, seed=50)
        style = torch.from_numpy(np.random.default_rng(60).normal(0, 0.5, size=(1, 1, 128)).astype(np.float32))
        r = model(dict(s), style_emb=style, is_external_style_emb=True)
        out["ext_predicted_exp"], out["ext_predicted_jaw"] = r["predicted_exp"].numpy(), r["predicted_jaw"].numpy()
        out["ext_predicted_vertices_chk"] = checksum(r["predicted_vertices"])
        # only_style_emb branch (FaceFormerDecoder.py:599-601)
        out["style_only"] = model(dict(s), only_style_emb=True).numpy()
        mask = ii.load("inferno.models.temporal.TransformerMasking")[0].init_alibi_biased_mask_future(8, 600)
        out["alibi_future_8_40"] = mask[:, :40, :40].numpy()
    np.savez(os.path.join(GOLD, "emote.npz"), **out)
    print("emote.npz", {k: v.shape for k, v in out.items()})


def build_reference_faceformer_vert(ffv, fd, sd_ff, sd_w2v, flame, fan_emb, period=30):
    """models/faceformer_vert.Faceformer through __new__ (the constructor needs the network and private assets), carrying exactly
    the attributes forward_switch_frame reads (:360-519)."""
    args = types.SimpleNamespace(dataset="vocaset", feature_dim=fd, vertice_dim=15069, period=period,
                                 train_subjects="a b c d e f g h", device="cpu", is_only_emo=True)
    m = ffv.Faceformer.__new__(ffv.Faceformer)
    nn.Module.__init__(m)
    m.args, m.dataset, m.device, m.vertice_scale = args, "vocaset", "cpu", 1.0
    m.audio_encoder = _ref_wav2vec2(sd_w2v)
    m.audio_encoder.feature_extractor._freeze_parameters()                       # :154
    m.audio_feature_map = nn.Linear(768, fd)
    m.vertice_map = nn.Linear(15069, fd)
    m.vertice_map_r = nn.Linear(fd, 15069)
    m.obj_vector = nn.Linear(8, fd, bias=False)
    m.PPE = ffv.PeriodicPositionalEncoding(fd, period=period)
    m.biased_mask = ffv.init_biased_mask(n_head=4, max_seq_len=600, period=period)
    layer = nn.TransformerDecoderLayer(d_model=fd, nhead=4, dim_feedforward=2 * fd, batch_first=True)
    m.transformer_decoder = nn.TransformerDecoder(layer, num_layers=1)
    m.flame = flame
    m.template = flame.v_template.reshape(1, 1, 15069) * m.vertice_scale         # :184
    m.fan_net = _FanStub(fan_emb)
    m.meters, m.losses_dict = {}, {}
    own = {k: v for k, v in m.state_dict().items()
           if not k.startswith(("audio_encoder.", "PPE.", "flame."))}
    assert set(own) == set(sd_ff), (set(own) ^ set(sd_ff))
    m.load_state_dict(sd_ff, strict=False)
    return m.eval()          # dropout / SpecAugment / LayerDrop inactive: the step is then a deterministic function of its inputs


GRAD_STRIDE = 509  # big gradient tensors are stored flattened [::GRAD_STRIDE]


def train_inputs(B, T, seed=90):
    """coeff [B,T,53] normalised, pose [B,T,6], shape [B,T,100], coeff_mean/std [1,1,53] (config 5 inputs)."""
    rng = np.random.default_rng(seed)
    f = lambda *s, sc=1.0: torch.from_numpy((sc * rng.normal(size=s)).astype(np.float32))  # noqa: E731
    coeff = f(B, T, 53)
    pose = torch.cat([f(B, T, 3, sc=0.2), f(B, T, 3, sc=0.1)], -1)
    shape = f(B, T, 100)
    mean = f(1, 1, 53, sc=0.1)
    std = 0.5 + 0.1 * f(1, 1, 53).abs()
    std[..., 50:] = 0.1
    return coeff, pose, shape, mean, std


def golden_train():
    """The reference's OWN faceformer_vert forward_switch_frame (:360-542) -> loss.backward() -> torch.optim.Adam step."""
    import pdb
    from gdl.models.DecaFLAME import FLAME_mediapipe
    from . import synth
    _import_faceformer()
    import models.faceformer_vert as ffv
    pdb.set_trace = lambda *a, **k: None                     # left enabled upstream at :541
    cwd = os.getcwd()
    os.chdir("/tmp")                                           # the debug visualisation makes ./intermediate_res (:527-528)
    out = {}
    try:
        sd_w2v = synth.wav2vec2_state(0)
        cfg = synth.write_flame_assets("/tmp/avi_flame_assets")
        cfg.n_shape = 100
        flame = FLAME_mediapipe(cfg)
        for tag, fd, B, n_samples, T in (("a", 64, 2, 16000, 24), ("b", 128, 1, 16000, 20)):
            sd_ff = synth.faceformer_state(fd=fd, seed=200 + fd, variant="vert")
            m = build_reference_faceformer_vert(ffv, fd, sd_ff, sd_w2v, flame, synth.fan_embeddings(T, seed=20))
            coeff, pose, shape, mean, std = train_inputs(B, T, seed=90 + fd)
            m.coeff_mean, m.coeff_std = mean, std
            audio = synth.audio(B, n_samples, seed=4321)
            img = torch.zeros(B, T, 3, 4, 4)
            img[:, :, 0, 0, 0] = torch.arange(T).float()
            opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-4)
            opt.zero_grad(set_to_none=True)
            loss = m.forward_switch_frame(audio, coeff, pose.clone(), shape, img=img, criterion=nn.MSELoss(reduction="none"),
                                          teacher_forcing=True)
            loss.backward()
            out[f"{tag}_loss"] = np.array([loss.item()])
            names = [n for n, p in m.named_parameters() if p.requires_grad]
            for n, p in m.named_parameters():
                if not p.requires_grad:
                    continue
                g = p.grad if p.grad is not None else torch.zeros_like(p)
                out[f"{tag}_gchk/{n}"] = checksum(g)
                out[f"{tag}_g/{n}"] = (g.reshape(-1)[::GRAD_STRIDE] if g.numel() > 4096 else g.reshape(-1)).numpy().copy()
            before = {n: p.detach().clone() for n, p in m.named_parameters() if p.requires_grad}
            opt.step()
            for n, p in m.named_parameters():
                if p.requires_grad:
                    q = p.detach() - before[n]          # the Adam update itself (|dp| ~ lr), not the parameter
                    out[f"{tag}_dp/{n}"] = (q.reshape(-1)[::GRAD_STRIDE] if q.numel() > 4096 else q.reshape(-1)).numpy().copy()
            out[f"{tag}_names"] = np.array(names)
            print("train", tag, "loss", loss.item(), "trainable tensors", len(names))
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(GOLD, "train.npz"), **out)
    print("train.npz", len(out), "arrays")


def golden_train_reg():
    """The same step in TRAIN mode: the reference's own forward_switch_frame with dropout (HF Wav2Vec2: feat_proj / hidden / attention /
    activation; PeriodicPositionalEncoding; nn.TransformerDecoderLayer), SpecAugment (models/lib/wav2vec.py:120-131) and LayerDrop active.
    Every draw is injected: torch.nn.functional.dropout pops the pre-scaled masks of synth.train_regularisers in call order (shapes are
    asserted, which pins the site order), torch.rand([]) yields the LayerDrop decisions, _compute_mask_indices returns the SpecAugment
    mask. nn.MultiheadAttention's need_weights=False path hides its dropout inside the fused scaled_dot_product_attention, so that
    function is replaced by its definition (softmax -> dropout -> @ v), which is also what the pinned torch 1.9 executed."""
    import math as _math
    import pdb
    from gdl.models.DecaFLAME import FLAME_mediapipe
    from . import synth
    _import_faceformer()
    import models.faceformer_vert as ffv
    import models.lib.wav2vec as ref_w2v
    pdb.set_trace = lambda *a, **k: None
    cwd = os.getcwd()
    os.chdir("/tmp")
    out = {}
    Fn = torch.nn.functional
    saved = (Fn.dropout, Fn.scaled_dot_product_attention, torch.rand, ref_w2v._compute_mask_indices)
    try:
        sd_w2v = synth.wav2vec2_state(0)
        cfg = synth.write_flame_assets("/tmp/avi_flame_assets")
        cfg.n_shape = 100
        flame = FLAME_mediapipe(cfg)
        fd, B, n_samples, T = 64, 2, 16000, 24
        sd_ff = synth.faceformer_state(fd=fd, seed=200 + fd, variant="vert")
        m = build_reference_faceformer_vert(ffv, fd, sd_ff, sd_w2v, flame, synth.fan_embeddings(T, seed=20))
        m.train()
        reg = synth.train_regularisers(B, T, fd, seed=300)
        mk = reg["masks"]
        queue = [(n, mk[n]) for n in reg["order"] if not n.startswith(("ppe", "dec."))]
        for j in range(B):                                   # the decoder runs clip by clip (:437-454)
            queue += [("ppe", mk["ppe"].view(B, T, fd)[j:j + 1]), ("dec.sa", mk["dec.sa"][j:j + 1]), ("dec.d1", mk["dec.d1"].view(B, T, fd)[j:j + 1]),
                      ("dec.ca", mk["dec.ca"][j:j + 1]), ("dec.d2", mk["dec.d2"].view(B, T, fd)[j:j + 1]),
                      ("dec.act", mk["dec.act"].view(B, T, 2 * fd)[j:j + 1]), ("dec.d3", mk["dec.d3"].view(B, T, fd)[j:j + 1])]
        used = []

        def fake_dropout(input, p=0.5, training=True, inplace=False):
            if not training or p == 0.0:
                return input
            name, mask = queue.pop(0)
            assert abs(p - reg["p"]) < 1e-9 and mask.numel() == input.numel(), (name, p, tuple(mask.shape), tuple(input.shape))
            used.append(name)
            return input * mask.reshape(input.shape)

        def sdpa(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None, **kw):
            s_ = q @ k.transpose(-2, -1) / _math.sqrt(q.shape[-1])
            if attn_mask is not None:
                s_ = s_ + attn_mask
            return fake_dropout(torch.softmax(s_, dim=-1), dropout_p, True) @ v

        keep = list(reg["layer_keep"])

        def fake_rand(*a, **k):
            if len(a) == 1 and a[0] == [] and keep:
                return torch.tensor(0.5 if keep.pop(0) else 0.0)        # < layerdrop (0.1) skips the layer
            return saved[2](*a, **k)

        Fn.dropout, Fn.scaled_dot_product_attention, torch.rand = fake_dropout, sdpa, fake_rand
        ref_w2v._compute_mask_indices = lambda *a, **k: reg["spec_mask"].numpy()
        assert m.audio_encoder.config.apply_spec_augment and m.audio_encoder.config.mask_time_prob > 0
        assert abs(m.audio_encoder.config.layerdrop - 0.1) < 1e-9
        coeff, pose, shape, mean, std = train_inputs(B, T, seed=90 + fd)
        m.coeff_mean, m.coeff_std = mean, std
        audio = synth.audio(B, n_samples, seed=4321)
        img = torch.zeros(B, T, 3, 4, 4)
        img[:, :, 0, 0, 0] = torch.arange(T).float()
        opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-4)
        opt.zero_grad(set_to_none=True)
        loss = m.forward_switch_frame(audio, coeff, pose.clone(), shape, img=img, criterion=nn.MSELoss(reduction="none"),
                                      teacher_forcing=True)
        assert not queue and not keep, (len(queue), keep)
        loss.backward()
        out["loss"] = np.array([loss.item()])
        names = [n for n, p in m.named_parameters() if p.requires_grad]
        for n, p in m.named_parameters():
            if not p.requires_grad:
                continue
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            out[f"gchk/{n}"] = checksum(g)
            out[f"g/{n}"] = (g.reshape(-1)[::GRAD_STRIDE] if g.numel() > 4096 else g.reshape(-1)).numpy().copy()
        before = {n: p.detach().clone() for n, p in m.named_parameters() if p.requires_grad}
        opt.step()
        for n, p in m.named_parameters():
            if p.requires_grad:
                q = p.detach() - before[n]
                out[f"dp/{n}"] = (q.reshape(-1)[::GRAD_STRIDE] if q.numel() > 4096 else q.reshape(-1)).numpy().copy()
        out["names"] = np.array(names)
        out["sites_in_call_order"] = np.array(used)
        print("train_reg loss", loss.item(), "dropout calls", len(used), "trainable tensors", len(names))
    finally:
        Fn.dropout, Fn.scaled_dot_product_attention, torch.rand, ref_w2v._compute_mask_indices = saved
        os.chdir(cwd)
    np.savez_compressed(os.path.join(GOLD, "train_reg.npz"), **out)
    print("train_reg.npz", len(out), "arrays")


def _exec_reference_functions(path, names, extra_globals):
    """Compile only the named top-level functions of a reference file (its imports need librosa / pytorch_lightning / ...)."""
    import ast
    src = open(path).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(keep) == len(names), [n.name for n in keep]
    g = {"np": np, "__name__": "reference_subset"}
    g.update(extra_globals)
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), g)
    return g


def frontend_wav(n, seed):
    return (np.random.default_rng(seed).normal(size=n) * 3000).clip(-32767, 32767).astype(np.int16)


def golden_frontend():
    """The reference's OWN process_audio / create_base_sample (evaluation_functions.py:141-160,690-714), compiled from source."""
    path = os.path.join(REF, "third_party", "inferno", "inferno_apps", "TalkingHead", "evaluation", "evaluation_functions.py")
    cur = {}
    g = _exec_reference_functions(path, ["process_audio", "create_base_sample"], {
        "read_audio": lambda p: (cur["wav"], 16000),
        "create_condition": lambda th, sample: sample,
    })
    th = types.SimpleNamespace(cfg=types.SimpleNamespace(data=types.SimpleNamespace(reconstruction_type=["rec"])))
    out = {}
    for i, (n, kw) in enumerate([(16000 * 4 + 123, {}), (16000 * 2, dict(smallest_unit=8)), (9999, dict(silent_frames_start=3, silent_frames_end=2)),
                                 (640 * 7, dict(smallest_unit=4, silence_all=True)), (100, {})]):
        cur["wav"] = frontend_wav(n, 700 + i)
        pa = g["process_audio"](cur["wav"], 16000, 25)
        out[f"pa_{i}"] = pa["raw_audio"]
        s = g["create_base_sample"](th, "unused.wav", **kw)
        out[f"cbs_{i}_raw"] = s["raw_audio"]
        for k in ("gt_exp", "gt_shape", "gt_jaw", "gt_tex"):
            out[f"cbs_{i}_{k}_shape"] = np.array(s["reconstruction"]["rec"][k].shape)
    np.savez_compressed(os.path.join(GOLD, "frontend.npz"), **out)
    print("frontend.npz", {k: v.shape for k, v in out.items() if "raw" in k or k.startswith("pa_")})


def golden_clip_text():
    """transformers.CLIPTextModel - the class models/diffusion_prior.py:19,37 instantiates - on the seeded synthetic state dict."""
    from transformers import CLIPTextConfig, CLIPTextModel
    from . import synth
    out = {}
    for tag, layers, B in (("l12", 12, 3), ("l2", 2, 2)):
        cfg = CLIPTextConfig(vocab_size=synth.CLIP_TEXT.vocab, hidden_size=768, intermediate_size=3072, num_hidden_layers=layers,
                             num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu", projection_dim=768)
        m = CLIPTextModel(cfg).eval()
        missing, unexpected = m.load_state_dict(synth.clip_text_state(60, layers), strict=False)
        assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
        ids = synth.clip_tokens(B, seed=61)
        with torch.no_grad():
            o = m(input_ids=ids)
        out[f"{tag}_last_sub"] = o.last_hidden_state[:, ::4, ::3].numpy()
        out[f"{tag}_voxel"] = o.last_hidden_state.mean(dim=1).numpy()
        out[f"{tag}_chk"] = checksum(o.last_hidden_state)
    np.savez_compressed(os.path.join(GOLD, "clip_text.npz"), **out)
    print("clip_text.npz", {k: v.shape for k, v in out.items()})


def golden_subject_labels():
    """The reference's own TalkingHeadWrapper.get_subject_labels (TalkingHeadWrapper.py:168-236), compiled from source."""
    import ast
    import json
    import random
    path = os.path.join(REF, "third_party", "inferno", "inferno_apps", "TalkingHead", "evaluation", "TalkingHeadWrapper.py")
    cls = [n for n in ast.parse(open(path).read()).body if isinstance(n, ast.ClassDef) and n.name == "TalkingHeadWrapper"][0]
    fn = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "get_subject_labels"][0]
    g = {"rand": random}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), g)
    res = {}
    for split in ("random_by_identityV2_sorted_70_15_15", "random_by_identityV2_random_70_15_15", "random_by_identityV2_sorted_85_15_0"):
        me = types.SimpleNamespace(cfg=types.SimpleNamespace(data=types.SimpleNamespace(data_class="MEADPseudo3DDM", split=split)))
        for which in ("training", "validation", "testing"):
            res[f"{split}/{which}"] = g["get_subject_labels"](me, which)
    json.dump(res, open(os.path.join(GOLD, "subject_labels.json"), "w"), indent=0)


def _import_reference_fan():
    """third_party/pd_fgc_inference ... encoder.FanEncoder; its constructor only reads pose_dim / eye_dim from a YAML through omegaconf
    (absent here): a two-field stand-in for omegaconf.OmegaConf.load is registered before the import."""
    _paths()
    om = types.ModuleType("omegaconf")
    net_motion = types.SimpleNamespace(pose_dim=6, eye_dim=6, motion_dim=18)
    om.OmegaConf = types.SimpleNamespace(load=lambda path: types.SimpleNamespace(model=types.SimpleNamespace(net_motion=net_motion)))
    saved = sys.modules.get("omegaconf")
    sys.modules["omegaconf"] = om
    # _import_faceformer() registers MagicMock stand-ins for `third_party` / `third_party.pirender...`: lift them while the REAL
    # third_party.pd_fgc_inference package is imported, then put them back
    mocks = {k: v for k, v in sys.modules.items() if (k == "third_party" or k.startswith("third_party.")) and isinstance(v, MagicMock)}
    for k in mocks:
        del sys.modules[k]
    try:
        import third_party.pd_fgc_inference.lib.models.networks.encoder as enc
    finally:
        if saved is not None:
            sys.modules["omegaconf"] = saved
        for k, v in mocks.items():
            sys.modules.setdefault(k, v)
    return enc


def golden_fan():
    from . import synth
    enc = _import_reference_fan()
    m = enc.FanEncoder().eval()
    missing, unexpected = m.load_state_dict(synth.fan_state(80), strict=True)
    x = synth.fan_images(3, seed=81)
    with torch.no_grad():
        head, eye, emo, mouth = m(x)
        feat = m.forward_feature(x)
    out = dict(head=head.numpy(), eye=eye.numpy(), emo=emo.numpy(), mouth=mouth.numpy(), feat=feat.numpy())
    np.savez_compressed(os.path.join(GOLD, "fan.npz"), **out)
    print("fan.npz", {k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})


def main():
    _paths()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    golden_flame()
    golden_lbs_rotmat()
    golden_masks()
    golden_host_helpers()
    golden_wav2vec2()
    golden_faceformer()
    golden_prior()
    golden_prior_cfg()
    golden_prior_train()
    golden_emote()
    golden_train()
    golden_train_reg()
    golden_frontend()
    golden_clip_text()
    golden_subject_labels()
    golden_fan()


if __name__ == "__main__":
    main()
