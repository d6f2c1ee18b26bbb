"""Drop-in for the EMOTE talking-head inference path that ``experiments/diffusion_test.sh`` runs (Path B, third_party/inferno):

  TalkingHeadWrapper.forward(sample, style_emb=None, only_style_emb=False, is_external_style_emb=False)
      inferno_apps/TalkingHead/evaluation/TalkingHeadWrapper.py:123-138
  -> TalkingHeadBase.forward                     inferno/models/talkinghead/TalkingHeadBase.py:503-553
       FlamePreprocessor (gt_exp/gt_jaw/gt_shape -> gt_vertices, template)   inferno/models/temporal/Preprocessors.py:62-186
       Wav2Vec2Encoder (z-norm + Wav2Vec2ModelResampled)                       inferno/models/temporal/AudioEncoders.py:38-90,168-201
       LinearSequenceEncoder                                                   inferno/models/temporal/SequenceEncoders.py:180-197
       BertPriorDecoder (style add, encoder layer, decoder, stack-linear squash, L2L motion-prior decoder, FLAME, offsets)
                                                                               inferno/models/talkinghead/FaceFormerDecoder.py:987-1224
       L2lDecoder.forward / MotionPrior.decoding_step                          inferno/models/temporal/motion_prior/{L2lMotionPrior.py:460-495,MotionPrior.py:316-380}

The module tree (and therefore every ``state_dict`` key below ``sequence_encoder`` / ``sequence_decoder`` / ``audio_model.model``)
mirrors the reference, so an EMOTE checkpoint loads with ``load_state_dict``. torch modules are parameter containers only; the
arithmetic runs in libavi_b200.so and is batched over clips (the reference is called with B = 1, evaluation_functions.py:381-383;
the audio z-norm is therefore applied per clip). Renderers, losses, texture and Lightning hooks are out of scope
(``render_results=False`` at train_diffusion_prior.py:956). There is no CPU path.
"""
from __future__ import annotations

import math
import os
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import ops
from .faceformer import get_slopes
from .ops import ACT_GELU, ACT_NONE
from .wav2vec import Wav2Vec2Model


def default_precision() -> str:
    p = os.environ.get("AVI_B200_PRECISION", "bf16").lower()
    if p not in ("bf16", "fp32"):
        raise ValueError("AVI_B200_PRECISION must be bf16 or fp32")
    return p


# ------------------------------------------------------------------------------------------------ parameter containers
class Wav2Vec2Encoder(nn.Module):
    """AudioEncoders.py:130-241: holds ``model`` (Wav2Vec2ModelResampled there, the B200 Wav2Vec2Model here)."""

    def __init__(self, model: Wav2Vec2Model):
        super().__init__()
        self.model = model

    def output_feature_dim(self):
        return self.model.config.hidden_size


class LinearSequenceEncoder(nn.Module):
    def __init__(self, input_feature_dim=768, feature_dim=128):
        super().__init__()
        self.linear = nn.Linear(input_feature_dim, feature_dim)


class LinearEmotionCondition(nn.Module):
    """FaceFormerDecoder.py:129-267 with the EMOTE flags: per-frame one-hots of expression label, intensity and identity."""

    def __init__(self, cfg, output_dim):
        super().__init__()
        self.cfg = cfg
        self.condition_dim = cfg.n_expression + cfg.n_intensities + cfg.n_identities
        self.map = nn.Linear(self.condition_dim, output_dim, bias=getattr(cfg, "use_bias", True))

    def gather_condition(self, sample, T):
        """:166-254 for gt_expression_label / gt_expression_intensity / gt_expression_identity."""
        parts = []
        for key, idx_key, n, off in (("gt_expression_label_condition", "gt_expression_label", self.cfg.n_expression, 0),
                                     ("gt_expression_intensity_condition", "gt_expression_intensity", self.cfg.n_intensities, 1),
                                     ("gt_expression_identity_condition", "gt_expression_identity", self.cfg.n_identities, 0)):
            if key in sample:
                c = sample[key]
            else:
                c = torch.nn.functional.one_hot(sample[idx_key] - off, num_classes=n)
            if c.ndim == 2:
                c = c.unsqueeze(1)
            if c.shape[1] == 1:
                c = c.expand(-1, T, -1)
            parts.append(c.to(dtype=torch.float32))
        return torch.cat(parts, dim=-1)


class StackLinearSquash(nn.Module):
    def __init__(self, input_dim, latent_frame_size, output_dim):
        super().__init__()
        self.input_dim, self.latent_frame_size, self.output_dim = input_dim, latent_frame_size, output_dim
        self.linear = nn.Linear(input_dim * latent_frame_size, output_dim)


class L2lDecoder(nn.Module):
    """L2lMotionPrior.py:361-455 (l2l_decoder.yaml: feature_dim 256, 8 heads, ff 384, gelu, alibi_future, quant_factor 3)."""

    def __init__(self, feature_dim=256, nhead=8, intermediate_size=384, quant_factor=3, out_dim=53):
        super().__init__()
        d = feature_dim
        self.expander = nn.ModuleList([nn.Sequential(nn.ConvTranspose1d(d, d, 5, stride=2, padding=2, output_padding=1),
                                                     nn.LeakyReLU(0.2, True), nn.BatchNorm1d(d))])
        for _ in range(1, quant_factor):
            self.expander.append(nn.Sequential(nn.Conv1d(d, d, 5, stride=1, padding=2, padding_mode="replicate"),
                                               nn.LeakyReLU(0.2, True), nn.BatchNorm1d(d)))
        layer = nn.TransformerEncoderLayer(d_model=d, nhead=nhead, dim_feedforward=intermediate_size, activation="gelu", dropout=0.0,
                                           batch_first=True)
        self.decoder_transformer = nn.TransformerEncoder(layer, num_layers=1, enable_nested_tensor=False)
        self.decoder_linear_embedding = nn.Linear(d, d)
        self.cross_smooth_layer = nn.Conv1d(d, out_dim, 5, padding=2)
        self.nhead, self.quant_factor = nhead, quant_factor


class MotionPrior(nn.Module):
    def __init__(self, motion_decoder: L2lDecoder, flame):
        super().__init__()
        self.motion_decoder = motion_decoder
        self._flame = [flame]          # not registered twice: the decoder owns it (sequence_decoder.flame)
        self.cfg = SimpleNamespace(model=SimpleNamespace(sequence_components={"exp": 50, "jaw": "rot"}, rotation_representation="aa"))

    def get_flame(self):
        return self._flame[0]

    def latent_frame_size(self):
        return 2 ** self.motion_decoder.quant_factor


class BertPriorDecoder(nn.Module):
    def __init__(self, cfg, flame):
        super().__init__()
        self.cfg = cfg
        self.obj_vector = LinearEmotionCondition(cfg.style_embedding, cfg.feature_dim)
        layer = nn.TransformerEncoderLayer(d_model=cfg.feature_dim, nhead=cfg.nhead, dim_feedforward=cfg.feature_dim, activation="gelu",
                                           dropout=0.25, batch_first=True)
        self.bert_decoder = nn.TransformerEncoder(layer, num_layers=1, enable_nested_tensor=False)
        self.flame = flame
        self.motion_prior = MotionPrior(L2lDecoder(), flame)
        self.latent_frame_size = self.motion_prior.latent_frame_size()
        self.decoder = nn.Linear(cfg.feature_dim, 256)
        nn.init.constant_(self.decoder.weight, 0)          # FaceFormerDecoder.py:1050-1051
        nn.init.constant_(self.decoder.bias, 0)
        self.squasher_2 = StackLinearSquash(256, self.latent_frame_size, 256)

    def get_shape_model(self):
        return self.motion_prior.get_flame()


class AttrDict(dict):
    """cfg.yaml as the callers read it: nested attribute AND item access (the reference uses an OmegaConf DictConfig, which is not a
    dependency here; `cfg.model.sequence_decoder.style_embedding.n_identities`, `cfg.data.split`, `cfg.learning.losses = {}` all work)."""

    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            return AttrDict(v)
        if isinstance(v, (list, tuple)):
            return type(v)(AttrDict._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, AttrDict._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, AttrDict) else v) for k, v in self.items()}


def emote_cfg(n_identities=32, n_expression=8, n_intensities=3, flame=None, checkpoint_dir=None):
    """The slice of the EMOTE cfg.yaml this path reads (talkinghead_conf/bertprior_wild.yaml and its model/* groups): decoder sizes and
    style embedding, the audio model specifier, FLAME asset paths (`flame`: an object with flame_model_path / flame_lmk_embedding_path /
    n_shape / n_exp) and inout.checkpoint_dir."""
    style = dict(type="emotion_linear", n_expression=n_expression, n_intensities=n_intensities, n_identities=n_identities,
                 gt_expression_label=True, gt_expression_intensity=True, gt_expression_identity=True, use_bias=True)
    dec = dict(type="BertPriorDecoder", feature_dim=128, nhead=8, num_layers=1, activation="gelu", post_bug_fix=True,
               squash_after=True, squash_type="stack_linear", style_op="add", style_embedding=style)
    if flame is not None:
        dec["flame"] = dict(flame_model_path=str(flame.flame_model_path), flame_lmk_embedding_path=str(flame.flame_lmk_embedding_path),
                            n_shape=int(flame.n_shape), n_exp=int(flame.n_exp))
    data = dict(data_class="MEADPseudo3DDM", split="random_by_identityV2_sorted_70_15_15", reconstruction_type=["EMICA-MEAD_flame2020"])
    cfg = dict(model=dict(sequence_decoder=dec, audio=dict(type="wav2vec2", model_specifier="facebook/wav2vec2-base-960h", trainable=True),
                          sequence_encoder=dict(type="LinearSequenceEncoder", feature_dim=128)),
               data=data, learning=dict(losses={}, metrics={}), inout=dict(checkpoint_dir=checkpoint_dir or "checkpoints"))
    return AttrDict(cfg)


# ------------------------------------------------------------------------------------------------ checkpoint directories
def get_path_to_assets():
    """inferno/utils/other.py get_path_to_assets: the directory relative checkpoint_dirs are resolved against (INFERNO_ASSETS, else
    ./assets)."""
    from pathlib import Path
    return Path(os.environ.get("INFERNO_ASSETS", "assets"))


def locate_checkpoint(cfg_or_checkpoint_dir, replace_root=None, relative_to=None, mode=None, pattern=None):
    """inferno/models/IO.py:26-90, same contract: the checkpoint file chosen from `cfg.inout.checkpoint_dir` (or a directory given
    directly) - mode 'latest' = the first *.ckpt in sorted order, which must be called last.ckpt; 'best' = the smallest value after the
    last '=' of the file stem; an int indexes the sorted list. Returns None when nothing qualifies (upstream prints and returns None)."""
    from pathlib import Path
    checkpoint_dir = str(cfg_or_checkpoint_dir) if isinstance(cfg_or_checkpoint_dir, (str, Path)) else cfg_or_checkpoint_dir.inout.checkpoint_dir
    if replace_root is not None and relative_to is not None:
        try:
            checkpoint_dir = str(Path(replace_root) / Path(checkpoint_dir).relative_to(relative_to))
        except ValueError:
            pass
    if not Path(checkpoint_dir).is_absolute():
        checkpoint_dir = str(get_path_to_assets() / checkpoint_dir)
    checkpoints = sorted(Path(checkpoint_dir).rglob("*.ckpt"))
    if pattern is not None:
        checkpoints = [c for c in checkpoints if pattern in str(c)]
    if not checkpoints:
        return None
    if isinstance(mode, int):
        return str(checkpoints[mode])
    if mode == "latest":
        return str(checkpoints[0]) if checkpoints[0].name == "last.ckpt" else None
    if mode == "best":
        best, best_val = None, float("inf")
        for c in checkpoints:
            if c.stem == "last":
                continue
            try:
                val = float(c.stem[c.stem.rfind("=") + 1:])
            except ValueError:
                continue
            if val <= best_val:
                best, best_val = c, val
        if best is None:
            raise FileNotFoundError("Finding the best checkpoint failed")
        return str(best)
    raise ValueError(f"Invalid checkpoint loading mode '{mode}'")


def _load_cfg_yaml(path):
    import yaml
    with open(path) as fh:
        return AttrDict(yaml.safe_load(fh))


def _flame_from_cfg(cfg, run_path):
    """FLAME(n_shape, n_exp) for the decoder: asset paths from cfg.model.sequence_decoder.flame (bertprior_wild.yaml:38-44), else from
    cfg.model.preprocessor.flame (preprocessor/flame_tex.yaml). Relative paths are tried against the run directory and the assets root."""
    from pathlib import Path

    from .flame import FLAME
    fc = None
    for holder in (cfg.model.get("sequence_decoder", {}), cfg.model.get("preprocessor", {})):
        if isinstance(holder, dict) and isinstance(holder.get("flame"), dict):
            fc = holder["flame"]
            break
    if fc is None:
        raise KeyError("cfg.yaml has no FLAME section (model.sequence_decoder.flame / model.preprocessor.flame)")

    def resolve(p):
        p = Path(p)
        for cand in ([p] if p.is_absolute() else [Path(run_path) / p, get_path_to_assets() / p, p]):
            if cand.exists():
                return str(cand)
        raise FileNotFoundError(f"FLAME asset '{p}' named by cfg.yaml does not exist (looked under the run directory and "
                                f"{get_path_to_assets()}; set INFERNO_ASSETS)")

    return FLAME(SimpleNamespace(flame_model_path=resolve(fc["flame_model_path"]),
                                 flame_lmk_embedding_path=resolve(fc["flame_lmk_embedding_path"]),
                                 n_shape=int(fc.get("n_shape", 300)), n_exp=int(fc.get("n_exp", 50))))


def _audio_from_cfg(cfg, state_dict):
    """The wav2vec2 encoder named by cfg.model.audio.model_specifier (AudioEncoders.py:130-166). Its weights are in the checkpoint (the
    EMOTE audio model is trainable); the hub is only asked for the ARCHITECTURE, and only from the local cache (no network needed)."""
    from transformers import Wav2Vec2Config
    spec = cfg.model.get("audio", {}).get("model_specifier", "facebook/wav2vec2-base-960h")
    try:
        config = Wav2Vec2Config.from_pretrained(spec, local_files_only=True)
    except Exception:  # noqa: BLE001 - not cached: the named model must then be the base architecture the checkpoint tensors describe
        config = Wav2Vec2Config()
        n_layers = 1 + max((int(k.split(".")[4]) for k in state_dict if k.startswith("audio_model.model.encoder.layers.")), default=-1)
        width = next((v.shape[0] for k, v in state_dict.items() if k == "audio_model.model.encoder.layer_norm.weight"), config.hidden_size)
        if (n_layers and n_layers != config.num_hidden_layers) or width != config.hidden_size:
            raise RuntimeError(f"'{spec}' is not in the local transformers cache and the checkpoint is not wav2vec2-base shaped "
                               f"({n_layers} layers, width {width}); cache the model's config.json first")
    return Wav2Vec2Model(config)


def load_model(path_to_models, run_name, mode="latest", with_losses=True):
    """inferno_apps/TalkingHead/utils/load.py:28-61 (load_model + load_faceformer + LightningModule.load_from_checkpoint(strict=False)):
    reads <path_to_models>/<run_name>/cfg.yaml, locates the checkpoint, builds the model the config describes and loads the
    checkpoint's `state_dict`. Returns (model, cfg). Where upstream calls sys.exit(0) for a missing checkpoint this raises."""
    from pathlib import Path
    run_path = Path(path_to_models) / run_name
    cfg = _load_cfg_yaml(run_path / "cfg.yaml")
    if not with_losses:
        cfg.setdefault("learning", AttrDict())
        cfg.learning.losses = {}
        cfg.learning.metrics = {}
    ckpt_dir = Path(str(cfg.inout.checkpoint_dir))
    checkpoint = locate_checkpoint(cfg, mode=mode) if (ckpt_dir.is_absolute() and ckpt_dir.exists()) or (get_path_to_assets() / ckpt_dir).exists() else None
    if checkpoint is None:     # released runs carry paths of the training cluster: fall back to the run directory itself
        checkpoint = locate_checkpoint(run_path, mode=mode)
    if checkpoint is None:
        raise FileNotFoundError(f"no '{mode}' checkpoint (*.ckpt; 'latest' needs last.ckpt first in sorted order) under "
                                f"'{cfg.inout.checkpoint_dir}' or '{run_path}'")
    blob = torch.load(checkpoint, map_location="cpu", weights_only=False)
    state = blob["state_dict"] if isinstance(blob, dict) and "state_dict" in blob else blob
    dec_type = cfg.model.sequence_decoder.get("type")
    if dec_type != "BertPriorDecoder":
        raise NotImplementedError(f"sequence_decoder.type '{dec_type}': only the EMOTE BertPriorDecoder stack is built")
    model = TalkingHeadModel(cfg, _audio_from_cfg(cfg, state), _flame_from_cfg(cfg, run_path))
    # strict=False as upstream (TalkingHead/utils/load.py:60): renderer / loss / texture tensors of the checkpoint have no module here.
    # What must NOT be missing are this model's own trainable tensors.
    missing, unexpected = model.load_state_dict(state, strict=False)
    own_missing = [k for k in missing if ".flame." not in k and not k.endswith("num_batches_tracked")]
    if own_missing:
        raise RuntimeError(f"checkpoint '{checkpoint}' lacks {len(own_missing)} tensors of the talking-head model, e.g. {own_missing[:5]}")
    model.checkpoint_path, model.unexpected_keys = checkpoint, unexpected
    return model, cfg


# ------------------------------------------------------------------------------------------------ the model
class TalkingHeadModel(nn.Module):
    """TalkingHeadBase: audio_model, sequence_encoder, sequence_decoder (+ FLAME as the preprocessor)."""

    max_seq_length = 5000                                                                         # TalkingHeadBase.py:87-89

    def __init__(self, cfg, audio_encoder: Wav2Vec2Model, flame):
        super().__init__()
        self.cfg = cfg
        self.audio_model = Wav2Vec2Encoder(audio_encoder)
        self.sequence_encoder = LinearSequenceEncoder(768, cfg.model.sequence_decoder.feature_dim)
        self.sequence_decoder = BertPriorDecoder(cfg.model.sequence_decoder, flame)
        self.renderer, self.neural_losses = None, {}
        self.precision = default_precision()
        self._packed, self._packed_key = None, None

    # -- packing -------------------------------------------------------------------------------------------
    def _own_tensors(self):
        return [t for n, t in list(self.named_parameters()) + list(self.named_buffers())
                if not n.startswith("audio_model.") and ".flame." not in n]

    @torch.no_grad()
    def _pack(self):
        key = (self.precision,) + tuple((t.data_ptr(), t._version) for t in self._own_tensors())
        if self._packed is not None and key == self._packed_key:
            return self._packed
        bf16 = self.precision == "bf16"
        wdt = (lambda t: ops.cast_bf16(t)) if bf16 else (lambda t: t.detach().float().contiguous())
        f32 = lambda t: t.detach().float().contiguous()  # noqa: E731

        def layer(lyr):
            return dict(qkv_w=wdt(lyr.self_attn.in_proj_weight), qkv_b=f32(lyr.self_attn.in_proj_bias),
                        o_w=wdt(lyr.self_attn.out_proj.weight), o_b=f32(lyr.self_attn.out_proj.bias),
                        ff1_w=wdt(lyr.linear1.weight), ff1_b=f32(lyr.linear1.bias), ff2_w=wdt(lyr.linear2.weight), ff2_b=f32(lyr.linear2.bias),
                        ln1=(f32(lyr.norm1.weight), f32(lyr.norm1.bias)), ln2=(f32(lyr.norm2.weight), f32(lyr.norm2.bias)))

        dec = self.sequence_decoder
        l2l = dec.motion_prior.motion_decoder
        P = dict(enc_w=wdt(self.sequence_encoder.linear.weight), enc_b=f32(self.sequence_encoder.linear.bias),
                 style_w=f32(dec.obj_vector.map.weight), style_b=None if dec.obj_vector.map.bias is None else f32(dec.obj_vector.map.bias),
                 bert=layer(dec.bert_decoder.layers[0]), dec_w=wdt(dec.decoder.weight), dec_b=f32(dec.decoder.bias),
                 sq_w=wdt(dec.squasher_2.linear.weight), sq_b=f32(dec.squasher_2.linear.bias),
                 l2l=layer(l2l.decoder_transformer.layers[0]),
                 emb_w=wdt(l2l.decoder_linear_embedding.weight), emb_b=f32(l2l.decoder_linear_embedding.bias),
                 slopes=torch.tensor(get_slopes(l2l.nhead), dtype=torch.float32, device=dec.decoder.weight.device))
        P["exp"] = []
        for i, seq in enumerate(l2l.expander):
            conv, bn = seq[0], seq[2]
            w = conv.weight.detach().float()
            if isinstance(conv, nn.ConvTranspose1d):
                # ConvTranspose1d(k5, s2, p2, op1) == zero insertion + correlation with the flipped kernel, in/out swapped:
                # W_eq[co, j, ci] = w[ci, co, 4 - j]
                w = w.flip(2).permute(1, 2, 0)
            else:
                w = w.permute(0, 2, 1)                                                           # [co, j, ci] tap-major
            scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float().contiguous()
            shift = (bn.bias - bn.running_mean * scale).detach().float().contiguous()
            P["exp"].append((wdt(w.reshape(w.shape[0], -1).contiguous()), f32(conv.bias), scale, shift))
        cs = l2l.cross_smooth_layer
        P["cs_w"] = wdt(cs.weight.detach().float().permute(0, 2, 1).reshape(cs.weight.shape[0], -1).contiguous())
        P["cs_b"] = f32(cs.bias)
        self._packed, self._packed_key = P, key
        return P

    # -- pieces --------------------------------------------------------------------------------------------
    def _lin(self, x32, x16, w, b, *, act=ACT_NONE, residual=None, want32=True, want16=False, **kw):
        """nn.Linear on [rows, K]: bf16 mode feeds the tcgen05 GEMM with the bf16 copy, fp32 mode the CUDA-core GEMM."""
        bf16 = self.precision == "bf16"
        a = x16 if bf16 else x32
        if bf16 and want16 and not want32:
            return None, ops.linear(a, w, b, act=act, residual=residual, out_dtype=torch.bfloat16)
        if bf16 and want16:
            return ops.linear(a, w, b, act=act, residual=residual, out_dtype=torch.float32, out2_dtype=torch.bfloat16)
        return ops.linear(a, w, b, act=act, residual=residual, out_dtype=torch.float32), None

    def _conv5(self, x32, B, L_in, mode, front, Lp, rows, w, b, N):
        """5-tap Conv1d over time as a conv-mode GEMM on a staged (padded / zero-inserted) copy of x [B, L_in, C]."""
        bf16 = self.precision == "bf16"
        Cc = x32.shape[-1]
        xs = ops.stage_rows(x32.reshape(B, L_in, Cc), B, L_in, Lp, front, mode, dtype=torch.bfloat16 if bf16 else torch.float32)
        out = torch.empty((B, rows, N), dtype=torch.float32, device=x32.device)
        ops.gemm(xs, w, b, out, batch=B, rows=rows, N=N, K=5 * Cc, conv_taps=5, conv_stride=1, a_ld=Cc, a_batch_stride=Lp * Cc,
                 a_rows_alloc=Lp, c_ld=N, c_batch_stride=rows * N)
        return out

    def _encoder_layer(self, x32, x16, L, B, T, nhead, slopes):
        """nn.TransformerEncoderLayer (post-LN, GELU, eval)."""
        bf16 = self.precision == "bf16"
        D = x32.shape[-1] // nhead
        qkv, _ = self._lin(x32, x16, L["qkv_w"], L["qkv_b"])
        a32, a16 = ops.mha_small(qkv, B, T, nhead, D, slopes=slopes, want_bf16=bf16)
        y, _ = self._lin(a32, a16, L["o_w"], L["o_b"], residual=x32)
        h32, h16 = ops.layernorm(y, *L["ln1"], want_bf16=bf16)
        f32_, f16 = self._lin(h32, h16, L["ff1_w"], L["ff1_b"], act=ACT_GELU, want32=not bf16, want16=bf16)
        y2, _ = self._lin(f32_, f16, L["ff2_w"], L["ff2_b"], residual=h32)
        return ops.layernorm(y2, *L["ln2"], want_bf16=bf16)

    def _flame_sequence(self, shape, exp, jaw):
        """FlamePreprocessor._forward (Preprocessors.py:62-186): vertices [B,T,V*3] for per-frame exp / jaw, pose = [0,0,0,jaw]."""
        flame = self.sequence_decoder.flame
        pose = torch.cat([torch.zeros_like(jaw), jaw], dim=-1)
        return flame.vertices_sequence(shape, exp.contiguous(), pose)

    def _neutral(self, shape):
        flame = self.sequence_decoder.flame
        return flame.vertices_only(shape, torch.zeros(shape.shape[0], flame.cfg.n_exp, device=shape.device), None).reshape(shape.shape[0], -1).contiguous()   # dense rows for avi_sub_add_rows (the tensor-core FLAME path returns a padded row pitch)

    def style_embedding(self, sample, T):
        """LinearEmotionCondition.forward (FaceFormerDecoder.py:256-267): [B,T,cond] -> [B,T,128] (fp32 GEMM: K = 43)."""
        P = self._pack()
        cond = self.sequence_decoder.obj_vector.gather_condition(sample, T).contiguous()
        B = cond.shape[0]
        return ops.linear(cond.reshape(B * T, -1).contiguous(), P["style_w"], P["style_b"]).view(B, T, -1)

    # -- forward -------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, sample, style_emb=None, train=False, validation=False, only_style_emb=False, is_external_style_emb=False, **kw):
        if train or validation:
            raise NotImplementedError("training / validation passes (disentangle, losses) are outside the inference hot path")
        raw = sample["raw_audio"]
        if not raw.is_cuda:
            raise RuntimeError("avi_talking_b200 talking head runs on CUDA only (no CPU fallback)")
        if "gt_shape" not in sample or sample["gt_shape"].ndim != 2:
            raise NotImplementedError("gt_shape [B, n_shape] is required (as create_base_sample provides, evaluation_functions.py:155); "
                                      "the template-mesh fallback of FlamePreprocessor is not implemented")
        B, T = raw.shape[:2]
        if T > self.max_seq_length:
            raise NotImplementedError("sequence longer than max_seq_length")
        if only_style_emb:                                                                        # FaceFormerDecoder.py:599-601
            return self.style_embedding(sample, T)
        P = self._pack()
        bf16 = self.precision == "bf16"
        shape = sample["gt_shape"].float()
        # 1. preprocess: pseudo-GT vertices and the per-clip template (TalkingHeadBase.py:516)
        sample["gt_vertices"] = self._flame_sequence(shape, sample["gt_exp"].float(), sample["gt_jaw"].float())
        neutral = self._neutral(shape)
        sample["template"] = neutral
        # 2. audio (AudioEncoders.py:168-201): z-norm, wav2vec2 resampled to T frames
        audio = ops.audio_znorm(raw.reshape(B, -1))
        sample["processed_audio"] = audio
        w2v = self.audio_model.model
        feat = w2v(audio, frame_num=T).last_hidden_state
        sample["audio_feature"] = feat
        feat16 = w2v.last_hidden_state_bf16.reshape(B * T, -1) if bf16 else None
        # 3. sequence encoder + style (SequenceEncoders.py:189-197, FaceFormerDecoder.py:652-670)
        if not (style_emb is not None and is_external_style_emb):
            style_emb = self.style_embedding(sample, T)
        style = style_emb.float().contiguous()
        if style.ndim == 2:
            style = style[:, None]
        Cf = style.shape[-1]
        h32 = torch.empty((B * T, Cf), dtype=torch.float32, device=raw.device)
        h16 = torch.empty((B * T, Cf), dtype=torch.bfloat16, device=raw.device) if bf16 else None
        ops.gemm(feat16 if bf16 else feat.reshape(B * T, -1), P["enc_w"], P["enc_b"], h32, batch=B, rows=T, N=Cf, K=feat.shape[-1],
                 residual=style, out2=h16, a_ld=feat.shape[-1], a_batch_stride=T * feat.shape[-1], a_rows_alloc=T, c_ld=Cf,
                 c_batch_stride=T * Cf, res_ld=0 if style.shape[1] == 1 else Cf, res_batch_stride=style.shape[1] * Cf)
        # (the un-styled "seq_encoder_output" is not materialised: the style add is fused into the encoder GEMM as its residual)
        # 4. BertPriorDecoder._decode (FaceFormerDecoder.py:1194-1224)
        dec = self.sequence_decoder
        x32, x16 = self._encoder_layer(h32, h16, P["bert"], B, T, dec.cfg.nhead, None)
        z, _ = self._lin(x32, x16, P["dec_w"], P["dec_b"])
        # 5. _apply_motion_prior (:1104-1182): pad to a multiple of the latent frame size, stack-linear squash, L2L decoder
        fs = dec.latent_frame_size
        Tp = int(math.ceil(T / fs) * fs)
        zp = ops.stage_rows(z.view(B, T, -1), B, T, Tp, 0, 0, dtype=torch.bfloat16 if bf16 else torch.float32)
        L = Tp // fs
        zs = zp.view(B * L, -1)
        lat, _ = self._lin(zs, zs, P["sq_w"], P["sq_b"])
        sample["prior_input_sequence"] = lat.view(B, L, -1)
        x = lat.view(B, L, -1)
        for i, (w, b, scale, shift) in enumerate(P["exp"]):                                       # L2lMotionPrior.py:465-470
            Lc = x.shape[1]
            if i == 0:
                y = self._conv5(x, B, Lc, 2, 2, 2 * Lc + 4, 2 * Lc, w, b, w.shape[0])
                x = ops.lrelu_bn_repeat(y, scale, shift, B, 2 * Lc, 1)
            else:
                y = self._conv5(x, B, Lc, 1, 2, Lc + 4, Lc, w, b, w.shape[0])
                x = ops.lrelu_bn_repeat(y, scale, shift, B, Lc, 2)
        assert x.shape[1] == Tp
        x2 = x.reshape(B * Tp, -1)
        e32, e16 = self._lin(x2, ops.cast_bf16(x2) if bf16 else None, P["emb_w"], P["emb_b"], want16=bf16)   # :472
        l2l = dec.motion_prior.motion_decoder
        t32, _ = self._encoder_layer(e32, e16, P["l2l"], B, Tp, l2l.nhead, P["slopes"])          # :477-483
        seq = self._conv5(t32, B, Tp, 0, 2, Tp + 4, Tp, P["cs_w"], P["cs_b"], P["cs_w"].shape[0])[:, :T]   # :488, crop :1143-1146
        exp, jaw = seq[..., :50].contiguous(), seq[..., 50:53].contiguous()                       # MotionPrior.py:316-329
        sample["predicted_exp"], sample["predicted_jaw"] = exp, jaw
        # 6. FLAME on the predicted coefficients, offsets from the neutral shape, + template (MotionPrior.py:331-351, :1163-1175, :690-694)
        verts = self._flame_sequence(shape, exp, jaw)
        sample["predicted_vertices"] = ops.sub_add_rows_(verts, neutral, sample["template"])
        return sample


class TalkingHeadWrapper(nn.Module):
    """inferno_apps/TalkingHead/evaluation/TalkingHeadWrapper.py:76-138, same constructor: ``TalkingHeadWrapper(path_to_model,
    render_results=False)`` opens the run directory (cfg.yaml + a Lightning ``last.ckpt``) exactly as train_diffusion_prior.py:954-958
    does. Rendering is out of scope: ``render_results=True`` (the upstream default) raises - the inference script passes False.
    ``TalkingHeadWrapper.from_parts(audio_encoder, flame, cfg)`` assembles the same object from modules already in memory."""

    def __init__(self, path_to_model, render_results=True, use_preprocessor=True, apply_mask=True):
        super().__init__()
        if render_results:
            raise NotImplementedError("rendering (FixedViewFlameRenderer / pytorch3d) is out of scope of the audio -> FLAME path; "
                                      "construct with render_results=False as train_diffusion_prior.py:956 does")
        from pathlib import Path
        self.dim = 128                 # style-code width of EMOTE (TalkingHeadWrapper.py:80)
        self.self_cond = None
        path_to_model = Path(path_to_model)
        self.talking_head_model, self.cfg = load_model(path_to_model.parent, path_to_model.name, mode="latest", with_losses=False)
        self.talking_head_model.eval()
        self.renderer = None
        self.render_results, self.use_preprocessor, self.apply_mask = render_results, use_preprocessor, apply_mask

    @classmethod
    def from_parts(cls, audio_encoder: Wav2Vec2Model, flame, cfg=None, use_preprocessor=True, apply_mask=True):
        """The same wrapper around modules already in memory (no run directory): tests, smoke, benchmarks with synthetic weights."""
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        self.dim, self.self_cond = 128, None
        self.cfg = cfg if cfg is not None else emote_cfg()
        self.talking_head_model = TalkingHeadModel(self.cfg, audio_encoder, flame)
        self.renderer = None
        self.render_results, self.use_preprocessor, self.apply_mask = False, use_preprocessor, apply_mask
        return self

    def set_neutral_mesh(self, neutral_v):
        """TalkingHeadWrapper.py:140-158: overwrite the FLAME template in place (every FLAME instance of the reference's model; here
        the decoder, the motion prior and the preprocessor share ONE)."""
        self.talking_head_model.sequence_decoder.get_shape_model().v_template[...] = neutral_v

    def get_num_intensities(self):
        return self.cfg.model.sequence_decoder.style_embedding.n_intensities

    def get_num_emotions(self):
        return self.cfg.model.sequence_decoder.style_embedding.n_expression

    def get_num_identities(self):
        return self.cfg.model.sequence_decoder.style_embedding.n_identities

    MEAD_IDENTITIES = ("M003 M005 M007 M009 M011 M012 M013 M019 M022 M023 M024 M025 M026 M027 M028 M029 M030 M031 M032 M033 M034 M035 "
                       "M037 M039 M040 M041 M042 W009 W011 W014 W015 W016 W017 W018 W019 W021 W023 W024 W025 W026 W028 W029 W033 W035 "
                       "W036 W037 W038 W040")

    def get_subject_labels(self, train_val_test):
        """TalkingHeadWrapper.py:168-236: the MEAD identity split named by cfg.data.split (male and female lists are cut separately;
        upstream's `random` shuffle is applied to a list that is not used afterwards, so both variants give the sorted split)."""
        assert self.cfg.data.data_class == "MEADPseudo3DDM"
        assert "random_by_identityV2" in self.cfg.data.split
        res = self.cfg.data.split.split("_")
        assert res[3] in ("random", "sorted"), f"Unknown random_or_sorted value: '{res[3]}'"
        train, val, test = float(res[-3]), float(res[-2]), float(res[-1])
        train_, val_ = train / (train + val + test), val / (train + val + test)
        ids = sorted(TalkingHeadWrapper.MEAD_IDENTITIES.split())
        male, female = [i for i in ids if i.startswith("M")], [i for i in ids if i.startswith("W")]
        out = {"training": [], "validation": [], "testing": []}
        for grp in (male, female):
            a, b = int(len(grp) * train_), int(len(grp) * (train_ + val_))
            out["training"] += grp[:a]
            out["validation"] += grp[a:b]
            out["testing"] += grp[b:]
        if train_val_test not in out:
            raise RuntimeError(f"Unknown set_type: '{train_val_test}'")
        return out[train_val_test]

    def forward(self, sample, style_emb=None, only_style_emb=False, is_external_style_emb=False):
        return self.talking_head_model(sample, style_emb=style_emb, only_style_emb=only_style_emb,
                                       is_external_style_emb=is_external_style_emb)
