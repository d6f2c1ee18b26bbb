"""The teacher-forced faceformer_vert training step (BASELINE configs[4]) on libavi_b200.so.

Reference: models/faceformer_vert.py  Faceformer.forward_switch_frame :360-371 (style, wav2vec2 with frame_num, audio_feature_map),
:405-412 (gt_verts = convert_coeff2verts(coeff[..., :53], pose, zeros)), :437-454 (teacher-forced decoder), :475-482 (loss =
mean(criterion(out + template, gt)) * 10); the conv feature extractor is frozen (:154). No trainer for this class is published
(SURVEY 3.3): Adam (torch.optim.Adam semantics) and the data-parallel gradient all-reduce are assembled here.

Design (B200-first, not a port of autograd):
  * forward keeps exactly the tensors the hand-written backward needs (pre-LayerNorm sums, pre-activations, attention
    probabilities); all clips of the step are batched, rows = clips x frames;
  * every dense contraction, forward and backward (dX = dY W, dW = dY^T X), runs on the tcgen05 GEMM (bf16 operands, fp32
    accumulation; `precision="fp32"` routes the same calls to the CUDA-core GEMM for the <=1e-5 parity mode). Operands whose
    contraction dimension is not innermost are laid out by avi_transpose_cast; 15069 is padded to 15104 (= 236 x 64);
  * gradients are written straight into ONE flat fp32 buffer laid out in backward-completion order (`FlatLayout`), so that the
    data-parallel all-reduce is a few large NCCL calls on contiguous ranges launched while backward is still running
    (`GradBuckets`), and Adam is ONE kernel over the flat parameter / gradient / moment buffers (`FlatAdam`);
  * q/k/v weights of an encoder layer are adjacent in the flat buffers: the fused [2304, 768] QKV weight and its gradient are
    views, no concatenation per step.

Regularisers. Without `reg` the step is the deterministic .eval() function (tests/golden/train.npz). With `reg` it is the reference's
TRAIN mode: dropout 0.1 in wav2vec2 (hidden / attention / activation, optional feat_proj), the PPE and the five sites of
nn.TransformerDecoderLayer, SpecAugment along time (models/lib/wav2vec.py:16-63,120-131) and LayerDrop - every draw an INPUT tensor
(`synth.train_regularisers` for the seeded test case, `draw_regularisers` on the device for real runs), so that the step stays a
deterministic function the oracle and tests/golden/train_reg.npz (the reference in .train() mode with the same draws injected) pin.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops
from .ops import ACT_GELU, ACT_RELU

V3_PAD = 64  # contraction dims are padded to a multiple of this (tcgen05 GEMM k-block)


def _pad(n, m=V3_PAD):
    return ((n + m - 1) // m) * m


# ------------------------------------------------------------------------------------------------ flat parameter / gradient layout
class FlatLayout:
    """Order and offsets of the trainable parameters inside the flat buffers. Order = the order in which backward finishes
    their gradients (head and decoder first, wav2vec2 layer 11 .. 0, positional conv, feature projection last), so that a
    bucket of the all-reduce is a contiguous range that is complete once backward has passed its last member."""

    def __init__(self, model):
        w2v = model.audio_encoder
        named = dict(model.named_parameters())
        lyr = "transformer_decoder.layers.0."
        order = ["vertice_map_r.weight", "vertice_map_r.bias"]
        order += [lyr + n for n in ("norm3.weight", "norm3.bias", "linear2.weight", "linear2.bias", "linear1.weight", "linear1.bias",
                                    "norm2.weight", "norm2.bias", "multihead_attn.out_proj.weight", "multihead_attn.out_proj.bias",
                                    "multihead_attn.in_proj_weight", "multihead_attn.in_proj_bias", "norm1.weight", "norm1.bias",
                                    "self_attn.out_proj.weight", "self_attn.out_proj.bias", "self_attn.in_proj_weight",
                                    "self_attn.in_proj_bias")]
        order += ["vertice_map.weight", "vertice_map.bias", "obj_vector.weight", "audio_feature_map.weight", "audio_feature_map.bias"]
        self.segments = [len(order)]                        # parameter-count boundaries where a bucket may end
        for l in reversed(range(len(w2v.encoder.layers))):
            p = f"audio_encoder.encoder.layers.{l}."
            order += [p + n for n in ("final_layer_norm.weight", "final_layer_norm.bias", "feed_forward.output_dense.weight",
                                      "feed_forward.output_dense.bias", "feed_forward.intermediate_dense.weight",
                                      "feed_forward.intermediate_dense.bias", "layer_norm.weight", "layer_norm.bias",
                                      "attention.out_proj.weight", "attention.out_proj.bias",
                                      "attention.q_proj.weight", "attention.k_proj.weight", "attention.v_proj.weight",
                                      "attention.q_proj.bias", "attention.k_proj.bias", "attention.v_proj.bias")]
            self.segments.append(len(order))
        pc = "audio_encoder.encoder.pos_conv_embed.conv."
        g_name = pc + ("parametrizations.weight.original0" if pc + "parametrizations.weight.original0" in named else "weight_g")
        v_name = pc + ("parametrizations.weight.original1" if pc + "parametrizations.weight.original1" in named else "weight_v")
        self.pos_g, self.pos_v = g_name, v_name
        order += ["audio_encoder.encoder.layer_norm.weight", "audio_encoder.encoder.layer_norm.bias", pc + "bias", g_name, v_name,
                  "audio_encoder.feature_projection.projection.weight", "audio_encoder.feature_projection.projection.bias",
                  "audio_encoder.feature_projection.layer_norm.weight", "audio_encoder.feature_projection.layer_norm.bias"]
        if "audio_encoder.masked_spec_embed" in named and named["audio_encoder.masked_spec_embed"].requires_grad:
            order.append("audio_encoder.masked_spec_embed")     # gradient only under SpecAugment (train mode); zero otherwise
        self.segments.append(len(order))
        missing = [n for n in order if n not in named]
        if missing:
            raise RuntimeError(f"FlatLayout: model lacks parameters {missing}")
        # everything else that requires grad gets no gradient from this step (obj_embedding, and for the disentangle variant
        # v_merge2hidden / learnable_eye_embed, which forward_switch_frame never touches)
        self.unused = [n for n, p in named.items() if p.requires_grad and n not in order]
        frozen = [n for n in order if not named[n].requires_grad]
        if frozen:
            raise RuntimeError(f"FlatLayout: parameters on the trained path are frozen: {frozen}")
        self.names = order
        self.offsets, off = {}, 0
        for n in order:
            self.offsets[n] = off
            k = named[n].numel()
            # keep q|k|v (weights and biases) densely adjacent; everything else starts on a 64-element (256 B) boundary
            dense = n.endswith(("attention.q_proj.weight", "attention.k_proj.weight", "attention.q_proj.bias", "attention.k_proj.bias"))
            off += k if dense else _pad(k, 64)
        self.total = off
        self.shapes = {n: tuple(named[n].shape) for n in order}

    def view(self, flat, name):
        o = self.offsets[name]
        shp = self.shapes[name]
        n = 1
        for s in shp:
            n *= s
        return flat[o:o + n].view(shp)

    def span(self, flat, first, rows, cols):
        """[rows, cols] view starting at parameter `first` (fused q|k|v)."""
        o = self.offsets[first]
        return flat[o:o + rows * cols].view(rows, cols)

    def bucket_ranges(self, max_buckets=8):
        """<= max_buckets contiguous element ranges [a, b) in backward-completion order, ending on segment boundaries."""
        ends = [(self.offsets[self.names[s]] if s < len(self.names) else self.total) for s in self.segments]
        ends[-1] = self.total
        per = -(-len(ends) // max_buckets)
        picked = [ends[min(i + per - 1, len(ends) - 1)] for i in range(0, len(ends), per)]
        ranges, a = [], 0
        for b in picked:
            if b > a:
                ranges.append((a, b))
                a = b
        return ranges


def flatten_parameters(model, layout=None):
    """Move the trainable parameters into one flat fp32 buffer (each nn.Parameter becomes a view; values preserved)."""
    layout = layout or FlatLayout(model)
    named = dict(model.named_parameters())
    dev = named[layout.names[0]].device
    flat = torch.zeros(layout.total, dtype=torch.float32, device=dev)
    with torch.no_grad():
        for n in layout.names:
            v = layout.view(flat, n)
            v.copy_(named[n].data)
            named[n].data = v
    model._flat_params, model._flat_layout = flat, layout
    return flat, layout


class GradBuckets:
    """Data-parallel gradient exchange: sum all-reduce of contiguous ranges of the flat gradient buffer, each launched
    on a communication stream as soon as backward has finished the range. The producer scales its upstream gradient by `prescale`
    (= 1 / world) so the reduced buffer IS the rank-averaged gradient, as DistributedDataParallel leaves it in `.grad`."""

    def __init__(self, layout: FlatLayout, group=None, max_buckets=8):
        self.ranges = layout.bucket_ranges(max_buckets)
        self.group = group
        self.works = []
        self.stream = None
        self._next = 0

    @property
    def world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    @property
    def prescale(self):
        return 1.0 / self.world

    def begin(self):
        self.works, self._next = [], 0

    def ready(self, flat_grad, upto):
        """Backward has finished every gradient whose flat offset is < upto: launch the buckets that are now complete."""
        if self.world == 1:
            return
        while self._next < len(self.ranges) and self.ranges[self._next][1] <= upto:
            a, b = self.ranges[self._next]
            self._next += 1
            if flat_grad.is_cuda:
                if self.stream is None:
                    self.stream = torch.cuda.Stream(device=flat_grad.device)
                self.stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self.stream):
                    self.works.append(dist.all_reduce(flat_grad[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            else:
                self.works.append(dist.all_reduce(flat_grad[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self, flat_grad):
        self.ready(flat_grad, flat_grad.numel())
        for w in self.works:
            w.wait()
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.works = []
        return 1.0      # the 1/world factor was folded into the upstream gradient (TrainStep.backward): the buffer holds the AVERAGE


def shadow_tag(model):
    """What the bf16 shadow of the flat parameter buffer is valid for: the fused optimizer's epoch and every parameter's version."""
    lay = model._flat_layout
    named = dict(model.named_parameters())
    return (ops.WEIGHT_EPOCH, model._flat_params.data_ptr()) + tuple(named[n]._version for n in lay.names)


def bf16_shadow(model):
    """bf16 copy of the flat parameter buffer (same offsets): refreshed by the Adam kernel itself, or here by one cast launch when
    the weights were changed by anything else (load_state_dict, a torch optimizer)."""
    tag = shadow_tag(model)
    if getattr(model, "_flat_params_bf16", None) is None:
        model._flat_params_bf16 = torch.empty(model._flat_params.shape, dtype=torch.bfloat16, device=model._flat_params.device)
        model._flat_params_bf16_tag = None
    if model._flat_params_bf16_tag != tag:                  # in place: a captured graph may hold views of this buffer
        ops.cast_bf16(model._flat_params, out=model._flat_params_bf16)
        model._flat_params_bf16_tag = tag
    return model._flat_params_bf16


class FlatAdam:
    """torch.optim.Adam (no weight decay, no amsgrad) as one kernel over the flat parameter / gradient / moment buffers."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
        if getattr(model, "_flat_params", None) is None:
            flatten_parameters(model)
        self.model = model
        self.p = model._flat_params
        self.m = torch.zeros_like(self.p)
        self.v = torch.zeros_like(self.p)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.t = 0

    def zero_grad(self, set_to_none=True):
        self.model._flat_grad = None
        for p in self.model.parameters():
            p.grad = None

    def step(self, grad_scale=1.0):
        g = getattr(self.model, "_flat_grad", None)
        if g is None:
            raise RuntimeError("FlatAdam.step: no gradient (run the training forward and loss.backward() first)")
        self.t += 1
        bf16 = getattr(self.model, "precision", "bf16") == "bf16"
        if bf16 and getattr(self.model, "_flat_params_bf16", None) is None:
            self.model._flat_params_bf16 = torch.empty(self.p.shape, dtype=torch.bfloat16, device=self.p.device)
        p16 = self.model._flat_params_bf16 if bf16 else None
        ops.adam_step(self.p, g, self.m, self.v, self.lr, self.betas[0], self.betas[1], self.eps, self.t, grad_scale, p_bf16=p16)
        ops.WEIGHT_EPOCH += 1          # the kernel rewrote the weights in place: every cached operand pack is stale
        if bf16:                       # ... except the bf16 shadow of the flat buffer, which the same kernel just refreshed
            self.model._flat_params_bf16_tag = shadow_tag(self.model)


# ------------------------------------------------------------------------------------------------ train-mode regularisers
ENC_SITES = ("attn", "h1", "act", "h3")
DEC_SITES = ("ppe", "dec.sa", "dec.d1", "dec.ca", "dec.d2", "dec.act", "dec.d3")


def regulariser_shapes(B, T, fd, cfg, dec_heads=4):
    """site -> shape of its dropout draw (the dict layout of synth.train_regularisers)."""
    M, C, F, H = B * T, cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads
    shp = {"enc_in": (M, C), "ppe": (M, fd), "dec.sa": (B, dec_heads, T, T), "dec.d1": (M, fd), "dec.ca": (B, dec_heads, T, T),
           "dec.d2": (M, fd), "dec.act": (M, 2 * fd), "dec.d3": (M, fd)}
    if getattr(cfg, "feat_proj_dropout", 0.0) > 0:
        shp["featproj"] = (M, C)
    for l in range(cfg.num_hidden_layers):
        shp.update({f"l{l}.attn": (B, H, T, T), f"l{l}.h1": (M, C), f"l{l}.act": (M, F), f"l{l}.h3": (M, C)})
    return shp


def draw_regularisers(B, T, fd, cfg, device, generator=None, p_dec=0.1, spec_augment=True):
    """One step's draws on the device, in the reference's configuration: HF Wav2Vec2Config's hidden / attention / activation dropout
    and layerdrop, SpecAugment with mask_time_prob / mask_time_length and at least two spans per clip (wav2vec.py:120-131 calls
    _compute_mask_indices(min_masks=2)), dropout 0.1 for the PPE and nn.TransformerDecoderLayer. The random stream is torch's device
    generator, not the reference's (which interleaves its draws with the forward): parity is pinned with INJECTED draws instead."""
    def mask(shape, p):
        if p <= 0:
            return None
        return (torch.rand(shape, device=device, generator=generator) >= p).float().div_(1.0 - p)

    shp = regulariser_shapes(B, T, fd, cfg)
    p_site = site_dropout_p(cfg, p_dec)
    masks = {}
    for name, s_ in shp.items():
        mk = mask(s_, p_site[name])
        if mk is not None:
            masks[name] = mk
    keep = (torch.rand(cfg.num_hidden_layers, generator=generator, device=device) >= cfg.layerdrop).tolist()
    spec = None
    if spec_augment and getattr(cfg, "apply_spec_augment", True) and cfg.mask_time_prob > 0:
        L = cfg.mask_time_length
        n_spans = max(2, int(cfg.mask_time_prob * T / L + torch.rand((), device=device, generator=generator).item()))
        if L < T:
            starts = torch.randint(0, T - L + 1, (B, n_spans), device=device, generator=generator)
            idx = (starts[:, :, None] + torch.arange(L, device=device)[None, None, :]).reshape(B, -1)
            spec = torch.zeros(B, T, dtype=torch.bool, device=device).scatter_(1, idx, True)
    return {"p": p_dec, "spec_mask": spec, "layer_keep": keep, "masks": masks}


def site_dropout_p(cfg, p_dec=0.1):
    """site -> dropout probability in the reference's configuration (HF Wav2Vec2Config; 0.1 for the PPE and the decoder layer)."""
    p_site = {"featproj": getattr(cfg, "feat_proj_dropout", 0.0), "enc_in": cfg.hidden_dropout}
    for l in range(cfg.num_hidden_layers):
        p_site.update({f"l{l}.attn": cfg.attention_dropout, f"l{l}.h1": cfg.hidden_dropout, f"l{l}.act": cfg.activation_dropout,
                       f"l{l}.h3": cfg.hidden_dropout})
    for n in DEC_SITES:
        p_site[n] = p_dec
    return p_site


class DeviceDraws:
    """Every draw of a TRAIN-mode step made by the library on the device (csrc/train_draw.cu): all dropout sites are views of ONE flat
    fp32 buffer filled by one launch per distinct probability (normally one), LayerDrop becomes the 0 / 1 blend rows of the graph-stable
    form, SpecAugment spans are drawn in the same launch, and the step counter lives in device memory - so `draw()` can be captured
    in the step's CUDA graph and every replay sees fresh draws with no host work. Counter-based Philox4x32-10: the draws are a pure
    function of (seed, step, element), reproduced bit for bit by oracle/philox_oracle.py (tests/test_gpu_train.py)."""

    def __init__(self, B, T, fd, cfg, device, seed=0, p_dec=0.1, spec_augment=True):
        self.B, self.T, self.cfg, self.device, self.seed = B, T, cfg, torch.device(device), int(seed)
        shapes = regulariser_shapes(B, T, fd, cfg)
        p_site = site_dropout_p(cfg, p_dec)
        by_p = {}
        for name, shp in shapes.items():
            if p_site[name] > 0:
                by_p.setdefault(float(p_site[name]), []).append((name, shp))
        self.groups, self.layout, off = [], {}, 0          # groups: (p, first element, elements); layout: site -> (group, offset in group)
        for gi, (p, sites) in enumerate(sorted(by_p.items())):
            start = off
            for name, shp in sites:
                n = 1
                for d_ in shp:
                    n *= d_
                self.layout[name] = (gi, off - start, shp)
                off += (n + 3) // 4 * 4
            self.groups.append((p, start, off - start))
        self.flat = torch.empty(max(off, 4), dtype=torch.float32, device=self.device)
        self.masks = {}
        for name, (gi, o, shp) in self.layout.items():
            n = 1
            for d_ in shp:
                n *= d_
            a = self.groups[gi][1] + o
            self.masks[name] = self.flat[a:a + n].view(shp)
        lo, hi = self.seed & 0xFFFFFFFF, (self.seed >> 32) & 0xFFFFFFFF
        as_i32 = lambda v: v - (1 << 32) if v >= (1 << 31) else v  # noqa: E731
        self.state = torch.tensor([as_i32(lo), as_i32(hi), 0, 0], dtype=torch.int32, device=self.device)
        L, self.rows = cfg.num_hidden_layers, B * T * cfg.hidden_size
        self.layerdrop = float(cfg.layerdrop)
        self.blend = torch.ones((2, L, self.rows), dtype=torch.float32, device=self.device)
        self.keep_flags = torch.ones(L, dtype=torch.float32, device=self.device)
        self.spec = None
        if spec_augment and getattr(cfg, "apply_spec_augment", True) and cfg.mask_time_prob > 0:
            self.spec = torch.zeros(B * T, dtype=torch.uint8, device=self.device)
        self.span_len, self.span_rate, self.min_spans = int(cfg.mask_time_length), float(cfg.mask_time_prob * T / cfg.mask_time_length), 2
        self.reg = {"p": p_dec, "masks": self.masks, "spec_mask": self.spec, "layer_keep": [True] * L, "layer_blend": self.blend,
                    "_on_device": str(self.device)}
        if "dec.ca" in self.masks:
            self.reg["ca_diag"] = torch.ones((B, T, self.masks["dec.ca"].shape[1]), dtype=torch.float32, device=self.device)

    def draw(self):
        """Launch the draws of the next step on the current stream (capturable) -> the `reg` dict TrainStep.forward takes."""
        for gi, (p, start, n) in enumerate(self.groups):
            ops.dropout_masks_(self.flat[start:start + n], p, self.state, stream_id=gi)
        ops.layerdrop_spec_draw_(self.blend, self.keep_flags, self.rows, self.layerdrop, self.spec, self.B, self.T, self.span_len,
                                 self.span_rate, self.min_spans, self.state)
        ops.draw_bump_step_(self.state)
        if "ca_diag" in self.reg:
            self.reg["ca_diag"].copy_(torch.diagonal(self.masks["dec.ca"], dim1=2, dim2=3).permute(0, 2, 1))
        return self.reg


def regularisers_to_device(reg, device):
    """The dict of synth.train_regularisers / draw_regularisers with every tensor on `device` (fp32 contiguous masks, uint8 span rows)
    and the cross-attention draw reduced to what the degenerate cross-attention sees: the ONE visible key of query t is key t
    (enc_dec_mask, vocaset), so head h of the value row (b, t) is scaled by mask[b, h, t, t]."""
    if reg is None or reg.get("_on_device") == str(device):
        return reg
    out = {"p": reg.get("p"), "layer_keep": [bool(k) for k in reg["layer_keep"]], "_on_device": str(device)}
    out["masks"] = {k: v.to(device=device, dtype=torch.float32).contiguous() for k, v in reg["masks"].items()}
    sm = reg.get("spec_mask")
    out["spec_mask"] = None if sm is None else sm.to(device).reshape(-1).to(torch.uint8).contiguous()
    ca = out["masks"].get("dec.ca")
    if ca is not None:
        out["ca_diag"] = torch.diagonal(ca, dim1=2, dim2=3).permute(0, 2, 1).contiguous()          # [B, T, heads]
    if reg.get("layer_blend") is not None:
        out["layer_blend"] = reg["layer_blend"].to(device)
    return out


# ------------------------------------------------------------------------------------------------ the step
class _Lin:
    """Operand plumbing of one precision."""

    def __init__(self, bf16, flat32=None, flat16=None):
        self.bf16 = bf16
        self.dt = torch.bfloat16 if bf16 else torch.float32
        self.flat32, self.flat16 = flat32, flat16

    def w(self, W32):
        """GEMM operand of a weight: for a dense view of the flat parameter buffer, the same view of its bf16 shadow (no launch)."""
        if not self.bf16:
            return W32
        if self.flat16 is not None and W32.is_contiguous():
            off = (W32.data_ptr() - self.flat32.data_ptr()) // 4
            if 0 <= off and off + W32.numel() <= self.flat32.numel():
                return self.flat16[off:off + W32.numel()].view(W32.shape)
        return ops.cast_bf16(W32)

    def a(self, x32):                     # [M, K] fp32 contiguous -> GEMM operand
        return ops.cast_bf16(x32) if self.bf16 else x32

    def t(self, x32, R_pad=None):         # [R, C] fp32 -> operand [C, R_pad]
        return ops.transpose_cast(x32, self.dt, R_pad)

    def fwd(self, x_op, W32, b, residual=None, act=0):
        return ops.linear(x_op, self.w(W32), b, residual=residual, act=act, out_dtype=torch.float32)

    def bwd(self, dy32, x32, W32, gW, gb, Mp, want_dx=True, residual=None, dy_op=None, xT=None):
        """y = x W^T + b. gW [N, K] <- dy^T x ; gb [N] <- colsum(dy) ; returns dx = dy W (+ residual) or None."""
        M, N = dy32.shape
        K = W32.shape[1]
        dyT = self.t(dy32, Mp)
        xT = self.t(x32, Mp) if xT is None else xT
        ops.gemm(dyT, xT, None, gW, rows=N, N=K, K=Mp, a_rows_alloc=N, c_ld=gW.stride(0))
        if gb is not None:
            ops.colsum(dy32, out=gb)
        if not want_dx:
            return None
        dy_op = self.a(dy32) if dy_op is None else dy_op
        WT = self.t(W32, N)                # [K, N]
        return ops.linear(dy_op, WT, None, residual=residual, out_dtype=torch.float32)


class TrainStep:
    """forward -> loss, backward -> flat gradient, for a FaceformerVert drop-in. `precision` follows the model."""

    def __init__(self, model, buckets: GradBuckets | None = None):
        if getattr(model, "variant", None) != "vert":
            raise NotImplementedError("the training step follows models/faceformer_vert.py (audio-only hidden states)")
        self.model = model
        if getattr(model, "_flat_params", None) is None:
            flatten_parameters(model)
        self.layout = model._flat_layout
        self.buckets = buckets
        self.saved = None

    # ---------------------------------------------------------------------------- positional conv (grouped, weight-normed)
    def _posconv(self, x32, w, B, T, flip, bias=None):
        """Grouped Conv1d(k=128, pad=64, groups=16)[..., :-1] over time as 4 block-diagonal conv-mode GEMMs (4 groups = 192 channels
        each). flip=False: forward, out[t] = sum_j x[t + j - 64] w[:, :, j]. flip=True: the input gradient,
        dx[r] = sum_j dpc[r + 64 - j] w[:, :, j]^T (a correlation with the tap-flipped, channel-transposed weight)."""
        cfg = self.model.audio_encoder.config
        k, g, Cc = cfg.num_conv_pos_embeddings, cfg.num_conv_pos_embedding_groups, cfg.hidden_size
        cg = Cc // g
        if cg * 4 % 64 != 0 or g % 4 != 0:
            raise NotImplementedError("positional conv layout other than 16 groups x 48 channels")
        front = k - 1 - k // 2 if flip else k // 2
        Tp = T + k
        xpad = torch.zeros((B, Tp, Cc), dtype=torch.float32, device=x32.device)
        xpad[:, front:front + T] = x32.view(B, T, Cc)
        wg = w.reshape(g, cg, cg, k)                                   # [group, co, ci, tap]
        out = torch.empty((B * T, Cc), dtype=torch.float32, device=x32.device)
        lin = self.lin
        W4 = 4 * cg
        for q in range(g // 4):
            blk = torch.zeros(4, cg, k, 4, cg, dtype=torch.float32, device=w.device)
            for gl in range(4):
                blk[gl, :, :, gl, :] = wg[4 * q + gl].flip(-1).permute(1, 2, 0) if flip else wg[4 * q + gl].permute(0, 2, 1)
            wq = lin.a(blk.reshape(W4, k * W4).contiguous())
            xs = lin.a(xpad[:, :, W4 * q:W4 * (q + 1)].contiguous())
            ops.gemm(xs, wq, None if bias is None else bias[W4 * q:W4 * (q + 1)], out[:, W4 * q:], batch=B, rows=T, N=W4, K=k * W4, conv_taps=k, conv_stride=1, a_ld=W4,
                     a_batch_stride=Tp * W4, a_rows_alloc=Tp, c_ld=Cc, c_batch_stride=T * Cc,
                     algorithmic_flops=2.0 * B * T * W4 * k * cg)
        return out

    def _posconv_dw(self, x32, dpc32, B, T):
        """dW [C, C/groups, k] of the grouped positional conv on the tensor cores: per block of 4 groups (192 channels),
        dW_blk[co, (j, ci)] = sum_{b,t} dpc[b,t,co] * xpad[b,t+j,ci] is one GEMM (rows = 192 output channels, N = k*192, contraction
        = clips x frames padded to 64 per clip) over the transposed unfold of x; the 4 diagonal 48x48 blocks per tap are the groups'
        gradients (the off-diagonal ones are cross-group products that the grouped conv does not have)."""
        cfg = self.model.audio_encoder.config
        k, g, Cc = cfg.num_conv_pos_embeddings, cfg.num_conv_pos_embedding_groups, cfg.hidden_size
        cg = Cc // g
        W4 = 4 * cg
        Tq = _pad(T)
        lin = self.lin
        dw = torch.empty((Cc, cg, k), dtype=torch.float32, device=x32.device)
        dpc_pad = torch.zeros((B, Tq, Cc), dtype=torch.float32, device=x32.device)
        dpc_pad[:, :T] = dpc32.view(B, T, Cc)
        dpc_pad = dpc_pad.view(B * Tq, Cc)
        for q in range(g // 4):
            xu = ops.posconv_unfold_t(x32, B, T, Tq, W4 * q, W4, k, lin.dt)                      # [k*192, B*Tq]
            dT = lin.t(dpc_pad[:, W4 * q:W4 * (q + 1)], B * Tq)                                    # [192, B*Tq]
            blk = torch.empty((W4, k * W4), dtype=torch.float32, device=x32.device)
            ops.gemm(dT, xu, None, blk, rows=W4, N=k * W4, K=B * Tq, a_rows_alloc=W4)
            blk = blk.view(4, cg, k, 4, cg)                                                        # [gl_out, co, tap, gl_in, ci]
            for gl in range(4):
                dw[(4 * q + gl) * cg:(4 * q + gl + 1) * cg] = blk[gl, :, :, gl, :].permute(0, 2, 1)  # [co, ci, tap]
        return dw

    # ---------------------------------------------------------------------------- forward
    @torch.no_grad()
    def forward(self, audio, gt_verts, reg=None):
        """audio [B, N] fp32, gt_verts [B, T, V*3] fp32 (row stride may be padded) -> loss (0-dim fp32 tensor).
        reg: the draws of one TRAIN-mode step (module docstring); None = the .eval() arithmetic."""
        m = self.model
        w2v, cfg = m.audio_encoder, m.audio_encoder.config
        if not audio.is_cuda:
            raise RuntimeError("avi_talking_b200 training runs on CUDA only (no CPU fallback)")
        bf16 = m.precision == "bf16"
        fd = m.args.feature_dim
        if bf16 and fd % 64 != 0:
            raise NotImplementedError("bf16 training needs feature_dim % 64 == 0 (use precision='fp32')")
        self.lin = lin = _Lin(bf16, m._flat_params, bf16_shadow(m) if bf16 else None)
        B, T, V3 = gt_verts.shape
        M = B * T
        if T > 128:
            raise NotImplementedError("training clips of more than 128 frames (VOCASET clips are ~100-150 at 30 fps; split longer ones)")
        S = {"B": B, "T": T, "M": M}
        S["reg"] = reg = regularisers_to_device(reg, audio.device)
        mk = reg["masks"] if reg is not None else {}
        keep = reg["layer_keep"] if reg is not None else [True] * len(w2v.encoder.layers)
        # graph-stable LayerDrop (GraphedTrainStep): every layer runs and its output is blended with its input by 0 / 1 row masks
        # ([n_layers, M * C] keep and 1 - keep), so the launch sequence does not depend on the draw
        blend = reg.get("layer_blend") if reg is not None else None

        def drop(name, t, residual=None):
            """nn.Dropout at site `name` (+ the residual the reference adds right after); the plain add / identity without a draw."""
            if name in mk:
                return ops.mask_mul(t, mk[name], residual)
            return t if residual is None else ops.add_f32(t, residual)
        eps = cfg.layer_norm_eps
        H, D = cfg.num_attention_heads, cfg.hidden_size // cfg.num_attention_heads
        # -- frozen conv feature extractor (no gradient, :154) and the 50 -> frame_num resample (wav2vec.py:97-108)
        w2v.precision = m.precision
        feats, T50, La = w2v._feature_extractor(audio.contiguous().float(), w2v._pack_extractor())
        Cf = cfg.conv_dim[-1]
        S["h"] = h = ops.w2v_lerp(feats, La * Cf, B, T50, T, Cf)
        fp = w2v.feature_projection
        hn, _ = ops.layernorm(h, fp.layer_norm.weight, fp.layer_norm.bias, eps=eps)
        S["hn"] = hn
        proj = lin.fwd(lin.a(hn), fp.projection.weight, fp.projection.bias)                         # wav2vec.py:120
        if "featproj" in mk:                                                                        # Wav2Vec2FeatureProjection.dropout
            proj = ops.mask_mul(proj, mk["featproj"])
        if reg is not None and reg["spec_mask"] is not None:                                        # SpecAugment, wav2vec.py:120-131
            ops.spec_augment_fwd_(proj, reg["spec_mask"], w2v.masked_spec_embed)
        S["proj"] = proj
        # -- positional conv embedding + encoder LayerNorm
        S["pos_w"] = pos_w = w2v._posconv_weight().float().contiguous()
        S["pc"] = pcb = self._posconv(proj, pos_w, B, T, flip=False, bias=w2v.encoder.pos_conv_embed.conv.bias)
        gpc, _ = ops.act_fwd(pcb, ACT_GELU)
        S["h0pre"] = h0pre = ops.add_f32(proj, gpc)
        x, _ = ops.layernorm(h0pre, w2v.encoder.layer_norm.weight, w2v.encoder.layer_norm.bias, eps=eps)
        if "enc_in" in mk:                                                                          # Wav2Vec2Encoder.dropout
            x = ops.mask_mul(x, mk["enc_in"])
        # -- 12 post-LN encoder layers (fused q|k|v weight = a view of the flat parameter buffer)
        P, lay = m._flat_params, self.layout
        S["layers"] = []
        for l, lyr in enumerate(w2v.encoder.layers):
            if blend is None and not keep[l]:                  # LayerDrop: the layer is the identity and gets no gradient
                S["layers"].append(None)
                continue
            x_in = x
            pre = f"audio_encoder.encoder.layers.{l}.attention."
            Wqkv = lay.span(P, pre + "q_proj.weight", 3 * cfg.hidden_size, cfg.hidden_size)
            bqkv = lay.span(P, pre + "q_proj.bias", 1, 3 * cfg.hidden_size).view(-1)
            a = lyr.attention
            L = {"x": x, "Wqkv": Wqkv}
            L["qkv"] = qkv = lin.fwd(lin.a(x), Wqkv, bqkv)
            L["att"], L["P"] = att, _ = ops.attn_train_fwd(qkv, B, T, H, D, pmask=mk.get(f"l{l}.attn"))
            if f"l{l}.h1" in mk:
                y = drop(f"l{l}.h1", lin.fwd(lin.a(att), a.out_proj.weight, a.out_proj.bias), x)
            else:
                y = lin.fwd(lin.a(att), a.out_proj.weight, a.out_proj.bias, residual=x)
            L["y"] = y
            L["h1"] = h1 = ops.layernorm(y, lyr.layer_norm.weight, lyr.layer_norm.bias, eps=eps)[0]
            ff = lyr.feed_forward
            L["fpre"] = fpre = lin.fwd(lin.a(h1), ff.intermediate_dense.weight, ff.intermediate_dense.bias)
            L["f"] = f = drop(f"l{l}.act", ops.act_fwd(fpre, ACT_GELU)[0])          # the operand of output_dense (dropped)
            if f"l{l}.h3" in mk:
                y2 = drop(f"l{l}.h3", lin.fwd(lin.a(f), ff.output_dense.weight, ff.output_dense.bias), h1)
            else:
                y2 = lin.fwd(lin.a(f), ff.output_dense.weight, ff.output_dense.bias, residual=h1)
            L["y2"] = y2
            x = ops.layernorm(y2, lyr.final_layer_norm.weight, lyr.final_layer_norm.bias, eps=eps)[0]
            if blend is not None:                              # x <- keep * layer(x_in) + (1 - keep) * x_in
                x = ops.mask_mul(x_in, blend[1][l], residual=ops.mask_mul(x, blend[0][l]))
            S["layers"].append(L)
        S["h12"] = x
        # -- heads and the teacher-forced decoder layer (faceformer_vert.py:369,437-454)
        S["mem"] = mem = lin.fwd(lin.a(x), m.audio_feature_map.weight, m.audio_feature_map.bias)
        V3p = _pad(V3)
        if getattr(self, "_template", None) is None or self._template_src is not m.template:
            self._template_src = m.template                   # cached on the device (no H2D inside a graph capture)
            self._template = m.template.to(audio.device).reshape(-1).float().contiguous()
        template = self._template
        S["vin"] = vin = ops.tf_input_rows(gt_verts, template, V3p)                                    # :443-444
        Wvm_p = ops.cast_pad2d(m.vertice_map.weight, lin.dt, C_pad=V3p)
        xd = ops.linear(lin.a(vin), Wvm_p, m.vertice_map.bias, out_dtype=torch.float32)               # :445
        style = m.obj_vector.weight[:, 0].contiguous().view(1, fd)                                     # one_hot[:, 0] = 1 (:361-365)
        period = m.args.period
        ops.ff_add_style_pe(xd, style, m.PPE.pe[0, :period].contiguous(), B, T, fd, period)            # :446-447
        xd = drop("ppe", xd)                                                                           # PPE dropout
        dl = m.transformer_decoder.layers[0]
        S["xd"] = xd
        S["dqkv_in"] = qkv = lin.fwd(lin.a(xd), dl.self_attn.in_proj_weight, dl.self_attn.in_proj_bias)
        S["datt"], S["dP"] = att, _ = ops.attn_train_fwd(qkv, B, T, 4, fd // 4, bias_mode=1, period=period, pmask=mk.get("dec.sa"))
        if "dec.d1" in mk:
            y1 = drop("dec.d1", lin.fwd(lin.a(att), dl.self_attn.out_proj.weight, dl.self_attn.out_proj.bias), xd)
        else:
            y1 = lin.fwd(lin.a(att), dl.self_attn.out_proj.weight, dl.self_attn.out_proj.bias, residual=xd)
        S["dy1"] = y1
        S["x1"] = x1 = ops.layernorm(y1, dl.norm1.weight, dl.norm1.bias, eps=1e-5)[0]
        # cross-attention with enc_dec_mask (:80-88) sees exactly one key per query: softmax == 1, output = out_proj(v_proj(mem_t)),
        # and the q / k projections get exactly zero gradient
        Wc, bc = dl.multihead_attn.in_proj_weight, dl.multihead_attn.in_proj_bias
        cv = lin.fwd(lin.a(mem), Wc[2 * fd:], bc[2 * fd:])
        if reg is not None and "ca_diag" in reg:           # dropout on the single unit probability: head h of row (b, t) x mask[b, h, t, t]
            S["ca_scale"] = reg["ca_diag"].repeat_interleave(fd // 4, dim=2).reshape(M, fd).contiguous()
            cv = ops.mask_mul(cv, S["ca_scale"])
        S["cv"] = cv
        if "dec.d2" in mk:
            y2 = drop("dec.d2", lin.fwd(lin.a(cv), dl.multihead_attn.out_proj.weight, dl.multihead_attn.out_proj.bias), x1)
        else:
            y2 = lin.fwd(lin.a(cv), dl.multihead_attn.out_proj.weight, dl.multihead_attn.out_proj.bias, residual=x1)
        S["dy2"] = y2
        S["x2"] = x2 = ops.layernorm(y2, dl.norm2.weight, dl.norm2.bias, eps=1e-5)[0]
        S["f1pre"] = f1pre = lin.fwd(lin.a(x2), dl.linear1.weight, dl.linear1.bias)
        S["f1"] = f1 = drop("dec.act", ops.act_fwd(f1pre, ACT_RELU)[0])
        if "dec.d3" in mk:
            y3 = drop("dec.d3", lin.fwd(lin.a(f1), dl.linear2.weight, dl.linear2.bias), x2)
        else:
            y3 = lin.fwd(lin.a(f1), dl.linear2.weight, dl.linear2.bias, residual=x2)
        S["dy3"] = y3
        S["x3"] = x3 = ops.layernorm(y3, dl.norm3.weight, dl.norm3.bias, eps=1e-5)[0]
        out = ops.empty_rows(M, V3, audio.device)
        bias = (m.vertice_map_r.bias + template).contiguous()                                          # + template (:475)
        ops.gemm(lin.a(x3), lin.w(m.vertice_map_r.weight), bias, out, rows=M, N=V3, K=fd, a_rows_alloc=M, c_ld=out.stride(0))
        loss64, dout = ops.mse_loss_grad(out, gt_verts.reshape(M, V3), 10.0)                           # :481-482
        S["dout"] = dout
        self.saved = S
        return loss64.float().reshape(())

    # ---------------------------------------------------------------------------- backward
    @torch.no_grad()
    def backward(self, grad_scale=1.0, grad_tensor=None):
        """Gradient of `grad_scale * loss` (times the 0-dim device tensor `grad_tensor`, if given: no host sync) into a fresh flat
        buffer; sets model._flat_grad and every parameter's .grad view. Under data parallel the upstream gradient is pre-divided by
        the world size, so after the SUM all-reduce `.grad` holds the rank-AVERAGED gradient (DistributedDataParallel semantics):
        any optimizer or gradient clipping may consume it as is. Returns 1.0 (kept for callers that forward it to FlatAdam.step)."""
        S, m, lin, lay = self.saved, self.model, self.lin, self.layout
        if S is None:
            raise RuntimeError("TrainStep.backward without a forward")
        w2v, cfg = m.audio_encoder, m.audio_encoder.config
        B, T, M = S["B"], S["T"], S["M"]
        Mp = _pad(M)
        fd = m.args.feature_dim
        eps = cfg.layer_norm_eps
        G = torch.zeros(lay.total, dtype=torch.float32, device=S["h"].device)
        g = lambda n: lay.view(G, n)  # noqa: E731
        reg = S["reg"]
        mk = reg["masks"] if reg is not None else {}
        blend = reg.get("layer_blend") if reg is not None else None

        def dropb(name, dy):
            """Gradient through the dropout of site `name`: dy * mask (the same launch as the forward); identity without a draw."""
            return ops.mask_mul(dy, mk[name]) if name in mk else dy
        bk = self.buckets
        if bk is not None:
            bk.begin()
        dout = S["dout"]
        if bk is not None:
            grad_scale = grad_scale * bk.prescale
        if grad_scale != 1.0:
            dout = dout * grad_scale
        if grad_tensor is not None:
            dout = dout * grad_tensor.to(dout.dtype)
        dl = m.transformer_decoder.layers[0]
        d = "transformer_decoder.layers.0."
        V3 = dout.shape[1]
        V3p = _pad(V3)
        # vertice_map_r: out = x3 Wr^T + br
        ops.gemm(lin.t(dout, Mp), lin.t(S["x3"], Mp), None, g("vertice_map_r.weight"), rows=V3, N=fd, K=Mp, a_rows_alloc=V3)
        ops.colsum(dout, out=g("vertice_map_r.bias"))
        WrT = ops.transpose_cast(m.vertice_map_r.weight, lin.dt, V3p)                                   # [fd, 15104]
        dx3 = ops.linear(ops.cast_pad2d(dout, lin.dt, C_pad=V3p), WrT, None, out_dtype=torch.float32)
        dy3 = ops.layernorm_bwd(S["dy3"], dl.norm3.weight, dx3, g(d + "norm3.weight"), g(d + "norm3.bias"), eps=1e-5)
        df1 = lin.bwd(dropb("dec.d3", dy3), S["f1"], dl.linear2.weight, g(d + "linear2.weight"), g(d + "linear2.bias"), Mp)
        df1pre = ops.act_bwd(S["f1pre"], dropb("dec.act", df1), ACT_RELU)
        dx2 = lin.bwd(df1pre, S["x2"], dl.linear1.weight, g(d + "linear1.weight"), g(d + "linear1.bias"), Mp, residual=dy3)
        dy2 = ops.layernorm_bwd(S["dy2"], dl.norm2.weight, dx2, g(d + "norm2.weight"), g(d + "norm2.bias"), eps=1e-5)
        dcv = lin.bwd(dropb("dec.d2", dy2), S["cv"], dl.multihead_attn.out_proj.weight, g(d + "multihead_attn.out_proj.weight"),
                      g(d + "multihead_attn.out_proj.bias"), Mp)
        if "ca_scale" in S:
            dcv = ops.mask_mul(dcv, S["ca_scale"])
        Wc = dl.multihead_attn.in_proj_weight
        dmem = lin.bwd(dcv, S["mem"], Wc[2 * fd:], g(d + "multihead_attn.in_proj_weight")[2 * fd:],
                       g(d + "multihead_attn.in_proj_bias")[2 * fd:], Mp)
        dy1 = ops.layernorm_bwd(S["dy1"], dl.norm1.weight, dy2, g(d + "norm1.weight"), g(d + "norm1.bias"), eps=1e-5)
        datt = lin.bwd(dropb("dec.d1", dy1), S["datt"], dl.self_attn.out_proj.weight, g(d + "self_attn.out_proj.weight"),
                       g(d + "self_attn.out_proj.bias"), Mp)
        dqkv = ops.attn_train_bwd(S["dqkv_in"], S["dP"], datt, B, T, 4, fd // 4, pmask=mk.get("dec.sa"))
        dxd = lin.bwd(dqkv, S["xd"], dl.self_attn.in_proj_weight, g(d + "self_attn.in_proj_weight"), g(d + "self_attn.in_proj_bias"),
                      Mp, residual=dy1)
        dxd = dropb("ppe", dxd)
        # xd = vin Wvm^T + bvm + style + pe
        vinT = lin.t(S["vin"][:, :V3], Mp)                                                              # [15069, Mp]
        ops.gemm(lin.t(dxd, Mp), vinT, None, g("vertice_map.weight"), rows=fd, N=V3, K=Mp, a_rows_alloc=fd, c_ld=V3)
        ops.colsum(dxd, out=g("vertice_map.bias"))
        gobj = torch.empty(fd, dtype=torch.float32, device=G.device)
        ops.colsum(dxd, out=gobj)
        g("obj_vector.weight")[:, 0] = gobj                                                              # other subjects: one_hot = 0
        # audio_feature_map
        dh = lin.bwd(dmem, S["h12"], m.audio_feature_map.weight, g("audio_feature_map.weight"), g("audio_feature_map.bias"), Mp)
        if bk is not None:
            bk.ready(G, lay.offsets[lay.names[lay.segments[0]]])
        H, D = cfg.num_attention_heads, cfg.hidden_size // cfg.num_attention_heads
        C = cfg.hidden_size
        nl = len(w2v.encoder.layers)
        for l in reversed(range(nl)):
            lyr, L = w2v.encoder.layers[l], S["layers"][l]
            if L is None:                                       # LayerDrop: identity, gradients stay zero
                if bk is not None:
                    seg = lay.segments[nl - l]
                    bk.ready(G, lay.offsets[lay.names[seg]] if seg < len(lay.names) else lay.total)
                continue
            p = f"audio_encoder.encoder.layers.{l}."
            ff, a = lyr.feed_forward, lyr.attention
            dh_up = dh
            if blend is not None:                               # a dropped layer sees a zero upstream gradient: all its gradients are 0
                dh = ops.mask_mul(dh_up, blend[0][l])
            dy2 = ops.layernorm_bwd(L["y2"], lyr.final_layer_norm.weight, dh, g(p + "final_layer_norm.weight"),
                                    g(p + "final_layer_norm.bias"), eps=eps)
            df = lin.bwd(dropb(f"l{l}.h3", dy2), L["f"], ff.output_dense.weight, g(p + "feed_forward.output_dense.weight"),
                         g(p + "feed_forward.output_dense.bias"), Mp)
            dfpre = ops.act_bwd(L["fpre"], dropb(f"l{l}.act", df), ACT_GELU)
            dh1 = lin.bwd(dfpre, L["h1"], ff.intermediate_dense.weight, g(p + "feed_forward.intermediate_dense.weight"),
                          g(p + "feed_forward.intermediate_dense.bias"), Mp, residual=dy2)
            dy = ops.layernorm_bwd(L["y"], lyr.layer_norm.weight, dh1, g(p + "layer_norm.weight"), g(p + "layer_norm.bias"), eps=eps)
            datt = lin.bwd(dropb(f"l{l}.h1", dy), L["att"], a.out_proj.weight, g(p + "attention.out_proj.weight"),
                           g(p + "attention.out_proj.bias"), Mp)
            dqkv = ops.attn_train_bwd(L["qkv"], L["P"], datt, B, T, H, D, pmask=mk.get(f"l{l}.attn"))
            gW = lay.span(G, p + "attention.q_proj.weight", 3 * C, C)
            gb = lay.span(G, p + "attention.q_proj.bias", 1, 3 * C).view(-1)
            dh = lin.bwd(dqkv, L["x"], L["Wqkv"], gW, gb, Mp, residual=dy)
            if blend is not None:                               # + the identity path of a dropped layer
                dh = ops.mask_mul(dh_up, blend[1][l], residual=dh)
            if bk is not None:
                seg = lay.segments[nl - l]
                bk.ready(G, lay.offsets[lay.names[seg]] if seg < len(lay.names) else lay.total)
        # encoder LayerNorm, positional conv (weight norm), feature projection
        dh0pre = ops.layernorm_bwd(S["h0pre"], w2v.encoder.layer_norm.weight, dropb("enc_in", dh), g("audio_encoder.encoder.layer_norm.weight"),
                                   g("audio_encoder.encoder.layer_norm.bias"), eps=eps)
        dpc = ops.act_bwd(S["pc"], dh0pre, ACT_GELU)
        ops.colsum(dpc, out=g("audio_encoder.encoder.pos_conv_embed.conv.bias"))
        k, groups = cfg.num_conv_pos_embeddings, cfg.num_conv_pos_embedding_groups
        dw = self._posconv_dw(S["proj"], dpc, B, T)
        named = dict(m.named_parameters())
        dv, dg = ops.weightnorm_bwd(named[lay.pos_v], named[lay.pos_g], dw)
        g(lay.pos_v).copy_(dv)
        g(lay.pos_g).copy_(dg)
        dproj = ops.add_f32(dh0pre, self._posconv(dpc, S["pos_w"], B, T, flip=True))
        if reg is not None and reg["spec_mask"] is not None:    # the replaced rows feed masked_spec_embed, not the projection
            if "audio_encoder.masked_spec_embed" not in lay.offsets:
                raise RuntimeError("SpecAugment needs audio_encoder.masked_spec_embed among the trainable parameters")
            ops.spec_augment_bwd_(dproj, reg["spec_mask"], g("audio_encoder.masked_spec_embed"))
        dproj = dropb("featproj", dproj)
        fp = w2v.feature_projection
        fpn = "audio_encoder.feature_projection."
        dhn = lin.bwd(dproj, S["hn"], fp.projection.weight, g(fpn + "projection.weight"), g(fpn + "projection.bias"), Mp)
        ops.layernorm_bwd(S["h"], fp.layer_norm.weight, dhn, g(fpn + "layer_norm.weight"), g(fpn + "layer_norm.bias"), want_dx=False,
                          eps=eps)
        scale = bk.finish(G) if bk is not None else 1.0
        m._flat_grad = G
        for n in lay.names:
            named[n].grad = lay.view(G, n)
        self.saved = None
        return scale


class _LossFn(torch.autograd.Function):
    """`loss = model(...)` / `loss.backward()` as upstream: backward runs TrainStep.backward and fills the parameters' .grad."""

    @staticmethod
    def forward(ctx, anchor, step, audio, gt, reg):
        ctx.step = step
        return step.forward(audio, gt, reg=reg)

    @staticmethod
    def backward(ctx, gloss):
        step = ctx.step
        # the upstream gradient stays on the device (no float(gloss) host sync every step)
        step.grad_divisor = step.backward(grad_tensor=gloss if gloss.numel() == 1 else None)
        return None, None, None, None, None


def training_loss(step: TrainStep, audio, gt_verts, reg=None):
    """Differentiable-looking loss: a 0-dim tensor whose .backward() fills every trainable parameter's .grad (flat views).
    reg: the draws of a TRAIN-mode step (None = the .eval() arithmetic)."""
    anchor = torch.zeros((), device=audio.device, requires_grad=True)
    return _LossFn.apply(anchor, step, audio, gt_verts, reg)


class GraphedTrainStep:
    """forward + backward of one fixed-shape step captured ONCE in a CUDA graph (the batch-1 step is ~700 short launches: replaying
    the graph removes the per-launch host cost). Adam stays outside the graph (its bias corrections depend on the step count):

        gstep = GraphedTrainStep(model, audio.shape, gt.shape); opt = FlatAdam(model)
        loss = gstep(audio, gt); opt.step()

    Data parallel (`buckets` = a GradBuckets over the process group): the bucketed NCCL all-reduces are captured INSIDE the graph, on
    the communication stream forked from the backward stream (NCCL collectives are stream-ordered kernels, so the fork / join is
    ordinary graph structure); one replay then runs forward, backward and the overlapped exchange with no host launch in between.
    Round 1 ran the multi-GPU step eagerly and two GPUs were slower per step than one graph-replayed GPU.

    Train mode (`reg`, see draw_regularisers): the dropout / SpecAugment draws are copied into static buffers the graph reads, so a
    new draw every step replays the SAME graph. LayerDrop would change which kernels run; inside the graph every layer runs and its
    output is blended with its input by 0 / 1 masks (identical loss, exactly zero gradients for a dropped layer, no re-capture; the
    eager TrainStep skips dropped layers instead). One graph per set of active sites (`max_graphs` most recently used are kept)."""

    def __init__(self, model, audio_shape, gt_shape, warmup=2, buckets: GradBuckets | None = None, max_graphs=8):
        self.model = model
        self.step = TrainStep(model, buckets=buckets)
        dev = model._flat_params.device
        self.audio = torch.zeros(audio_shape, dtype=torch.float32, device=dev)
        self.gt = torch.zeros(gt_shape, dtype=torch.float32, device=dev)
        self.graphs = {}               # key -> dict(graph, loss, grad, reg)
        self.max_graphs = max_graphs
        self.warmup = warmup
        self.graph = None              # the most recently replayed graph (kept for callers that look at it)

    @staticmethod
    def _reg_key(reg):
        if reg is None:
            return None
        if isinstance(reg, DeviceDraws):
            return ("device draws", id(reg))
        return (reg.get("spec_mask") is not None, tuple(sorted(reg["masks"])))

    def _static_reg(self, reg):
        """Device-resident copy of the draws with fixed addresses (what the captured kernels read)."""
        dev = self.audio.device
        st = regularisers_to_device({k: v for k, v in reg.items() if k != "_on_device"}, dev)
        st["masks"] = {k: v.clone() for k, v in st["masks"].items()}
        if st["spec_mask"] is not None:
            st["spec_mask"] = st["spec_mask"].clone()
        if "ca_diag" in st:
            st["ca_diag"] = st["ca_diag"].clone()
        cfg = self.model.audio_encoder.config
        rows = self.gt.shape[0] * self.gt.shape[1] * cfg.hidden_size
        st["layer_blend"] = torch.ones((2, cfg.num_hidden_layers, rows), dtype=torch.float32, device=dev)
        self._load_keep(st, reg["layer_keep"])
        return st

    @staticmethod
    def _load_keep(static, keep):
        k = torch.tensor([1.0 if x else 0.0 for x in keep], dtype=torch.float32).to(static["layer_blend"].device, non_blocking=True)
        static["layer_blend"][0].copy_(k[:, None].expand_as(static["layer_blend"][0]))
        static["layer_blend"][1].copy_((1.0 - k)[:, None].expand_as(static["layer_blend"][1]))
        static["layer_keep"] = [bool(x) for x in keep]

    @staticmethod
    def _load_reg(static, reg):
        for k, v in static["masks"].items():
            v.copy_(reg["masks"][k].reshape(v.shape), non_blocking=True)
        if static["spec_mask"] is not None:
            static["spec_mask"].copy_(reg["spec_mask"].reshape(-1), non_blocking=True)
        if "ca_diag" in static:
            static["ca_diag"].copy_(torch.diagonal(static["masks"]["dec.ca"], dim1=2, dim2=3).permute(0, 2, 1))
        GraphedTrainStep._load_keep(static, reg["layer_keep"])

    def _capture(self, reg):
        m = self.model
        if m.precision == "bf16":
            bf16_shadow(m)
        on_device = isinstance(reg, DeviceDraws)               # the draw launches are part of the graph: fresh draws every replay
        static = None if reg is None or on_device else self._static_reg(reg)
        side = torch.cuda.Stream(device=self.audio.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):                       # one-time attribute calls, allocator warm-up
                self.step.forward(self.audio, self.gt, reg=reg.draw() if on_device else static)
                self.step.backward()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        # thread_local: the NCCL watchdog thread of a process group may query events while this thread captures
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            loss = self.step.forward(self.audio, self.gt, reg=reg.draw() if on_device else static)
            self.step.backward()
        # "draws" keeps a DeviceDraws alive as long as its graph exists: the cache key holds id(reg), which must not be reused
        return {"graph": graph, "loss": loss, "grad": m._flat_grad, "reg": static, "draws": reg if on_device else None}

    def __call__(self, audio, gt_verts, reg=None):
        m = self.model
        self.audio.copy_(audio)
        self.gt.copy_(gt_verts)
        key = (m.precision, m._flat_params.data_ptr(), self._reg_key(reg))
        ent = self.graphs.pop(key, None)
        if ent is None:
            while len(self.graphs) >= self.max_graphs:         # least recently used first (dicts keep insertion order)
                self.graphs.pop(next(iter(self.graphs)))
            ent = self._capture(reg)
        self.graphs[key] = ent                                 # (re-)inserted last = most recently used
        if ent["reg"] is not None:
            self._load_reg(ent["reg"], reg)
        if m.precision == "bf16":
            bf16_shadow(m)                                     # refreshed in place by FlatAdam; re-cast here only if someone else wrote
        self.graph = ent["graph"]
        ent["graph"].replay()
        self.loss, self.grad = ent["loss"], ent["grad"]
        m._flat_grad = self.grad
        lay = m._flat_layout
        for n, p in m.named_parameters():
            if n in lay.offsets:
                p.grad = lay.view(self.grad, n)
        return self.loss
