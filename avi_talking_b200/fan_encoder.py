"""Drop-in for the FanEncoder image branch: third_party/pd_fgc_inference/lib/models/networks/encoder.py:89-126
(FanEncoder) over FAN_feature_extractor.py:13-163 (ConvBlock, HourGlass, FAN_use). Same module tree and `state_dict` keys as the
reference (the nn.Modules below are parameter containers; their computation runs in libavi_b200.so), same
`forward(x [N,3,224,224]) -> (headpose_emb [N,6], eye_embed [N,6], emo_embed [N,30], mouth_feat [N,512])`, eval mode (BatchNorm with
running statistics: what `Faceformer.predict` uses under no_grad).

Data layout: activations are NHWC fp32 rows in a PADDED-WIDTH layout [N*H*(W+2), C] (the last two pixels of every image line are
don't-care). Every 3x3 convolution is implicit: `avi_pad_act` writes the activated input once, with a zero one-pixel border and the same
line pitch (it applies the PRE-activation BatchNorm + ReLU of the ConvBlock, FAN_feature_extractor.py:38-48: conv(relu(bn(x))), and pads
after the activation as F.conv2d does), so that output pixel r and tap (ky, kx) read operand row r + ky*(W+2) + kx - ONE conv-mode GEMM with
2-D taps (AviGemmArgs.conv_taps_x / conv_row_pitch) per convolution. No 9x im2col.
  * the 7x7 stride-2 stem, the 3x3 stride-2 tail conv and 32-channel layers in bf16 mode (k-block of 64) use `avi_im2col_affine` + GEMM;
  * the three convolutions of a ConvBlock write their raw outputs straight into channel slices of the concatenated tensor (c_ld = C_out),
    the residual (identity or the bn-relu-1x1 `downsample`) is added afterwards;
  * max-pool, bilinear upsample + add (hourglass skip connections) and per-channel affine (+ReLU) are small row kernels.
precision "tf32" (default): the tcgen05 GEMM on fp32 operands rounded to TF32 (fp32 accumulate); "fp32": CUDA-core GEMMs; "bf16":
tcgen05 GEMMs on bf16 operands.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import ops
from .ops import ACT_NONE, ACT_RELU


def conv3x3(cin, cout):
    return nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1, bias=False)


class ConvBlock(nn.Module):
    def __init__(self, in_planes, out_planes):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(in_planes)
        self.conv1 = conv3x3(in_planes, out_planes // 2)
        self.bn2 = nn.BatchNorm2d(out_planes // 2)
        self.conv2 = conv3x3(out_planes // 2, out_planes // 4)
        self.bn3 = nn.BatchNorm2d(out_planes // 4)
        self.conv3 = conv3x3(out_planes // 4, out_planes // 4)
        self.downsample = None
        if in_planes != out_planes:
            self.downsample = nn.Sequential(nn.BatchNorm2d(in_planes), nn.ReLU(True),
                                            nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=1, bias=False))


class HourGlass(nn.Module):
    def __init__(self, num_modules, depth, num_features):
        super().__init__()
        self.num_modules, self.depth, self.features = num_modules, depth, num_features
        self.dropout = nn.Dropout(0.5)
        self._generate_network(depth)

    def _generate_network(self, level):
        self.add_module("b1_" + str(level), ConvBlock(256, 256))
        self.add_module("b2_" + str(level), ConvBlock(256, 256))
        if level > 1:
            self._generate_network(level - 1)
        else:
            self.add_module("b2_plus_" + str(level), ConvBlock(256, 256))
        self.add_module("b3_" + str(level), ConvBlock(256, 256))


class FAN_use(nn.Module):
    def __init__(self):
        super().__init__()
        self.num_modules = 1
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3)
        self.bn1 = nn.BatchNorm2d(64)
        self.conv2 = ConvBlock(64, 128)
        self.conv3 = ConvBlock(128, 128)
        self.conv4 = ConvBlock(128, 256)
        self.add_module("m0", HourGlass(1, 4, 256))
        self.add_module("top_m_0", ConvBlock(256, 256))
        self.add_module("conv_last0", nn.Conv2d(256, 256, kernel_size=1, stride=1, padding=0))
        self.add_module("l0", nn.Conv2d(256, 68, kernel_size=1, stride=1, padding=0))
        self.add_module("bn_end0", nn.BatchNorm2d(256))
        self.avgpool = nn.MaxPool2d((2, 2), 2)
        self.conv6 = nn.Conv2d(68, 1, 3, 2, 1)
        self.fc = nn.Linear(28 * 28, 512)
        self.bn5 = nn.BatchNorm2d(68)
        self.relu = nn.ReLU(True)


def _bn_affine(bn):
    """eval-mode BatchNorm as y = x * scale + shift."""
    scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float().contiguous()
    shift = (bn.bias - bn.running_mean * scale).detach().float().contiguous()
    return scale, shift


class FanEncoder(nn.Module):
    def __init__(self, opt=None, pose_dim=6, eye_dim=6):
        super().__init__()
        self.opt = opt
        self.model = FAN_use()

        def head():
            return nn.Sequential(nn.Linear(512, 512), nn.ReLU(), nn.BatchNorm1d(512), nn.Linear(512, 512))

        self.to_mouth = head()
        self.mouth_embed = nn.Sequential(nn.ReLU(), nn.Linear(512, 512 - pose_dim - eye_dim))
        self.to_headpose = head()
        self.headpose_embed = nn.Sequential(nn.ReLU(), nn.Linear(512, pose_dim))
        self.to_eye = head()
        self.eye_embed = nn.Sequential(nn.ReLU(), nn.Linear(512, eye_dim))
        self.to_emo = head()
        self.emo_embed = nn.Sequential(nn.ReLU(), nn.Linear(512, 30))
        # 60 stacked convolutions on bf16 operands leave ~2 % relative error on the embeddings (measured, tests/test_gpu_fan.py) - more
        # than the 1e-2 the bf16 mode of the audio path is held to - so the tensor-core default here is TF32 (fp32 operands rounded to
        # a 10-bit significand, tcgen05.mma.kind::tf32: 3e-3 measured, 2.8x the CUDA-core fp32 rate). AVI_B200_PRECISION=fp32 keeps
        # the exact CUDA-core path; AVI_B200_FAN_PRECISION={fp32,tf32,bf16} (or .precision) overrides.
        self.precision = os.environ.get("AVI_B200_FAN_PRECISION",
                                        "fp32" if os.environ.get("AVI_B200_PRECISION", "bf16").lower() == "fp32" else "tf32").lower()
        if self.precision not in ("bf16", "tf32", "fp32"):
            raise ValueError("AVI_B200_FAN_PRECISION must be bf16, tf32 or fp32")
        self.max_images_per_call = 16          # bounds the transient operand buffers of one chunk
        self._packed, self._packed_key = None, None

    # ------------------------------------------------------------------ packing
    @torch.no_grad()
    def _pack(self):
        key = (self.precision, ops.WEIGHT_EPOCH) + tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        if self._packed is not None and key == self._packed_key:
            return self._packed
        bf16 = self.precision == "bf16"
        tf32 = self.precision == "tf32"
        dt = torch.bfloat16 if bf16 else torch.float32

        def conv_w(conv):
            """[Cout, Cin, kh, kw] -> [Cout, Kpad] with columns (ky, kx, c), zero padded to a multiple of 64."""
            w = conv.weight.detach().float()
            co = w.shape[0]
            flat = w.permute(0, 2, 3, 1).reshape(co, -1)
            K = flat.shape[1]
            Kp = (K + 63) // 64 * 64
            out = torch.zeros(co, Kp, dtype=torch.float32, device=w.device)
            out[:, :K] = flat
            return (ops.cast_bf16(out) if bf16 else (ops.round_tf32(out) if tf32 else out.contiguous())), Kp

        def lin_w(lin):
            w = lin.weight.detach().float().contiguous()
            K = w.shape[1]
            Kp = (K + 63) // 64 * 64
            if Kp != K:
                wp = torch.zeros(w.shape[0], Kp, dtype=torch.float32, device=w.device)
                wp[:, :K] = w
                w = wp
            return (ops.cast_bf16(w) if bf16 else (ops.round_tf32(w) if tf32 else w)), Kp

        def conv_w3(conv):
            """3x3 weights for the implicit convolution: [Cout, 9*Cin] with columns (ky, kx, c); None when the channel count does not
            fill a k-block of this precision (then the im2col path is used)."""
            w = conv.weight.detach().float()
            cin = w.shape[1]
            if cin % (64 if bf16 else 32) != 0:
                return None
            m = w.permute(0, 2, 3, 1).reshape(w.shape[0], 9 * cin).contiguous()
            return ops.cast_bf16(m) if bf16 else (ops.round_tf32(m) if tf32 else m)

        def block(cb):
            d = {"bn": [_bn_affine(cb.bn1), _bn_affine(cb.bn2), _bn_affine(cb.bn3)],
                 "w": [conv_w(cb.conv1), conv_w(cb.conv2), conv_w(cb.conv3)],
                 "w3": [conv_w3(cb.conv1), conv_w3(cb.conv2), conv_w3(cb.conv3)],
                 "cout": [cb.conv1.out_channels, cb.conv2.out_channels, cb.conv3.out_channels], "down": None}
            if cb.downsample is not None:
                d["down"] = (_bn_affine(cb.downsample[0]), conv_w(cb.downsample[2]))
            return d

        f = self.model
        P = {"dt": dt}
        P["conv1_w"] = conv_w(f.conv1)
        s1, h1 = _bn_affine(f.bn1)
        P["conv1_aff"] = (s1, (f.conv1.bias.detach().float() * s1 + h1).contiguous())       # bn1(conv + bias)
        P["conv2"], P["conv3"], P["conv4"] = block(f.conv2), block(f.conv3), block(f.conv4)
        P["hg"] = {n: block(m) for n, m in f.m0.named_children() if isinstance(m, ConvBlock)}
        P["top"] = block(f.top_m_0)
        P["last_w"] = conv_w(f.conv_last0)
        se, he = _bn_affine(f.bn_end0)
        P["last_aff"] = (se, (f.conv_last0.bias.detach().float() * se + he).contiguous())
        P["l_w"] = conv_w(f.l0)
        s5, h5 = _bn_affine(f.bn5)
        P["l_aff"] = (s5, (f.l0.bias.detach().float() * s5 + h5).contiguous())
        P["conv6_w"], P["conv6_b"] = conv_w(f.conv6), f.conv6.bias.detach().float().contiguous()
        P["fc_w"], P["fc_b"] = lin_w(f.fc), f.fc.bias.detach().float().contiguous()

        def head(seq, emb):
            sb, hb = _bn_affine(seq[2])
            return {"w0": lin_w(seq[0]), "b0": seq[0].bias.detach().float().contiguous(), "bn": (sb, hb),
                    "w1": lin_w(seq[3]), "b1": seq[3].bias.detach().float().contiguous(),
                    "we": lin_w(emb[1]), "be": emb[1].bias.detach().float().contiguous()}

        P["heads"] = {"mouth": head(self.to_mouth, self.mouth_embed), "headpose": head(self.to_headpose, self.headpose_embed),
                      "eye": head(self.to_eye, self.eye_embed), "emo": head(self.to_emo, self.emo_embed)}
        self._packed, self._packed_key = P, key
        return P

    # ------------------------------------------------------------------ pieces (rows = NHWC fp32, padded-width lines of W + 2 pixels)
    def _conv(self, x, N, H, W, C, wk, cout, k, stride, pad, pre=None, out=None, bias=None, act=ACT_NONE, Wp_in=None, Wo_extra=0):
        """im2col + GEMM convolution of [relu(x * scale + shift) if pre else x] -> (rows [N*Ho*(Wo+Wo_extra), cout] or `out`, Ho, Wo)."""
        w, Kp = wk
        tf32 = self.precision == "tf32"
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        cols = ops.im2col_affine(x, N, H, W, C, k, stride, pad, Kp, w.dtype, pre, tf32=tf32, Wp_in=Wp_in, Wo_extra=Wo_extra)
        rows = N * Ho * (Wo + Wo_extra)
        if out is None:
            out = torch.empty((rows, cout), dtype=torch.float32, device=x.device)
        ops.gemm(cols, w, bias, out, rows=rows, N=cout, K=Kp, act=act, a_rows_alloc=rows, c_ld=out.stride(0),
                 algorithmic_flops=2.0 * N * Ho * Wo * cout * k * k * C, tf32=tf32)
        return out, Ho, Wo

    def _conv3x3(self, x, N, H, W, C, w3, wk, cout, pre, out):
        """3x3 / stride 1 / pad 1 convolution in the padded-width layout, implicit (one 2-D-tap conv-mode GEMM over the activated operand);
        `out` [N*H*(W+2), cout] may be a channel slice of a wider buffer."""
        Wp = W + 2
        if w3 is None:       # channel count below one k-block of this precision: explicit im2col
            return self._conv(x, N, H, W, C, wk, cout, 3, 1, 1, pre=pre, out=out, Wp_in=Wp, Wo_extra=2)[0]
        tf32 = self.precision == "tf32"
        a = ops.pad_act(x, N, H, W, C, w3.dtype, pre, tf32=tf32)
        rows = H * Wp
        ld = out.stride(0)
        # ONE contraction over the 9 taps: tap (ky, kx) of output pixel r reads operand row r + ky*Wp + kx (2-D taps of the GEMM)
        ops.gemm(a, w3, None, out, batch=N, rows=rows, N=cout, K=9 * C, conv_taps=9, conv_stride=1, conv_taps_x=3, conv_row_pitch=Wp,
                 a_ld=C, a_batch_stride=(H + 2) * Wp * C, a_rows_alloc=(H + 2) * Wp + 2, c_ld=ld, c_batch_stride=rows * ld, tf32=tf32,
                 algorithmic_flops=2.0 * N * H * W * cout * 9 * C)
        return out

    def _block(self, x, N, H, W, cin, B):
        """ConvBlock.forward (FAN_feature_extractor.py:35-59) on padded-width rows [N*H*(W+2), cin]."""
        c1, c2, c3 = B["cout"]
        Wp = W + 2
        cat = torch.empty((N * H * Wp, c1 + c2 + c3), dtype=torch.float32, device=x.device)
        self._conv3x3(x, N, H, W, cin, B["w3"][0], B["w"][0], c1, B["bn"][0], cat[:, :c1])
        self._conv3x3(cat[:, :c1], N, H, W, c1, B["w3"][1], B["w"][1], c2, B["bn"][1], cat[:, c1:c1 + c2])
        self._conv3x3(cat[:, c1:c1 + c2], N, H, W, c2, B["w3"][2], B["w"][2], c3, B["bn"][2], cat[:, c1 + c2:])
        if B["down"] is not None:     # bn -> relu -> 1x1 conv: a row-wise GEMM, the don't-care pixels ride along
            res, _, _ = self._conv(x, N, H, Wp, cin, B["down"][1], c1 + c2 + c3, 1, 1, 0, pre=B["down"][0])
        else:
            res = x
        return ops.add_f32(cat, res if res.is_contiguous() else res.contiguous())

    def _hourglass(self, level, x, N, H, W, P):
        """HourGlass._forward (:81-101); dropout inactive in eval."""
        up1 = self._block(x, N, H, W, 256, P["hg"][f"b1_{level}"])
        H2, W2 = H // 2, W // 2
        low1 = ops.maxpool2x2(x, N, H, W, 256, Wp_in=W + 2, Wp_out=W2 + 2)
        low1 = self._block(low1, N, H2, W2, 256, P["hg"][f"b2_{level}"])
        if level > 1:
            low2 = self._hourglass(level - 1, low1, N, H2, W2, P)
        else:
            low2 = self._block(low1, N, H2, W2, 256, P["hg"][f"b2_plus_{level}"])
        low3 = self._block(low2, N, H2, W2, 256, P["hg"][f"b3_{level}"])
        return ops.upsample_bilinear_add(low3, up1, N, H2, W2, H, W, 256, Wp_in=W2 + 2, Wp_out=W + 2)

    def _features(self, img, P):
        """FAN_use.forward (:139-163): [n,3,224,224] -> [n,512]."""
        n, _, H, W = img.shape
        x = img.permute(0, 2, 3, 1).contiguous().float().reshape(n * H * W, 3)                       # dense NHWC rows of the image
        x, H, W = self._conv(x, n, H, W, 3, P["conv1_w"], 64, 7, 2, 3, Wo_extra=2)                   # -> padded-width rows
        ops.affine_act(x, *P["conv1_aff"], relu=True)                                                # relu(bn1(conv1(x)))
        x = self._block(x, n, H, W, 64, P["conv2"])
        x = ops.maxpool2x2(x, n, H, W, 128, Wp_in=W + 2, Wp_out=W // 2 + 2)
        H, W = H // 2, W // 2
        x = self._block(x, n, H, W, 128, P["conv3"])
        x = self._block(x, n, H, W, 128, P["conv4"])
        hg = self._hourglass(4, x, n, H, W, P)
        ll = self._block(hg, n, H, W, 256, P["top"])
        Wp = W + 2
        ll, _, _ = self._conv(ll, n, H, Wp, 256, P["last_w"], 256, 1, 1, 0)
        ops.affine_act(ll, *P["last_aff"], relu=True)                                                # relu(bn_end(conv_last(ll)))
        t, _, _ = self._conv(ll, n, H, Wp, 256, P["l_w"], 68, 1, 1, 0)
        ops.affine_act(t, *P["l_aff"], relu=True)                                                    # relu(bn5(l(.)))
        net, Ho, Wo = self._conv(t, n, H, W, 68, P["conv6_w"], 1, 3, 2, 1, bias=P["conv6_b"], act=ACT_RELU, Wp_in=Wp)
        net = net.view(n, Ho * Wo)                                                                   # [n, 784] (dense again)
        return self._linear(net, P["fc_w"], P["fc_b"])

    def _linear(self, x, wk, b, act=ACT_NONE):
        w, Kp = wk
        rows, K = x.shape
        xo = ops.cast_pad2d(x, w.dtype, C_pad=Kp) if (Kp != K or w.dtype != torch.float32) else x
        out = torch.empty((rows, w.shape[0]), dtype=torch.float32, device=x.device)
        ops.gemm(xo, w, b, out, rows=rows, N=w.shape[0], K=Kp, act=act, a_rows_alloc=rows, tf32=self.precision == "tf32")
        return out

    def _head(self, x, Hd):
        h = self._linear(x, Hd["w0"], Hd["b0"], act=ACT_RELU)
        ops.affine_act(h, *Hd["bn"], relu=False)
        feat = self._linear(h, Hd["w1"], Hd["b1"])
        r = feat.clone()
        ops.affine_act(r, None, None, relu=True)
        return feat, self._linear(r, Hd["we"], Hd["be"])

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def forward_feature(self, x):
        return torch.cat([self._features(x[i:i + self.max_images_per_call], self._pack())
                          for i in range(0, x.shape[0], self.max_images_per_call)], 0)

    @torch.no_grad()
    def forward(self, x):
        if self.training:
            raise NotImplementedError("FanEncoder drop-in is eval-only (BatchNorm running statistics); call .eval()")
        if not x.is_cuda:
            raise RuntimeError("avi_talking_b200.FanEncoder runs on CUDA only (no CPU fallback)")
        P = self._pack()
        net = self.forward_feature(x)
        mouth_feat, _ = self._head(net, P["heads"]["mouth"])
        _, headpose_emb = self._head(net, P["heads"]["headpose"])
        _, eye_embed = self._head(net, P["heads"]["eye"])
        _, emo_embed = self._head(net, P["heads"]["emo"])
        return headpose_emb, eye_embed, emo_embed, mouth_feat
