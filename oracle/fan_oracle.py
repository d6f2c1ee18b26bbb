"""Oracle: FanEncoder image branch (TEST INFRASTRUCTURE; see oracle/__init__.py).

fp32 functional-torch restatement, driven by a plain state dict, of
  third_party/pd_fgc_inference/lib/models/networks/FAN_feature_extractor.py  ConvBlock.forward :35-59, HourGlass._forward :81-101,
                                                                             FAN_use.forward :139-163
  third_party/pd_fgc_inference/lib/models/networks/encoder.py                FanEncoder.forward :116-126
in eval mode (BatchNorm running statistics, hourglass dropout inactive), as Faceformer.predict calls it under no_grad
(models/faceformer_disentangle.py:783-797). Pinned by tests/golden/fan.npz: outputs of the reference's own FanEncoder class (imported
with an omegaconf stub for its YAML read) on the seeded state dict of avi_talking_b200.synth.fan_state.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"], False, 0.0, 1e-5)


def conv_block(sd, p, x):
    """ConvBlock.forward :35-59."""
    o1 = F.conv2d(F.relu(_bn(sd, p + "bn1.", x)), sd[p + "conv1.weight"], padding=1)
    o2 = F.conv2d(F.relu(_bn(sd, p + "bn2.", o1)), sd[p + "conv2.weight"], padding=1)
    o3 = F.conv2d(F.relu(_bn(sd, p + "bn3.", o2)), sd[p + "conv3.weight"], padding=1)
    out = torch.cat((o1, o2, o3), 1)
    if p + "downsample.2.weight" in sd:
        x = F.conv2d(F.relu(_bn(sd, p + "downsample.0.", x)), sd[p + "downsample.2.weight"])
    return out + x


def hourglass(sd, p, level, x):
    """HourGlass._forward :81-101 (dropout inactive)."""
    up1 = conv_block(sd, f"{p}b1_{level}.", x)
    low1 = conv_block(sd, f"{p}b2_{level}.", F.max_pool2d(x, 2, stride=2))
    low2 = hourglass(sd, p, level - 1, low1) if level > 1 else conv_block(sd, f"{p}b2_plus_{level}.", low1)
    low3 = conv_block(sd, f"{p}b3_{level}.", low2)
    return up1 + F.interpolate(low3, size=up1.shape[2:], mode="bilinear", align_corners=False)


def fan_features(sd, x, p="model."):
    """FAN_use.forward :139-163: [N,3,224,224] -> [N,512]."""
    x = F.relu(_bn(sd, p + "bn1.", F.conv2d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"], stride=2, padding=3)))
    x = F.max_pool2d(conv_block(sd, p + "conv2.", x), 2)
    x = conv_block(sd, p + "conv4.", conv_block(sd, p + "conv3.", x))
    ll = conv_block(sd, p + "top_m_0.", hourglass(sd, p + "m0.", 4, x))
    ll = _bn(sd, p + "bn_end0.", F.conv2d(ll, sd[p + "conv_last0.weight"], sd[p + "conv_last0.bias"]))
    t = F.conv2d(F.relu(ll), sd[p + "l0.weight"], sd[p + "l0.bias"])
    net = F.conv2d(F.relu(_bn(sd, p + "bn5.", t)), sd[p + "conv6.weight"], sd[p + "conv6.bias"], stride=2, padding=1)
    net = F.relu(net.view(-1, net.shape[-2] * net.shape[-1]))
    return F.linear(net, sd[p + "fc.weight"], sd[p + "fc.bias"])


def _head(sd, p, x):
    h = F.relu(F.linear(x, sd[p + "0.weight"], sd[p + "0.bias"]))
    h = F.batch_norm(h, sd[p + "2.running_mean"], sd[p + "2.running_var"], sd[p + "2.weight"], sd[p + "2.bias"], False, 0.0, 1e-5)
    return F.linear(h, sd[p + "3.weight"], sd[p + "3.bias"])


@torch.no_grad()
def fan_encoder_forward(sd, x):
    """FanEncoder.forward :116-126 -> (headpose_emb [N,6], eye_embed [N,6], emo_embed [N,30], mouth_feat [N,512])."""
    net = fan_features(sd, x)
    mouth = _head(sd, "to_mouth.", net)
    emb = lambda feat, p: F.linear(F.relu(feat), sd[p + "1.weight"], sd[p + "1.bias"])  # noqa: E731
    return (emb(_head(sd, "to_headpose.", net), "headpose_embed."), emb(_head(sd, "to_eye.", net), "eye_embed."),
            emb(_head(sd, "to_emo.", net), "emo_embed."), mouth)
