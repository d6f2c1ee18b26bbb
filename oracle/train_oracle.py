"""Oracle: the teacher-forced faceformer_vert training step (TEST INFRASTRUCTURE; see oracle/__init__.py).

fp32 torch-on-CPU restatement, with torch.autograd supplying the gradients, of
  models/faceformer_vert.py  Faceformer.forward_switch_frame :360-371,405-412,434-454,475-482
     obj_vector(one_hot[:,0]=1) ; audio_encoder(audio, dataset, frame_num=T) ; audio_feature_map ;
     gt_verts = convert_coeff2verts(coeff[:, :, :53], pose, zeros_like(shape)) (:348-356) ;
     per clip: vertice_map(cat[template, gt[:-1]] - template) + style -> PPE -> TransformerDecoder(tgt_mask = biased mask,
     memory_mask = enc_dec_mask) -> vertice_map_r ; loss = mean(criterion(out + template, gt)) * 10
  models/faceformer_vert.py:154  feature_extractor._freeze_parameters()  (no gradients for the 7 conv layers + GroupNorm)
plus one torch.optim.Adam step (no trainer for this class is published, SURVEY 3.3: the harness assembles Adam, lr 1e-4).

Regularisers: by default every dropout / SpecAugment / LayerDrop is inactive (the modules are evaluated as in .eval()) so that the step
is a deterministic function of its inputs: the mode of the golden fixture tests/golden/train.npz (minted by oracle/make_golden.py from
the reference's OWN forward_switch_frame + loss.backward()). With `reg` (synth.train_regularisers) the step runs in TRAIN mode with every
draw given as a tensor - dropout masks of HF Wav2Vec2 / the PPE / nn.TransformerDecoderLayer, the SpecAugment time mask
(models/lib/wav2vec.py:120-131), LayerDrop - against tests/golden/train_reg.npz, minted from the reference in .train() mode with
torch.nn.functional.dropout, torch.rand and _compute_mask_indices replaced by injectors of the same tensors.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import faceformer_oracle as ffo
from . import wav2vec2_oracle as w2o

FROZEN_PREFIX = "feature_extractor."          # faceformer_vert.py:154
UNUSED_W2V = ("masked_spec_embed",)           # only touched by SpecAugment


def trainable(sd_ff: dict, sd_w2v: dict, spec_augment: bool = False) -> dict:
    """name -> tensor for everything optimizer would see (audio_encoder.* prefixed like the module's state_dict)."""
    out = {k: v for k, v in sd_ff.items()}
    for k, v in sd_w2v.items():
        if not k.startswith(FROZEN_PREFIX) and (spec_augment or k not in UNUSED_W2V):
            out["audio_encoder." + k] = v
    return out


def loss_fn(sd_ff, sd_w2v, template, audio, gt_verts, period=30, dataset="vocaset", n_subjects=8, reg=None):
    """:360-371,434-454,475-482 with gt_verts [B,T,V3] already converted. Differentiable w.r.t. the state-dict tensors."""
    B, T = gt_verts.shape[0], gt_verts.shape[1]
    one_hot = torch.zeros(B, n_subjects)
    one_hot[:, 0] = 1
    obj = F.linear(one_hot, sd_ff["obj_vector.weight"])                                        # :363-365
    ha = w2o.wav2vec2_forward.__wrapped__(sd_w2v, audio, frame_num=T, reg=reg)                 # :367-368
    ha = F.linear(ha, sd_ff["audio_feature_map.weight"], sd_ff["audio_feature_map.bias"])      # :369
    out = ffo.forward_ff.__wrapped__(sd_ff, template, ha, obj, T, True, gt_verts=gt_verts, period=period, dataset=dataset,
                                     merge=False, reg=reg)                                     # :437-454,475
    return torch.mean((out - gt_verts) ** 2) * 10.0                                            # :481-482 (criterion = MSE)


def train_step(sd_ff, sd_w2v, template, audio, gt_verts, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, period=30, steps=1, reg=None):
    """-> (losses [steps], grads of the first step {name: tensor}, params after `steps` Adam steps {name: tensor}).
    reg (avi_talking_b200.synth.train_regularisers): the step in TRAIN mode with every dropout mask, the SpecAugment mask and the
    LayerDrop decisions given (the same draws every step)."""
    sd_ff = {k: v.clone() for k, v in sd_ff.items()}
    sd_w2v = {k: v.clone() for k, v in sd_w2v.items()}
    params = trainable(sd_ff, sd_w2v, spec_augment=reg is not None and reg.get("spec_mask") is not None)
    for v in params.values():
        v.requires_grad_(True)
    opt = torch.optim.Adam(list(params.values()), lr=lr, betas=betas, eps=eps)
    losses, grads0 = [], None
    for s in range(steps):
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(sd_ff, sd_w2v, template, audio, gt_verts, period=period, reg=reg)
        loss.backward()
        if s == 0:
            grads0 = {k: (v.grad.detach().clone() if v.grad is not None else torch.zeros_like(v)) for k, v in params.items()}
        opt.step()
        losses.append(float(loss.detach()))
    return losses, grads0, {k: v.detach().clone() for k, v in params.items()}
