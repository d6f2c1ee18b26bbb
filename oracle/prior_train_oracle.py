"""Oracle: one TRAINING iteration of the text->style diffusion prior (TEST INFRASTRUCTURE; see oracle/__init__.py).

fp32 torch-on-CPU restatement, differentiated by torch autograd, of
  train_diffusion_prior.py    the iteration :434-486 (voxel2clip, diffusion_prior(text_embed, image_embed), normalize,
                              soft_clip_loss :125-133, loss = loss_nce + 30 * loss_prior :474), AdamW groups :996-1004
  models/diffusion_prior.py   InstructDiffusionPrior.forward :404-456, p_losses :369-402, VersatileDiffusionPriorNetwork.forward
                              :223-313 WITH the conditioning dropout (:258-281), BrainNetwork.forward :95-117 with its Dropouts
and of dalle2_pytorch's NoiseScheduler.q_sample / LayerNorm(stable=True) (row maximum detached) underneath.

PARITY UNPINNED for the dalle2_pytorch / rotary_embedding_torch semantics (see oracle/prior_oracle.py); the reference's own
arithmetic is pinned by tests/golden/prior_train.npz, minted by oracle/make_golden.golden_prior_train from the reference's class
sources executed over oracle/dalle2_standin.py. Every stochastic draw (timesteps, noise, keep masks, dropout masks) is an input.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import prior_oracle as po

NO_DECAY = ("bias", "LayerNorm.bias", "LayerNorm.weight")     # train_diffusion_prior.py:997 (substring match on parameter names)


def brain_network_train(sd, x, dropout_masks=None, p="voxel2clip.", n_blocks=4, clip_size=po.DIM):
    """BrainNetwork.forward :95-117; dropout_masks: 1 + n_blocks pre-scaled masks (nn.Dropout(0.5) :66, nn.Dropout(0.15) :72) or None."""
    def block(x, q, k):
        h = x @ sd[q + "0.weight"].t() + sd[q + "0.bias"]
        h = F.gelu(F.layer_norm(h, h.shape[-1:], sd[q + "1.weight"], sd[q + "1.bias"], 1e-5))
        return h if dropout_masks is None else h * dropout_masks[k]
    x = block(x, p + "lin0.", 0)
    residual = x
    for i in range(n_blocks):
        x = block(x, p + f"mlp.{i}.", i + 1) + residual
        residual = x
    x = x @ sd[p + "lin1.weight"].t() + sd[p + "lin1.bias"]
    h = x.reshape(len(x), -1, clip_size)
    q = p + "projector."
    for ln_i, lin_i in ((0, 2), (3, 5), (6, 8)):
        h = F.gelu(F.layer_norm(h, h.shape[-1:], sd[q + f"{ln_i}.weight"], sd[q + f"{ln_i}.bias"], 1e-5))
        h = h @ sd[q + f"{lin_i}.weight"].t() + sd[q + f"{lin_i}.bias"]
    return x, h


def layernorm_stable_detached(x, g, eps=1e-5):
    """dalle2_pytorch.LayerNorm(stable=True): x / x.amax(-1).detach() first."""
    x = x / x.amax(dim=-1, keepdim=True).detach()
    var = torch.var(x, dim=-1, unbiased=False, keepdim=True)
    mean = torch.mean(x, dim=-1, keepdim=True)
    return (x - mean) * (var + eps).rsqrt() * g


def prior_net_train(sd, image_embed, t, text_embed, keep_brain, keep_image, depth=po.DEPTH):
    """VersatileDiffusionPriorNetwork.forward :223-313 with the keep masks of :258-281 given (bool [B])."""
    B = image_embed.shape[0]
    image_embed = image_embed.view(B, -1, po.DIM)
    brain = text_embed.view(B, -1, po.DIM)
    brain = torch.where(keep_brain.view(B, 1, 1), brain, sd["net.null_brain_embeds"][None])           # :265-270
    image_embed = torch.where(keep_image.view(B, 1, 1), image_embed, sd["net.null_image_embed"][None])  # :273-278
    time_embed = po.time_mlp(sd, "net.to_time_embeds.0.1.", po.sinusoidal_pos_emb(t))[:, None, :]
    image_embed = image_embed + sd["net.learned_query"][None]                                          # :290-292
    x = torch.cat((brain, time_embed, image_embed), dim=-2)
    n = x.shape[1]
    bias = po.rel_pos_bias(sd["net.causal_transformer.rel_pos_bias.relative_attention_bias.weight"], n, n + 1)
    freqs = po.rotary_freqs()
    for l in range(depth):
        p = f"net.causal_transformer.layers.{l}."
        x = po.attention(sd, p + "0.", x, bias, freqs) + x
        x = po.feedforward(sd, p + "1.", x) + x
    out = layernorm_stable_detached(x, sd["net.causal_transformer.norm.g"])
    out = out @ sd["net.causal_transformer.project_out.weight"].t()
    return out[..., -1:, :]


def soft_clip_loss(preds, targs, temp=0.125):
    """train_diffusion_prior.py:125-133."""
    clip_clip = (targs @ targs.T) / temp
    brain_clip = (preds @ targs.T) / temp
    loss1 = -(brain_clip.log_softmax(-1) * clip_clip.softmax(-1)).sum(-1).mean()
    loss2 = -(brain_clip.T.log_softmax(-1) * clip_clip.softmax(-1)).sum(-1).mean()
    return (loss1 + loss2) / 2


def losses(sd, voxel, clip_target, times, noise, keep_brain, keep_image, temp, dropout_masks=None):
    """-> (loss_nce, loss_prior, pred) of one iteration (:441-468); clip_target [B,1,128]."""
    B = voxel.shape[0]
    sched = po.noise_schedule(100)
    clip_voxels, proj = brain_network_train(sd, voxel, dropout_masks)
    x0 = clip_target * (po.DIM ** 0.5)                                                                # :453
    sa, s1 = sched["sqrt_alphas_cumprod"][times].view(B, 1, 1), sched["sqrt_one_minus_alphas_cumprod"][times].view(B, 1, 1)
    xt = sa * x0 + s1 * noise                                                                         # q_sample :372
    pred = prior_net_train(sd, xt, times, clip_voxels.view(B, -1, po.DIM), keep_brain, keep_image)
    loss_prior = F.mse_loss(pred, x0)                                                                 # :401
    pn = F.normalize(proj.flatten(1), dim=-1)                                                         # :455-456
    tn = F.normalize(clip_target.flatten(1), dim=-1)
    return soft_clip_loss(pn, tn, temp), loss_prior, pred


def train_step(sd, voxel, clip_target, times, noise, keep_brain, keep_image, temp, dropout_masks=None, prior_mult=30.0, lr=3e-4,
               weight_decay=1e-2):
    """One iteration: losses, gradients of every trainable tensor, one torch.optim.AdamW step with the reference's groups.
    -> dict(loss_nce, loss_prior, pred, grads {name: tensor}, new {name: updated tensor})."""
    sd = {k: v.clone().requires_grad_(not k.endswith("rotary_emb.freqs")) for k, v in sd.items()}
    loss_nce, loss_prior, pred = losses(sd, voxel, clip_target, times, noise, keep_brain, keep_image, temp, dropout_masks)
    (loss_nce + prior_mult * loss_prior).backward()
    names = [k for k, v in sd.items() if v.requires_grad]

    def local(k):       # the name as the reference's named_parameters() of net / voxel2clip yields it
        return k.split(".", 1)[1]
    groups = [{"params": [sd[k] for k in names if not any(nd in local(k) for nd in NO_DECAY)], "weight_decay": weight_decay},
              {"params": [sd[k] for k in names if any(nd in local(k) for nd in NO_DECAY)], "weight_decay": 0.0}]
    grads = {k: sd[k].grad.detach().clone() for k in names}
    opt = torch.optim.AdamW(groups, lr=lr)
    opt.step()
    return dict(loss_nce=loss_nce.detach(), loss_prior=loss_prior.detach(), pred=pred.detach(), grads=grads,
                new={k: sd[k].detach() for k in names})
