"""Training step of the text -> style diffusion prior on the CUDA library.

What the reference does per iteration (train_diffusion_prior.py:434-486, models/diffusion_prior.py:369-456):

    clip_voxels, clip_voxels_proj = diffusion_prior.voxel2clip(voxel)                       # BrainNetwork.forward :95-117
    loss_prior, pred = diffusion_prior(text_embed=clip_voxels, image_embed=clip_target)      # forward :404-456 -> p_losses :369-402
    loss_nce = soft_clip_loss(normalize(clip_voxels_proj), normalize(clip_target), temp)     # :125-133, :455-468
    loss = loss_nce + 30 * loss_prior ; loss.backward() ; AdamW step                         # :474-486, :996-1004

Here the three differentiable pieces are explicit forward / backward pairs over the .so (dense contractions: ``avi_gemm_*``
through ``train._Lin``, bf16 tensor-core or fp32 CUDA-core by ``precision``; everything else: csrc/prior_train.cu and the
LayerNorm / activation kernels of csrc/train.cu), glued by ``torch.autograd.Function`` so that the reference's own loop
(``loss.backward()``) runs unchanged on the drop-in classes, and ``PriorTrainStep`` runs the whole iteration without autograd.
Stochastic inputs (timesteps, noise, the two classifier-free-guidance keep masks, BrainNetwork's dropout masks) are drawn with
torch's generator when not supplied; the parity tests supply them.  No CPU or eager fallback.
"""
from __future__ import annotations

import math

import torch

from . import ops
from .ops import ACT_GELU, ACT_SILU
from .train import _Lin


def _acc(p, g):
    """p.grad (+)= g, g in the parameter's shape."""
    g = g.view(p.shape)
    p.grad = g if p.grad is None else ops.add_f32(p.grad.contiguous(), g.contiguous())


def _pad64(n):
    return ((n + 63) // 64) * 64


def _zeros(n, dev):
    return torch.zeros((n,), dtype=torch.float32, device=dev)


class _LinearGrad:
    """One nn.Linear inside a saved graph: fwd through _Lin, bwd writes the weight / bias gradients into fresh buffers."""

    def __init__(self, lin: _Lin):
        self.lin = lin

    def fwd(self, x32, mod, residual=None):
        return self.lin.fwd(self.lin.a(x32), mod.weight, mod.bias, residual=residual)

    def bwd(self, dy32, x32, mod, want_dx=True, residual=None):
        W = mod.weight
        gW = torch.empty_like(W)
        gb = torch.empty_like(mod.bias) if mod.bias is not None else None
        dx = self.lin.bwd(dy32.contiguous(), x32, W, gW, gb, _pad64(dy32.shape[0]), want_dx=want_dx, residual=residual)
        _acc(W, gW)
        if gb is not None:
            _acc(mod.bias, gb)
        return dx


# ------------------------------------------------------------------------------------------------ BrainNetwork
class BrainTrain:
    """BrainNetwork.forward (models/diffusion_prior.py:95-117) with saved activations, and its backward.
    dropout_masks: None (dropout inactive = .eval()) or 1 + n_blocks fp32 [B,h] tensors ALREADY scaled by 1/(1-p)
    (nn.Dropout(0.5) after lin0, nn.Dropout(0.15) in every residual block, :66,:72)."""

    def __init__(self, net, precision):
        self.net, self.lin = net, _LinearGrad(_Lin(precision == "bf16"))
        self.S = None

    def forward(self, voxel, dropout_masks=None):
        net, L = self.net, self.lin
        x = voxel.reshape(voxel.shape[0], -1).contiguous().float()
        S = {"blocks": [], "proj": []}
        res = None
        for k, seq in enumerate([net.lin0] + list(net.mlp)):
            y = L.fwd(x, seq[0])
            z = ops.layernorm(y, seq[1].weight, seq[1].bias, eps=seq[1].eps)[0]
            g = ops.act_fwd(z, ACT_GELU)[0]
            m = None if dropout_masks is None else dropout_masks[k].contiguous()
            if m is not None:
                g = ops.mul_f32(g, m)
            out = g if res is None else ops.add_f32(g, res)                                  # x += residual (:109-110)
            S["blocks"].append((seq, x, y, z, m))
            res = x = out
        S["x_last"] = x
        o = L.fwd(x, net.lin1)                                                               # :113
        h = o.reshape(-1, net.clip_size)
        if net.use_projector:
            pr = net.projector
            for ln_i, lin_i in ((0, 2), (3, 5), (6, 8)):
                z = ops.layernorm(h.contiguous(), pr[ln_i].weight, pr[ln_i].bias, eps=pr[ln_i].eps)[0]
                g = ops.act_fwd(z, ACT_GELU)[0]
                S["proj"].append((pr[ln_i], pr[lin_i], h, z, g))
                h = L.fwd(g, pr[lin_i])
        self.S = S
        return o, (h.reshape(len(o), -1, net.clip_size) if net.use_projector else None)

    def backward(self, d_o=None, d_proj=None):
        """d_o / d_proj: gradients of the two outputs (either may be None). Parameter gradients are accumulated into .grad."""
        net, L, S = self.net, self.lin, self.S
        dev = S["x_last"].device
        dh = None
        if d_proj is not None:
            dh = d_proj.reshape(-1, net.clip_size).contiguous().float()
            for ln, lin_mod, h, z, g in reversed(S["proj"]):
                dg = L.bwd(dh, g, lin_mod)
                dz = ops.act_bwd(z, dg, ACT_GELU)
                dw, db = _zeros(ln.weight.numel(), dev), _zeros(ln.weight.numel(), dev)
                dh = ops.layernorm_bwd(h.contiguous(), ln.weight, dz, dw, db, eps=ln.eps)
                _acc(ln.weight, dw)
                _acc(ln.bias, db)
            dh = dh.reshape(len(S["x_last"]), -1)
        if d_o is not None:
            d_o = d_o.reshape(len(S["x_last"]), -1).contiguous().float()
            dh = d_o if dh is None else ops.add_f32(dh.contiguous(), d_o)
        if dh is None:
            return
        dx = L.bwd(dh, S["x_last"], net.lin1)
        for k in range(len(S["blocks"]) - 1, -1, -1):
            seq, x_in, y, z, m = S["blocks"][k]
            dg = dx if m is None else ops.mul_f32(dx.contiguous(), m)
            dz = ops.act_bwd(z, dg, ACT_GELU)
            dw, db = _zeros(seq[1].weight.numel(), dev), _zeros(seq[1].weight.numel(), dev)
            dy = ops.layernorm_bwd(y, seq[1].weight, dz, dw, db, eps=seq[1].eps)
            _acc(seq[1].weight, dw)
            _acc(seq[1].bias, db)
            # block k >= 1: out_k = act_k + out_{k-1}  ->  d out_{k-1} = dy W + d out_k ; block 0 reads the (constant) voxels
            dx = L.bwd(dy, x_in, seq[0], want_dx=k > 0, residual=dx.contiguous() if k > 0 else None)
        self.S = None


# ------------------------------------------------------------------------------------------------ prior network + p_losses
class PriorLossTrain:
    """InstructDiffusionPrior.forward / p_losses (models/diffusion_prior.py:369-456): q_sample, VersatileDiffusionPriorNetwork.forward
    (:223-313) with conditioning dropout, FlaggedCausalTransformer.forward (:154-166) over dalle2_pytorch's Attention /
    FeedForward / LayerNorm, l2 loss to x_start — forward with saved activations and the backward of all of it."""

    def __init__(self, prior, precision):
        self.prior, self.lin = prior, _LinearGrad(_Lin(precision == "bf16"))
        self.S = None
        self._const = None

    def _constants(self, dev):
        """Architecture constants, built once: sinusoidal features of every integer timestep (SinusoidalPosEmb), the rotary table,
        the relative-position bucket of every (query, key) pair as a one-hot matrix, a zero LayerNorm bias."""
        if self._const is not None and self._const["dev"] == dev:
            return self._const
        net = self.prior.net
        ct = net.causal_transformer
        T = self.prior.noise_scheduler.num_timesteps
        half = net.dim // 2
        f = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
        e = torch.arange(T, dtype=torch.float32)[:, None] * f[None, :]
        sin_table = torch.cat((e.sin(), e.cos()), dim=-1).contiguous().to(dev)
        freqs = ct.layers[0][0].rotary_emb.freqs.detach().float().cpu()
        ang = torch.arange(3, dtype=torch.float32)[:, None] * freqs[None, :]
        rot = torch.stack((ang.cos(), ang.sin()), dim=-1).contiguous().to(dev)
        rpb = ct.rel_pos_bias
        n = torch.clamp(-(torch.arange(4)[None, :] - torch.arange(3)[:, None]), min=0)
        max_exact = rpb.num_buckets // 2
        large = max_exact + (torch.log(n.float() / max_exact) / math.log(rpb.max_distance / max_exact) * (rpb.num_buckets - max_exact)).long()
        large = torch.min(large, torch.full_like(large, rpb.num_buckets - 1))
        bucket = torch.where(n < max_exact, n, large).reshape(-1)                            # [12]
        onehot = torch.zeros((rpb.num_buckets, 64), dtype=torch.float32)                     # K padded to 64 for the GEMM paths
        onehot[bucket, torch.arange(12)] = 1.0
        self._const = dict(dev=dev, sin=sin_table, rot=rot, bucket=bucket.to(dev), onehot=onehot.to(dev), zero=_zeros(4096, dev))
        return self._const

    def forward(self, text_embed, image_embed, times=None, noise=None, keep_brain=None, keep_image=None, generator=None, loss_scale=1.0):
        """image_embed is x_start, i.e. ALREADY multiplied by image_embed_scale (:453). Returns (loss * loss_scale, pred); the stored
        loss gradient carries loss_scale (the prior_mult of train_diffusion_prior.py:474 when the whole iteration is fused)."""
        prior, L = self.prior, self.lin
        net = prior.net
        ct = net.causal_transformer
        dev = image_embed.device
        B, dim = image_embed.shape[0], net.dim
        Cn = self._constants(dev)
        sched = prior.noise_scheduler
        x0 = image_embed.reshape(B, dim).contiguous().float()                                # already * image_embed_scale (:453)
        brain = text_embed.reshape(B, dim).contiguous().float()
        if times is None:                                                                    # sample_random_times (:446)
            times = torch.randint(0, sched.num_timesteps, (B,), device=dev, generator=generator)
        if noise is None:                                                                    # :370
            noise = torch.randn(x0.shape, device=dev, generator=generator)
        if keep_brain is None:                                                               # prob_mask_like(1 - drop_prob) :258-262
            keep_brain = torch.rand((B,), device=dev, generator=generator) < (1.0 - prior.text_cond_drop_prob)
        if keep_image is None:
            keep_image = torch.rand((B,), device=dev, generator=generator) < (1.0 - prior.image_cond_drop_prob)
        keep_b, keep_i = keep_brain.reshape(B).float().contiguous(), keep_image.reshape(B).float().contiguous()
        # time embedding: SinusoidalPosEmb (table row gather) -> MLP(dim, dim, expansion 2, depth 2) :186-189,286
        mlp = net.to_time_embeds[0][1].net
        e0 = Cn["sin"].index_select(0, times.long())
        a0 = L.fwd(e0, mlp[0][0])
        s0 = ops.act_fwd(a0, ACT_SILU)[0]
        a1 = L.fwd(s0, mlp[1][0])
        s1 = ops.act_fwd(a1, ACT_SILU)[0]
        temb = L.fwd(s1, mlp[2])
        tokens, _ = ops.prior_tokens_fwd(brain, net.null_brain_embeds.reshape(-1), keep_b, x0, noise.reshape(B, dim).contiguous().float(),
                                         sched.sqrt_alphas_cumprod, sched.sqrt_one_minus_alphas_cumprod, times.to(torch.int32).contiguous(),
                                         net.null_image_embed.reshape(-1), keep_i, net.learned_query.reshape(-1), temb)
        bias = ct.rel_pos_bias.relative_attention_bias.weight.detach()[Cn["bucket"]].t().contiguous()     # gather: [heads, 12]
        zero = Cn["zero"]
        x = tokens.view(B * 3, dim)
        S = dict(B=B, keep_b=keep_b, keep_i=keep_i, time=(e0, a0, s0, a1, s1), layers=[])
        for attn, ff in ct.layers:
            xn = ops.layernorm(x, attn.norm.g, zero[:dim], eps=attn.norm.eps)[0]
            q = L.fwd(xn, attn.to_q)
            kv = L.fwd(xn, attn.to_kv)
            att, P = ops.prior_attn_fwd(q, kv, attn.null_kv, Cn["rot"], bias, B, attn.heads, attn.null_kv.shape[1])
            o = L.fwd(att, attn.to_out[0])
            on = ops.layernorm(o, attn.to_out[1].g, zero[:dim], eps=attn.to_out[1].eps)[0]
            x1 = ops.add_f32(on, x)                                                          # attn(x) + x :162
            fn = ops.layernorm(x1, ff[0].g, zero[:dim], eps=ff[0].eps)[0]
            h = L.fwd(fn, ff[1])
            s = ops.swiglu_fwd(h)
            x2 = L.fwd(s, ff[5], residual=x1)                                                # ff(x) + x :163
            S["layers"].append((attn, ff, x, xn, q, kv, P, att, o, x1, fn, h, s))
            x = x2
        xs, amax = ops.rows_stat_div(x, 0)                                                   # LayerNorm(stable=True): x / amax(x).detach()
        xf = ops.layernorm(xs, ct.norm.g, zero[:dim], eps=ct.norm.eps)[0]
        out = L.fwd(xf, ct.project_out)
        pred = out.view(B, 3, dim)[:, 2]                                                     # tokens[..., -1:, :] :311
        loss, dpred = ops.mse_loss_grad(pred, x0, float(loss_scale))                         # F.mse_loss(pred, image_embed) :399-401
        S.update(xs=xs, amax=amax, xf=xf, dpred=dpred)
        self.S = S
        return loss, pred.contiguous().view(B, 1, dim)            # a fresh tensor: the caller divides it in place (:450)

    def backward(self, gloss=None, dpred_extra=None):
        """-> d loss / d text_embed [B,1,dim]; parameter gradients are accumulated into .grad. gloss (autograd glue only): the
        upstream gradient of the loss as a 0-d tensor; dpred_extra: an upstream gradient on the returned prediction (normally none)."""
        prior, L, S = self.prior, self.lin, self.S
        net = prior.net
        ct = net.causal_transformer
        B, dim = S["B"], net.dim
        dev = S["xs"].device
        Cn = self._constants(dev)
        zero = Cn["zero"]
        dpred = S["dpred"] if gloss is None else S["dpred"] * gloss
        if dpred_extra is not None:
            dpred = dpred + dpred_extra.reshape(B, dim)
        dout = torch.zeros((B, 3, dim), dtype=torch.float32, device=dev)
        dout[:, 2] = dpred
        dxf = L.bwd(dout.view(B * 3, dim), S["xf"], ct.project_out)
        dg, dummy = _zeros(dim, dev), _zeros(dim, dev)
        dxs = ops.layernorm_bwd(S["xs"], ct.norm.g, dxf, dg, dummy, eps=ct.norm.eps)
        _acc(ct.norm.g, dg)
        dx = ops.rows_stat_div_bwd(S["xs"], dxs, S["amax"], 0)
        dbias = None
        for attn, ff, x, xn, q, kv, P, att, o, x1, fn, h, s in reversed(S["layers"]):
            ds = L.bwd(dx, s, ff[5])
            dh = ops.swiglu_bwd(h, ds)
            dfn = L.bwd(dh, fn, ff[1])
            dg = _zeros(dim, dev)
            dx1 = ops.add_f32(ops.layernorm_bwd(x1, ff[0].g, dfn, dg, dummy, eps=ff[0].eps), dx.contiguous())
            _acc(ff[0].g, dg)
            dg = _zeros(dim, dev)
            do = ops.layernorm_bwd(o, attn.to_out[1].g, dx1, dg, dummy, eps=attn.to_out[1].eps)
            _acc(attn.to_out[1].g, dg)
            datt = L.bwd(do, att, attn.to_out[0])
            dq, dkv, dnull, db = ops.prior_attn_bwd(q, kv, attn.null_kv, Cn["rot"], P, datt, B, attn.heads, attn.null_kv.shape[1])
            _acc(attn.null_kv, dnull)
            dbias = db if dbias is None else ops.add_f32(dbias, db)
            dxn = ops.add_f32(L.bwd(dq, xn, attn.to_q), L.bwd(dkv, xn, attn.to_kv))
            dg = _zeros(dim, dev)
            dx = ops.add_f32(ops.layernorm_bwd(x, attn.norm.g, dxn, dg, dummy, eps=attn.norm.eps), dx1)
            _acc(attn.norm.g, dg)
        # relative-position bias: the [heads, 3, 4] gradient scattered onto the 32 x heads embedding = one-hot[32, 12] @ dbias^T
        emb = ct.rel_pos_bias.relative_attention_bias.weight
        db_pad = torch.zeros((emb.shape[1], 64), dtype=torch.float32, device=dev)
        db_pad[:, :12] = dbias.view(emb.shape[1], 12)
        _acc(emb, ops.linear(Cn["onehot"], db_pad, None))
        dnb, dni, dlq = _zeros(dim, dev), _zeros(dim, dev), _zeros(dim, dev)
        dbrain, dtemb = ops.prior_tokens_bwd(dx.view(B, 3, dim), S["keep_b"], S["keep_i"], dnb, dni, dlq)
        _acc(net.null_brain_embeds, dnb)
        _acc(net.null_image_embed, dni)
        _acc(net.learned_query, dlq)
        mlp = net.to_time_embeds[0][1].net
        e0, a0, s0, a1, s1 = S["time"]
        d = L.bwd(dtemb, s1, mlp[2])
        d = L.bwd(ops.act_bwd(a1, d, ACT_SILU), s0, mlp[1][0])
        L.bwd(ops.act_bwd(a0, d, ACT_SILU), e0, mlp[0][0], want_dx=False)
        self.S = None
        return dbrain.view(B, 1, dim)


# ------------------------------------------------------------------------------------------------ soft_clip_loss
class SoftClipTrain:
    """soft_clip_loss(preds, targs, temp) (train_diffusion_prior.py:125-133) on [B, C] rows; fp32 throughout (the B x B logits
    are divided by a temperature of ~0.005, so bf16 products would not survive)."""

    def __init__(self):
        self.S = None

    def forward(self, preds, targs, temp):
        p, t = preds.reshape(len(preds), -1).contiguous().float(), targs.reshape(len(targs), -1).contiguous().float()
        pt = ops.linear(p, t, None)                                                          # preds @ targs.T
        tt = ops.linear(t, t, None)
        loss, dsim = ops.soft_clip_loss_grad(pt, tt, float(temp))
        self.S = (dsim, t)
        return loss

    def backward(self, gloss=None):
        dsim, t = self.S
        # pt = p t^T -> dp = dsim t : a GEMM whose weight operand is t^T ([C, B] row-major)
        dp = ops.linear(dsim, ops.transpose_cast(t, torch.float32, R_pad=t.shape[0]), None)
        self.S = None
        return dp if gloss is None else dp * gloss


# ------------------------------------------------------------------------------------------------ autograd glue (drop-in surface)
def _params(mod):
    return [p for p in mod.parameters() if p.requires_grad]


class _BrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, step, voxel, masks, *params):
        o, proj = step.forward(voxel, masks)
        ctx.step = step
        ctx.set_materialize_grads(False)
        return o, proj

    @staticmethod
    def backward(ctx, d_o, d_proj):
        ps = _params(ctx.step.net)
        old = [p.grad for p in ps]
        for p in ps:
            p.grad = None
        ctx.step.backward(d_o, d_proj)
        grads = [p.grad for p in ps]
        for p, g in zip(ps, old):
            p.grad = g
        return (None, None, None, *grads)


class _PriorLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, step, text_embed, image_embed, kw, *params):
        loss, pred = step.forward(text_embed, image_embed, **kw)
        ctx.step, ctx.tshape = step, text_embed.shape
        ctx.set_materialize_grads(False)
        return loss.float().reshape(()).clone(), pred.clone()      # fresh tensors: the reference divides pred in place (:450)

    @staticmethod
    def backward(ctx, gloss, gpred):
        ps = _params(ctx.step.prior.net)
        old = [p.grad for p in ps]
        for p in ps:
            p.grad = None
        if gloss is None:
            gloss = torch.zeros((), dtype=torch.float32, device=gpred.device)
        dtext = ctx.step.backward(gloss.float(), gpred)
        grads = [p.grad for p in ps]
        for p, g in zip(ps, old):
            p.grad = g
        return (None, dtext.view(ctx.tshape), None, None, *grads)


class _SoftClipFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, preds, targs, temp):
        step = SoftClipTrain()
        loss = step.forward(preds, targs, temp)
        ctx.step, ctx.shape = step, preds.shape
        return loss.float().reshape(())

    @staticmethod
    def backward(ctx, gloss):
        return ctx.step.backward(gloss.float()).view(ctx.shape), None, None


def brain_forward_train(net, x, dropout_masks=None):
    """BrainNetwork.forward with gradient tracking (called by BrainNetwork.forward when autograd is recording)."""
    step = BrainTrain(net, net.precision)
    return _BrainFn.apply(step, x, dropout_masks, *_params(net))


def prior_loss(prior, text_embed, image_embed, **kw):
    """InstructDiffusionPrior.forward(text_embed=..., image_embed=...) -> (loss, pred) with gradient tracking."""
    step = PriorLossTrain(prior, getattr(prior, "precision", None) or prior.voxel2clip.precision)
    return _PriorLossFn.apply(step, text_embed, image_embed, kw, *_params(prior.net))


def soft_clip_loss(preds, targs, temp=0.125):
    return _SoftClipFn.apply(preds, targs, float(temp))


# ------------------------------------------------------------------------------------------------ optimiser + the whole iteration
class PriorAdamW:
    """torch.optim.AdamW over the four parameter groups of train_diffusion_prior.py:996-1004: weight decay 1e-2 except for
    parameters whose NAME contains 'bias', 'LayerNorm.bias' or 'LayerNorm.weight' (the reference's substring rule, which also
    exempts rel_pos_bias.relative_attention_bias.weight and does not exempt the LayerNorm gains)."""

    NO_DECAY = ("bias", "LayerNorm.bias", "LayerNorm.weight")

    def __init__(self, prior, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.lr, self.betas, self.eps, self.t = lr, betas, eps, 0
        self.items = []
        for mod in (prior.net, prior.voxel2clip):
            for n, p in mod.named_parameters():
                if not p.requires_grad:
                    continue
                wd = 0.0 if any(nd in n for nd in self.NO_DECAY) else weight_decay
                self.items.append((p, wd, torch.zeros_like(p, dtype=torch.float32), torch.zeros_like(p, dtype=torch.float32)))

    def zero_grad(self, set_to_none=True):
        for p, *_ in self.items:
            p.grad = None

    @torch.no_grad()
    def step(self, lr=None):
        """One update of every parameter that has a gradient: ONE launch over a device-resident pointer table (rebuilt only when a
        gradient tensor moved: a graph-replayed step writes its gradients in place, so the table is built once)."""
        self.t += 1
        lr = self.lr if lr is None else lr
        live = [(p.data, p.grad if p.grad.is_contiguous() else p.grad.contiguous(), m, v, wd) for p, wd, m, v in self.items if p.grad is not None]
        if not live:
            return
        key = tuple((p.data_ptr(), g.data_ptr()) for p, g, *_ in live)
        if getattr(self, "_table_key", None) != key:
            self._table, self._table_key, self._n = ops.adamw_table(live), key, len(live)
        self._keep = live                          # the table holds raw pointers: keep the tensors alive until the launch is queued
        ops.adamw_multi(self._table, self._n, lr, self.betas[0], self.betas[1], self.eps, self.t)
        ops.WEIGHT_EPOCH += 1


class PriorTrainStep:
    """One iteration of train_diffusion_prior.py:434-486 without autograd: voxel2clip -> prior loss + soft_clip_loss ->
    backward of both into voxel2clip -> (optionally) AdamW. Returns device scalars: loss_nce and prior_mult * loss_prior
    (``loss_prior_scaled``; the total loss of :474 is their sum), and the prediction."""

    def __init__(self, prior, precision=None, prior_mult=30.0):
        self.prior = prior
        self.precision = precision or prior.voxel2clip.precision
        self.prior_mult = prior_mult
        self.brain = BrainTrain(prior.voxel2clip, self.precision)
        self.net = PriorLossTrain(prior, self.precision)
        self.clip = SoftClipTrain()

    @torch.no_grad()
    def __call__(self, voxel, clip_target, temp, *, times=None, noise=None, keep_brain=None, keep_image=None, dropout_masks=None,
                 generator=None, optimizer=None):
        prior = self.prior
        B = voxel.shape[0]
        clip_voxels, proj = self.brain.forward(voxel, dropout_masks)                                      # :441
        x_start = ops.scale_f32(clip_target, float(prior.image_embed_scale))                              # :453
        loss_prior, pred = self.net.forward(clip_voxels.view(B, -1, prior.net.dim), x_start, times=times, noise=noise, keep_brain=keep_brain,
                                            keep_image=keep_image, generator=generator, loss_scale=self.prior_mult)       # :449, :474
        pn, pnorm = ops.rows_stat_div(proj.reshape(B, -1), 1)                                             # normalize(proj.flatten(1)) :455
        tn, _ = ops.rows_stat_div(clip_target.reshape(B, -1).float(), 1)
        loss_nce = self.clip.forward(pn, tn, temp)                                                        # :465-468
        d_text = self.net.backward()                                                                      # loss = nce + mult * prior :474
        d_pn = self.clip.backward()
        d_proj = ops.rows_stat_div_bwd(pn, d_pn, pnorm, 1)
        self.brain.backward(d_text.reshape(B, -1), d_proj)
        if optimizer is not None:
            optimizer.step()
        return dict(loss_prior_scaled=loss_prior, prior_mult=self.prior_mult, loss_nce=loss_nce, pred=pred)


class GraphedPriorTrainStep:
    """PriorTrainStep (forward + backward of the whole iteration, ~600 short launches) captured ONCE per (batch, temperature) in a
    CUDA graph; every stochastic draw is made with torch's generator OUTSIDE the graph and copied into static buffers, the graph
    writes the gradients in place, PriorAdamW then updates all parameters in one launch:

        gstep, opt = GraphedPriorTrainStep(prior, batch=256), PriorAdamW(prior)
        out = gstep(voxel, clip_target, temp); opt.step()
    """

    def __init__(self, prior, batch, precision=None, prior_mult=30.0, warmup=2, dropout=True):
        self.prior, self.B, self.warmup, self.dropout = prior, batch, warmup, dropout
        self.step = PriorTrainStep(prior, precision=precision, prior_mult=prior_mult)
        dev = next(prior.parameters()).device
        v2c = prior.voxel2clip
        h, dim = v2c.lin0[0].out_features, prior.net.dim
        f = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)  # noqa: E731
        self.buf = dict(voxel=f(batch, v2c.lin0[0].in_features), target=f(batch, 1, dim), noise=f(batch, 1, dim), keep_b=f(batch) + 1,
                        keep_i=f(batch) + 1, times=torch.zeros((batch,), dtype=torch.long, device=dev),
                        masks=[f(batch, h) + 1 for _ in range(1 + v2c.n_blocks)] if dropout else None)
        self.drop_p = [v2c.lin0[3].p] + [m[3].p for m in v2c.mlp]
        self.graph, self.key, self.out = None, None, None

    def _run(self, temp):
        b = self.buf
        return self.step(b["voxel"], b["target"], temp, times=b["times"], noise=b["noise"], keep_brain=b["keep_b"], keep_image=b["keep_i"],
                         dropout_masks=b["masks"])

    def _capture(self, temp):
        params = [p for p in self.prior.parameters() if p.requires_grad]
        side = torch.cuda.Stream(device=self.buf["voxel"].device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                for p in params:
                    p.grad = None
                self._run(temp)
        torch.cuda.current_stream().wait_stream(side)
        for p in params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._run(temp)
        self.grads = [p.grad for p in params]       # static tensors of the graph's pool, rewritten by every replay
        self.key = (float(temp),)

    @torch.no_grad()
    def __call__(self, voxel, clip_target, temp, *, times=None, noise=None, keep_brain=None, keep_image=None, dropout_masks=None,
                 generator=None):
        prior, b, B = self.prior, self.buf, self.B
        dev = b["voxel"].device
        b["voxel"].copy_(voxel.reshape(B, -1))
        b["target"].copy_(clip_target.reshape(B, 1, -1))
        T = prior.noise_scheduler.num_timesteps
        b["times"].copy_(times if times is not None else torch.randint(0, T, (B,), device=dev, generator=generator))
        b["noise"].copy_(noise.reshape(b["noise"].shape) if noise is not None else torch.randn(b["noise"].shape, device=dev, generator=generator))
        kb = keep_brain if keep_brain is not None else torch.rand((B,), device=dev, generator=generator) < (1.0 - prior.text_cond_drop_prob)
        ki = keep_image if keep_image is not None else torch.rand((B,), device=dev, generator=generator) < (1.0 - prior.image_cond_drop_prob)
        b["keep_b"].copy_(kb.reshape(B).float())
        b["keep_i"].copy_(ki.reshape(B).float())
        if self.dropout:
            for k, (m, pd) in enumerate(zip(b["masks"], self.drop_p)):
                if dropout_masks is not None:
                    m.copy_(dropout_masks[k])
                else:                                # Bernoulli(1 - p) / (1 - p), drawn by torch (data generation, not model arithmetic)
                    m.copy_((torch.rand(m.shape, device=dev, generator=generator) >= pd).float() / (1.0 - pd))
        if self.graph is None or self.key[0] != float(temp):
            self._capture(temp)
        self.graph.replay()
        for p, g in zip([p for p in prior.parameters() if p.requires_grad], self.grads):
            p.grad = g
        return self.out
