"""Stand-in for the un-vendored ``dalle2_pytorch`` / ``rotary_embedding_torch`` names that models/diffusion_prior.py imports
(:12-18) - TEST INFRASTRUCTURE (see oracle/__init__.py).

Purpose: let ``oracle/make_golden.py`` execute the REFERENCE'S OWN class sources (FlaggedCausalTransformer,
VersatileDiffusionPriorNetwork, InstructDiffusionPrior, BrainNetwork) in the build container, where the real packages are not
installed. The reference's own arithmetic (token assembly, transformer loop, p_sample, p_sample_loop_ddpm) is then pinned by
golden vectors; what stays UNPINNED is exactly this file: a restatement of the published upstream modules (lucidrains
dalle2_pytorch v1.x: LayerNorm, RelPosBias, Attention, FeedForward/SwiGLU, SinusoidalPosEmb, MLP, NoiseScheduler,
DiffusionPrior.{p_mean_variance, p_sample_loop, p_sample_loop_ddim}; rotary_embedding_torch.RotaryEmbedding), written as
nn.Modules with the upstream attribute / parameter names. Pieces with an independent public implementation in this image are pinned
on it (tests/test_oracle_prior.py::test_dalle2_standin_pieces_against_independent_public_implementations: RelPosBias bucketing vs
transformers' T5, the rotary embedding vs transformers' GPT-J, the cosine schedule vs its published formula); the rest is unpinned.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from einops import rearrange, repeat
from einops.layers.torch import Rearrange  # noqa: F401  (re-exported for the reference source)
from torch import einsum


def exists(v):
    return v is not None


def default(v, d):
    if exists(v):
        return v
    return d() if callable(d) else d


def l2norm(t):
    return F.normalize(t, dim=-1)


def prob_mask_like(shape, prob, device):
    if prob == 1:
        return torch.ones(shape, device=device, dtype=torch.bool)
    if prob == 0:
        return torch.zeros(shape, device=device, dtype=torch.bool)
    return torch.zeros(shape, device=device).float().uniform_(0, 1) < prob


class LayerNorm(nn.Module):
    def __init__(self, dim, eps=1e-5, fp16_eps=1e-3, stable=False):
        super().__init__()
        self.eps, self.fp16_eps, self.stable = eps, fp16_eps, stable
        self.g = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        eps = self.eps if x.dtype == torch.float32 else self.fp16_eps
        if self.stable:
            x = x / x.amax(dim=-1, keepdim=True).detach()
        var = torch.var(x, dim=-1, unbiased=False, keepdim=True)
        mean = torch.mean(x, dim=-1, keepdim=True)
        return (x - mean) * (var + eps).rsqrt() * self.g


class RelPosBias(nn.Module):
    def __init__(self, causal=False, num_buckets=32, max_distance=128, heads=8):
        super().__init__()
        self.num_buckets, self.max_distance = num_buckets, max_distance
        self.relative_attention_bias = nn.Embedding(num_buckets, heads)

    @staticmethod
    def _relative_position_bucket(relative_position, num_buckets=32, max_distance=128):
        n = -relative_position
        n = torch.max(n, torch.zeros_like(n))
        max_exact = num_buckets // 2
        is_small = n < max_exact
        val_if_large = max_exact + (torch.log(n.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)).long()
        val_if_large = torch.min(val_if_large, torch.full_like(val_if_large, num_buckets - 1))
        return torch.where(is_small, n, val_if_large)

    def forward(self, i, j, *, device):
        q_pos = torch.arange(i, dtype=torch.long, device=device)
        k_pos = torch.arange(j, dtype=torch.long, device=device)
        rel_pos = rearrange(k_pos, "j -> 1 j") - rearrange(q_pos, "i -> i 1")
        rp_bucket = self._relative_position_bucket(rel_pos, num_buckets=self.num_buckets, max_distance=self.max_distance)
        values = self.relative_attention_bias(rp_bucket)
        return rearrange(values, "i j h -> h i j")


class RotaryEmbedding(nn.Module):
    """rotary_embedding_torch.RotaryEmbedding(dim): freqs = 1 / theta^(2i/dim), interleaved-pair rotation of the first `dim` features."""

    def __init__(self, dim, theta=10000):
        super().__init__()
        self.freqs = nn.Parameter(1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim)), requires_grad=False)

    def rotate_queries_or_keys(self, t, seq_dim=-2):
        seq_len = t.shape[seq_dim]
        pos = torch.arange(seq_len, device=t.device).type_as(self.freqs)
        freqs = einsum("..., f -> ... f", pos, self.freqs)
        freqs = repeat(freqs, "... n -> ... (n r)", r=2)
        rot_dim = freqs.shape[-1]
        t_left, t_mid, t_right = t[..., :0], t[..., :rot_dim], t[..., rot_dim:]
        x = rearrange(t_mid, "... (d r) -> ... d r", r=2)
        x1, x2 = x.unbind(dim=-1)
        rot_half = rearrange(torch.stack((-x2, x1), dim=-1), "... d r -> ... (d r)")
        t_mid = (t_mid * freqs.cos()) + (rot_half * freqs.sin())
        return torch.cat((t_left, t_mid, t_right), dim=-1)


class SwiGLU(nn.Module):
    def forward(self, x):
        x, gate = x.chunk(2, dim=-1)
        return x * F.silu(gate)


def FeedForward(dim, mult=4, dropout=0.0, post_activation_norm=False):
    inner_dim = int(mult * dim)
    return nn.Sequential(
        LayerNorm(dim),
        nn.Linear(dim, inner_dim * 2, bias=False),
        SwiGLU(),
        LayerNorm(inner_dim) if post_activation_norm else nn.Identity(),
        nn.Dropout(dropout),
        nn.Linear(inner_dim, dim, bias=False),
    )


class Attention(nn.Module):
    def __init__(self, dim, *, dim_head=64, heads=8, dropout=0.0, causal=False, rotary_emb=None, cosine_sim=True, cosine_sim_scale=16):
        super().__init__()
        self.scale = cosine_sim_scale if cosine_sim else (dim_head ** -0.5)
        self.cosine_sim = cosine_sim
        self.heads = heads
        inner_dim = dim_head * heads
        self.causal = causal
        self.norm = LayerNorm(dim)
        self.dropout = nn.Dropout(dropout)
        self.null_kv = nn.Parameter(torch.randn(2, dim_head))
        self.to_q = nn.Linear(dim, inner_dim, bias=False)
        self.to_kv = nn.Linear(dim, dim_head * 2, bias=False)
        self.rotary_emb = rotary_emb
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim, bias=False), LayerNorm(dim))

    def forward(self, x, mask=None, attn_bias=None):
        b, n, device = *x.shape[:2], x.device
        x = self.norm(x)
        q, k, v = (self.to_q(x), *self.to_kv(x).chunk(2, dim=-1))
        q = rearrange(q, "b n (h d) -> b h n d", h=self.heads)
        q = q * self.scale
        if exists(self.rotary_emb):
            q, k = map(self.rotary_emb.rotate_queries_or_keys, (q, k))
        nk, nv = map(lambda t: repeat(t, "d -> b 1 d", b=b), self.null_kv.unbind(dim=-2))
        k = torch.cat((nk, k), dim=-2)
        v = torch.cat((nv, v), dim=-2)
        if self.cosine_sim:
            q, k = map(l2norm, (q, k))
        q, k = map(lambda t: t * math.sqrt(self.scale), (q, k))
        sim = einsum("b h i d, b j d -> b h i j", q, k)
        if exists(attn_bias):
            sim = sim + attn_bias
        max_neg_value = -torch.finfo(sim.dtype).max
        if exists(mask):
            mask = F.pad(mask, (1, 0), value=True)
            mask = rearrange(mask, "b j -> b 1 1 j")
            sim = sim.masked_fill(~mask, max_neg_value)
        if self.causal:
            i, j = sim.shape[-2:]
            causal_mask = torch.ones((i, j), dtype=torch.bool, device=device).triu(j - i + 1)
            sim = sim.masked_fill(causal_mask, max_neg_value)
        attn = sim.softmax(dim=-1, dtype=torch.float32)
        attn = attn.type(sim.dtype)
        attn = self.dropout(attn)
        out = einsum("b h i j, b j d -> b h i d", attn, v)
        out = rearrange(out, "b h n d -> b n (h d)")
        return self.to_out(out)


class CausalTransformer(nn.Module):  # imported by the reference but unused by it
    pass


class SinusoidalPosEmb(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        dtype, device = x.dtype, x.device
        half_dim = self.dim // 2
        emb = math.log(10000) / (half_dim - 1)
        emb = torch.exp(torch.arange(half_dim, device=device, dtype=dtype) * -emb)
        emb = rearrange(x, "i -> i 1") * rearrange(emb, "j -> 1 j")
        return torch.cat((emb.sin(), emb.cos()), dim=-1).type(dtype)


class MLP(nn.Module):
    def __init__(self, dim_in, dim_out, *, expansion_factor=2.0, depth=2, norm=False):
        super().__init__()
        hidden_dim = int(expansion_factor * dim_out)
        norm_fn = lambda: nn.LayerNorm(hidden_dim) if norm else nn.Identity()  # noqa: E731
        layers = [nn.Sequential(nn.Linear(dim_in, hidden_dim), nn.SiLU(), norm_fn())]
        for _ in range(depth - 1):
            layers.append(nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nn.SiLU(), norm_fn()))
        layers.append(nn.Linear(hidden_dim, dim_out))
        self.net = nn.Sequential(*layers)

    def forward(self, x):
        return self.net(x.float())


def extract(a, t, x_shape):
    b, *_ = t.shape
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


def cosine_beta_schedule(timesteps, s=0.008):
    steps = timesteps + 1
    x = torch.linspace(0, timesteps, steps, dtype=torch.float64)
    alphas_cumprod = torch.cos(((x / timesteps) + s) / (1 + s) * torch.pi * 0.5) ** 2
    alphas_cumprod = alphas_cumprod / alphas_cumprod[0]
    betas = 1 - (alphas_cumprod[1:] / alphas_cumprod[:-1])
    return torch.clip(betas, 0, 0.999)


class NoiseScheduler(nn.Module):
    def __init__(self, *, beta_schedule, timesteps, loss_type, p2_loss_weight_gamma=0.0, p2_loss_weight_k=1):
        super().__init__()
        assert beta_schedule == "cosine"
        betas = cosine_beta_schedule(timesteps)
        alphas = 1.0 - betas
        alphas_cumprod = torch.cumprod(alphas, axis=0)
        alphas_cumprod_prev = F.pad(alphas_cumprod[:-1], (1, 0), value=1.0)
        (timesteps,) = betas.shape
        self.num_timesteps = int(timesteps)
        self.loss_fn = F.mse_loss
        register_buffer = lambda name, val: self.register_buffer(name, val.to(torch.float32))  # noqa: E731
        register_buffer("betas", betas)
        register_buffer("alphas_cumprod", alphas_cumprod)
        register_buffer("alphas_cumprod_prev", alphas_cumprod_prev)
        register_buffer("sqrt_alphas_cumprod", torch.sqrt(alphas_cumprod))
        register_buffer("sqrt_one_minus_alphas_cumprod", torch.sqrt(1.0 - alphas_cumprod))
        register_buffer("log_one_minus_alphas_cumprod", torch.log(1.0 - alphas_cumprod))
        register_buffer("sqrt_recip_alphas_cumprod", torch.sqrt(1.0 / alphas_cumprod))
        register_buffer("sqrt_recipm1_alphas_cumprod", torch.sqrt(1.0 / alphas_cumprod - 1))
        posterior_variance = betas * (1.0 - alphas_cumprod_prev) / (1.0 - alphas_cumprod)
        register_buffer("posterior_variance", posterior_variance)
        register_buffer("posterior_log_variance_clipped", torch.log(posterior_variance.clamp(min=1e-20)))
        register_buffer("posterior_mean_coef1", betas * torch.sqrt(alphas_cumprod_prev) / (1.0 - alphas_cumprod))
        register_buffer("posterior_mean_coef2", (1.0 - alphas_cumprod_prev) * torch.sqrt(alphas) / (1.0 - alphas_cumprod))

    def sample_random_times(self, batch):
        return torch.randint(0, self.num_timesteps, (batch,), device=self.betas.device, dtype=torch.long)

    def q_sample(self, x_start, t, noise=None):
        noise = default(noise, lambda: torch.randn_like(x_start))
        return extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start + extract(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise

    def q_posterior(self, x_start, x_t, t):
        posterior_mean = extract(self.posterior_mean_coef1, t, x_t.shape) * x_start + extract(self.posterior_mean_coef2, t, x_t.shape) * x_t
        posterior_variance = extract(self.posterior_variance, t, x_t.shape)
        posterior_log_variance_clipped = extract(self.posterior_log_variance_clipped, t, x_t.shape)
        return posterior_mean, posterior_variance, posterior_log_variance_clipped

    def predict_noise_from_start(self, x_t, t, x0):
        return (extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t - x0) / extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape)


class DiffusionPrior(nn.Module):
    """The slice of dalle2_pytorch.DiffusionPrior the reference's subclass relies on at inference."""

    def __init__(self, net, *, clip=None, image_embed_dim=None, image_size=None, image_channels=3, timesteps=1000, sample_timesteps=None,
                 cond_drop_prob=0.0, text_cond_drop_prob=None, image_cond_drop_prob=None, loss_type="l2", predict_x_start=True,
                 predict_v=False, beta_schedule="cosine", condition_on_text_encodings=True, sampling_clamp_l2norm=False,
                 sampling_final_clamp_l2norm=False, training_clamp_l2norm=False, init_image_embed_l2norm=False, image_embed_scale=None,
                 clip_adapter_overrides=dict()):
        super().__init__()
        self.sample_timesteps = sample_timesteps
        self.noise_scheduler = NoiseScheduler(beta_schedule=beta_schedule, timesteps=timesteps, loss_type=loss_type)
        self.clip = None
        self.net = net
        self.image_embed_dim = image_embed_dim
        self.condition_on_text_encodings = condition_on_text_encodings
        self.text_cond_drop_prob = default(text_cond_drop_prob, cond_drop_prob)
        self.image_cond_drop_prob = default(image_cond_drop_prob, cond_drop_prob)
        self.can_classifier_guidance = self.text_cond_drop_prob > 0.0 and self.image_cond_drop_prob > 0.0
        self.predict_x_start = predict_x_start
        self.predict_v = predict_v
        self.image_embed_scale = default(image_embed_scale, self.image_embed_dim ** 0.5)
        self.sampling_clamp_l2norm = sampling_clamp_l2norm
        self.sampling_final_clamp_l2norm = sampling_final_clamp_l2norm
        self.training_clamp_l2norm = training_clamp_l2norm
        self.init_image_embed_l2norm = init_image_embed_l2norm
        self.register_buffer("_dummy", torch.tensor([True]), persistent=False)

    @property
    def device(self):
        return self._dummy.device

    def l2norm_clamp_embed(self, image_embed):
        return l2norm(image_embed) * self.image_embed_scale

    def p_mean_variance(self, x, t, text_cond, self_cond=None, clip_denoised=False, cond_scale=1.0):
        assert not (cond_scale != 1.0 and not self.can_classifier_guidance)
        pred = self.net.forward_with_cond_scale(x, t, cond_scale=cond_scale, self_cond=self_cond, **text_cond)
        assert self.predict_x_start and not self.predict_v
        x_start = pred
        if self.predict_x_start and self.sampling_clamp_l2norm:
            x_start = l2norm(x_start) * self.image_embed_scale
        model_mean, posterior_variance, posterior_log_variance = self.noise_scheduler.q_posterior(x_start=x_start, x_t=x, t=t)
        return model_mean, posterior_variance, posterior_log_variance, x_start

    @torch.no_grad()
    def p_sample_loop_ddim(self, shape, text_cond, *, timesteps, eta=1.0, cond_scale=1.0, image_embed=None, noises=None):
        """Upstream draws image_embed / the per-pair noise with torch.randn; `image_embed` / `noises` inject them (test hook)."""
        batch, device, alphas, total_timesteps = shape[0], self.device, self.noise_scheduler.alphas_cumprod_prev, self.noise_scheduler.num_timesteps
        times = torch.linspace(-1.0, total_timesteps, steps=timesteps + 1)[:-1]
        times = list(reversed(times.int().tolist()))
        time_pairs = list(zip(times[:-1], times[1:]))
        if image_embed is None:
            image_embed = torch.randn(shape, device=device)
        x_start = None
        for k, (time, time_next) in enumerate(time_pairs):
            alpha = alphas[time]
            alpha_next = alphas[time_next]
            time_cond = torch.full((batch,), time, device=device, dtype=torch.long)
            self_cond = x_start if self.net.self_cond else None
            pred = self.net.forward_with_cond_scale(image_embed, time_cond, self_cond=self_cond, cond_scale=cond_scale, **text_cond)
            x_start = pred
            if self.predict_x_start and self.sampling_clamp_l2norm:
                x_start = self.l2norm_clamp_embed(x_start)
            pred_noise = self.noise_scheduler.predict_noise_from_start(image_embed, t=time_cond, x0=x_start)
            if time_next < 0:
                image_embed = x_start
                continue
            c1 = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
            c2 = ((1 - alpha_next) - torch.square(c1)).sqrt()
            noise = (noises[k] if noises is not None else torch.randn_like(image_embed)) if time_next > 0 else 0.0
            image_embed = x_start * alpha_next.sqrt() + c1 * noise + c2 * pred_noise
        if self.predict_x_start and self.sampling_final_clamp_l2norm:
            image_embed = self.l2norm_clamp_embed(image_embed)
        return image_embed

    @torch.no_grad()
    def p_sample_loop(self, *args, timesteps=None, **kwargs):
        timesteps = default(timesteps, self.noise_scheduler.num_timesteps)
        assert timesteps <= self.noise_scheduler.num_timesteps
        is_ddim = timesteps < self.noise_scheduler.num_timesteps
        if not is_ddim:
            normalized_image_embed = self.p_sample_loop_ddpm(*args, **kwargs)
        else:
            normalized_image_embed = self.p_sample_loop_ddim(*args, **kwargs, timesteps=timesteps)
        image_embed = normalized_image_embed / self.image_embed_scale
        return image_embed
