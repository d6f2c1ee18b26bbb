"""CPU: the diffusion-prior oracle against golden vectors minted from the REFERENCE'S OWN classes (models/diffusion_prior.py
BrainNetwork / FlaggedCausalTransformer / VersatileDiffusionPriorNetwork / InstructDiffusionPrior, executed by
oracle/make_golden.py over oracle/dalle2_standin.py because dalle2_pytorch is not vendored), plus drop-in key parity."""
import numpy as np
import torch

from avi_talking_b200 import synth
from oracle import prior_oracle as po


def test_brain_network_matches_reference(golden):
    g = golden("prior")
    sd, inp = synth.prior_state(), synth.prior_inputs(4, 100)
    x, proj = po.brain_network(sd, inp["voxel"])
    assert np.abs(x.numpy() - g["brain_x"]).max() < 1e-5
    assert np.abs(proj.numpy() - g["brain_proj"]).max() < 1e-5


def test_prior_network_and_samplers_match_reference(golden):
    g = golden("prior")
    sd, inp = synth.prior_state(), synth.prior_inputs(4, 100)
    text = torch.from_numpy(g["brain_x"]).view(4, -1, 128)
    o = po.prior_net_forward(sd, inp["image_embed"], torch.full((4,), 37), text)
    assert np.abs(o.numpy() - g["net_t37"]).max() < 1e-5
    y = po.p_sample_loop(sd, text, inp["image_embed"], inp["noises"], timesteps=100)
    assert np.abs(y.numpy() - g["ddpm100"]).max() < 1e-5
    y = po.p_sample_loop(sd, text, inp["image_embed"], inp["noises"], timesteps=64)
    assert np.abs(y.numpy() - g["ddim64"]).max() < 1e-5


def test_noise_schedule(golden):
    g = golden("prior")
    ns = po.noise_schedule(100)
    for k in ("betas", "alphas_cumprod_prev", "posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped"):
        assert np.abs(ns[k].numpy() - g["sched_" + k]).max() == 0.0
    ac = ns["alphas_cumprod"]
    assert bool((ac[1:] < ac[:-1]).all()) and 0 < ac[-1] < 1e-3 + ac[-2]
    # q_posterior coefficients reproduce x_{t-1} = x0 when x_t = sqrt(ac_t) x0 has no noise ... at t = 0 the posterior is x0 itself
    assert abs(ns["posterior_mean_coef1"][0].item() - 1.0) < 1e-6 and abs(ns["posterior_mean_coef2"][0].item()) < 1e-6
    pairs = po.ddim_time_pairs(100, 64)
    assert len(pairs) == 63 and pairs[0] == (98, 96) and pairs[-1] == (0, -1)
    assert all(a > b for a, b in pairs)


def test_dropin_state_dict_keys_match_reference(golden):
    from avi_talking_b200.diffusion_prior import BrainNetwork, InstructDiffusionPrior, VersatileDiffusionPriorNetwork
    g = golden("prior")
    brain = BrainNetwork(in_dim=768, out_dim=128, clip_size=128, use_projector=True)
    net = VersatileDiffusionPriorNetwork(dim=128, depth=6, dim_head=64, heads=8, causal=False, num_tokens=1, learned_query_mode="pos_emb")
    prior = InstructDiffusionPrior(net=net, image_embed_dim=128, condition_on_text_encodings=False, timesteps=100, cond_drop_prob=0.2,
                                   image_embed_scale=None, voxel2clip=brain)
    own = sorted(k for k in prior.state_dict() if not k.startswith("noise_scheduler."))
    assert own == list(g["state_keys"])
    missing, unexpected = prior.load_state_dict(synth.prior_state(), strict=False)
    assert not unexpected and all(k.startswith("noise_scheduler.") for k in missing)
    assert abs(prior.image_embed_scale - 128 ** 0.5) < 1e-12
    ns = po.noise_schedule(100)
    for k in ("betas", "posterior_mean_coef1", "sqrt_recipm1_alphas_cumprod"):
        assert torch.equal(getattr(prior.noise_scheduler, k), ns[k])


def test_emote_oracle_matches_reference_golden(golden):
    """Path B: the EMOTE oracle against the reference's own TalkingHeadBase.forward (oracle/make_golden.py golden_emote)."""
    from oracle import emote_oracle as eo
    g = golden("emote")
    sd, w, buf = synth.emote_state(), synth.wav2vec2_state(0), synth.flame_buffers(300, 50)
    o = eo.talking_head_forward(sd, w, buf, synth.emote_sample(1, 27, seed=50))
    assert np.abs(o["predicted_exp"].numpy() - g["t27_predicted_exp"]).max() < 1e-5
    assert np.abs(o["predicted_jaw"].numpy() - g["t27_predicted_jaw"]).max() < 1e-5
    assert np.abs(o["prior_input_sequence"].numpy() - g["t27_prior_input_sequence"]).max() < 1e-5
    assert np.abs(o["predicted_vertices"].numpy()[:, :, ::7] - g["t27_predicted_vertices_sub"]).max() < 1e-6
    assert np.abs(o["gt_vertices"].numpy()[:, :, ::7] - g["t27_gt_vertices_sub"]).max() < 1e-6
    assert np.abs(eo.alibi_future_mask(8, 40).numpy() - g["alibi_future_8_40"]).max() == 0.0
    style = torch.from_numpy(np.random.default_rng(60).normal(0, 0.5, size=(1, 1, 128)).astype(np.float32))
    o = eo.talking_head_forward(sd, w, buf, synth.emote_sample(1, 27, seed=50), style_emb=style, is_external_style_emb=True)
    assert np.abs(o["predicted_exp"].numpy() - g["ext_predicted_exp"]).max() < 1e-5


def test_emote_dropin_state_dict_keys_match_reference(golden):
    from transformers import Wav2Vec2Config

    from avi_talking_b200.flame import FLAME
    from avi_talking_b200.talking_head import TalkingHeadWrapper, emote_cfg
    from avi_talking_b200.wav2vec import Wav2Vec2Model
    g = golden("emote")
    fcfg = synth.write_flame_assets("/tmp/avi_flame_assets_keys")
    fcfg.n_shape, fcfg.n_exp = 300, 50
    m = TalkingHeadWrapper.from_parts(Wav2Vec2Model(Wav2Vec2Config()), FLAME(fcfg), emote_cfg(n_identities=32))
    own = sorted(k for k in m.talking_head_model.state_dict() if k.startswith("sequence_") and ".flame." not in k)
    assert own == list(g["state_keys"])


def test_classifier_free_guidance_matches_reference(golden):
    """cond_scale = 2.5 through the reference's own forward_with_cond_scale / p_sample_loop (oracle/make_golden.golden_prior_cfg)."""
    g, gp = golden("prior_cfg"), golden("prior")
    cs = float(g["cond_scale"])
    sd, inp = synth.prior_state(), synth.prior_inputs(4, 100)
    text = torch.from_numpy(gp["brain_x"]).view(4, -1, 128)
    o = po.forward_with_cond_scale(sd, inp["image_embed"], torch.full((4,), 37), text, cs)
    assert np.abs(o.numpy() - g["net_t37_cfg"]).max() < 2e-5
    y = po.p_sample_loop(sd, text, inp["image_embed"], inp["noises"], timesteps=100, cond_scale=cs)
    assert np.abs(y.numpy() - g["ddpm100_cfg"]).max() < 2e-5
    y = po.p_sample_loop(sd, text, inp["image_embed"], inp["noises"], timesteps=64, cond_scale=cs)
    assert np.abs(y.numpy() - g["ddim64_cfg"]).max() < 2e-5
    assert np.abs(g["ddim64_cfg"] - gp["ddim64"]).max() > 1e-2      # guidance changes the result


def test_dalle2_standin_pieces_against_independent_public_implementations():
    """dalle2_pytorch / rotary_embedding_torch are un-vendored and absent from this image, so oracle/dalle2_standin.py stays UNPINNED as a
    whole. Two of its pieces do have independent public implementations installed here, and are pinned on them: the T5 relative-position
    bucketing behind RelPosBias (transformers' T5Attention._relative_position_bucket, causal form) and the interleaved-pair rotary
    embedding (transformers' GPT-J apply_rotary_pos_emb, the same convention as rotary_embedding_torch); LayerNorm (no bias) against
    torch.nn.functional.layer_norm and SwiGLU against its definition."""
    import torch.nn.functional as F
    from transformers.models.gptj import modeling_gptj as gptj
    from transformers.models.t5.modeling_t5 import T5Attention
    from oracle import dalle2_standin as ds
    rel = torch.arange(-300, 40)[None, :] - torch.zeros(1, 1, dtype=torch.long)
    for nb, md in ((32, 128), (16, 64)):
        got = ds.RelPosBias._relative_position_bucket(rel, num_buckets=nb, max_distance=md)
        want = T5Attention._relative_position_bucket(rel, bidirectional=False, num_buckets=nb, max_distance=md)
        assert torch.equal(got, want)
    torch.manual_seed(0)
    B, H, T, D = 2, 3, 7, 32
    x = torch.randn(B, H, T, D)
    rot = ds.RotaryEmbedding(dim=D)
    got = rot.rotate_queries_or_keys(x)                                  # [B, H, T, D], positions along dim -2
    sincos = gptj.create_sinusoidal_positions(T, D)                      # [T, D]: sin | cos halves
    sin, cos = sincos[None, :, : D // 2], sincos[None, :, D // 2:]
    want = gptj.apply_rotary_pos_emb(x.transpose(1, 2), sin, cos).transpose(1, 2)   # GPT-J layout is [B, T, H, D]
    assert (got - want).abs().max().item() < 1e-6
    ln = ds.LayerNorm(24)
    with torch.no_grad():
        ln.g.copy_(torch.randn(24))
    y = torch.randn(5, 24)
    assert (ln(y) - F.layer_norm(y, (24,), ln.g, None, 1e-5)).abs().max().item() < 1e-6
    z = torch.randn(4, 10)
    assert torch.equal(ds.SwiGLU()(z), z[:, :5] * F.silu(z[:, 5:]))
    betas = ds.cosine_beta_schedule(1000)                                # Nichol & Dhariwal 2021, eq. 17 restated directly
    t = torch.arange(1001, dtype=torch.float64) / 1000
    f = torch.cos((t + 0.008) / 1.008 * torch.pi / 2) ** 2
    assert (betas - torch.clip(1 - f[1:] / f[:-1], 0, 0.999)).abs().max().item() < 1e-12
