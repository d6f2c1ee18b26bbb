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
