"""GPU: the diffusion-prior TRAINING step - avi_talking_b200/prior_train.py over csrc/prior_train.cu and the GEMMs -
against (a) tests/golden/prior_train_*.npz, minted from the reference's own classes + soft_clip_loss + AdamW grouping
(train_diffusion_prior.py:434-486,996-1004; models/diffusion_prior.py:369-456) over the un-pinned dalle2 stand-in, and
(b) the autograd oracle (oracle/prior_train_oracle.py) at a larger batch. Tolerances: fp32 mode 5e-4 relative on gradients
(per-tensor scale), bf16 GEMM mode 5e-2."""
import numpy as np
import pytest
import torch

from avi_talking_b200 import synth
from avi_talking_b200.smoke import build_prior
from oracle import make_golden as mg
from oracle import prior_train_oracle as pto
from helpers import check_against_golden

pytestmark = pytest.mark.gpu


def _named(prior):
    return {n: p for n, p in prior.named_parameters() if p.requires_grad}


def _cuda_inputs(inp):
    return {k: ([m.cuda() for m in v] if isinstance(v, list) else v.cuda()) for k, v in inp.items()}


@pytest.mark.parametrize("variant", ["eval", "dropout"])
@pytest.mark.parametrize("precision,gtol,ptol,ltol", [("fp32", 5e-4, 5e-6, 2e-4), ("bf16", 6e-2, 7e-4, 2e-2)])
def test_fused_iteration_matches_reference_golden(golden, variant, precision, gtol, ptol, ltol):
    from avi_talking_b200.prior_train import PriorAdamW, PriorTrainStep
    g = golden(f"prior_train_{variant}")
    inp = _cuda_inputs(mg.prior_train_inputs())
    prior = build_prior(precision).train()
    step, opt = PriorTrainStep(prior, precision=precision), PriorAdamW(prior, lr=3e-4)
    opt.zero_grad()
    out = step(inp["voxel"], inp["clip_target"], float(g["temp"]), times=inp["times"], noise=inp["noise"], keep_brain=inp["keep_brain"],
               keep_image=inp["keep_image"], dropout_masks=inp["masks"] if variant == "dropout" else None)
    loss_prior = float(out["loss_prior_scaled"]) / out["prior_mult"]
    print(f"{variant} {precision}: loss_nce {float(out['loss_nce']):.5f} (ref {float(g['loss_nce']):.5f}) loss_prior {loss_prior:.5f} "
          f"(ref {float(g['loss_prior']):.5f})")
    assert abs(float(out["loss_nce"]) - float(g["loss_nce"])) < ltol * abs(float(g["loss_nce"])) * (10 if precision == "bf16" else 1)
    assert abs(loss_prior - float(g["loss_prior"])) < ltol * float(g["loss_prior"])
    assert np.abs(out["pred"].cpu().numpy() - g["pred"]).max() < (3e-5 if precision == "fp32" else 5e-2)
    named = _named(prior)
    assert set(named) == {str(x) for x in g["names"]}
    grads = {n: p.grad for n, p in named.items()}
    assert all(v is not None for v in grads.values())
    opt.step()
    wg, wp = check_against_golden(g, {n: v.cpu() for n, v in grads.items()}, {n: p.detach().cpu() for n, p in named.items()}, gtol, ptol)
    print(f"   worst relative gradient error {wg:.2e}, worst AdamW parameter error {wp:.2e}")


def test_reference_loop_verbatim_through_autograd_equals_fused_step():
    """The iteration as train_diffusion_prior.py:441-486 writes it (voxel2clip(...), diffusion_prior(text_embed=, image_embed=),
    F.normalize, soft_clip_loss, loss.backward()) on the drop-in classes gives the gradients of the fused PriorTrainStep."""
    from avi_talking_b200.diffusion_prior import soft_clip_loss
    from avi_talking_b200.prior_train import PriorTrainStep
    inp = _cuda_inputs(mg.prior_train_inputs())
    temp = 0.0045
    prior = build_prior("fp32").train()
    prior.voxel2clip.dropout_masks = inp["masks"]
    voxel = inp["voxel"].clone().requires_grad_(True)
    clip_voxels, clip_voxels_proj = prior.voxel2clip(voxel)
    clip_voxels = clip_voxels.view(len(voxel), -1, 128)
    loss_prior, aligned = prior(text_embed=clip_voxels, image_embed=inp["clip_target"], times=inp["times"], noise=inp["noise"],
                                keep_brain=inp["keep_brain"], keep_image=inp["keep_image"])
    aligned /= prior.image_embed_scale                                                               # :450
    clip_voxels_norm = torch.nn.functional.normalize(clip_voxels_proj.flatten(1), dim=-1)
    clip_target_norm = torch.nn.functional.normalize(inp["clip_target"].flatten(1), dim=-1)
    loss_nce = soft_clip_loss(clip_voxels_norm, clip_target_norm, temp=temp)
    loss = loss_nce + 30 * loss_prior
    loss.backward()
    got = {n: p.grad.clone() for n, p in _named(prior).items()}
    for p in prior.parameters():
        p.grad = None
    out = PriorTrainStep(prior, precision="fp32")(inp["voxel"], inp["clip_target"], temp, times=inp["times"], noise=inp["noise"],
                                                  keep_brain=inp["keep_brain"], keep_image=inp["keep_image"], dropout_masks=inp["masks"])
    assert abs(float(loss) - float(out["loss_nce"]) - float(out["loss_prior_scaled"])) < 1e-4 * abs(float(loss))
    worst = 0.0
    for n, p in _named(prior).items():
        scale = float(p.grad.abs().max()) + 1e-12
        worst = max(worst, float((p.grad - got[n]).abs().max()) / scale)
    print("autograd-glued loop vs fused step: worst relative gradient difference", worst)
    assert worst < 1e-4


def test_batch_64_against_the_autograd_oracle_and_drawn_inputs():
    """A batch the golden does not cover, against oracle.train_step on the same seeded inputs; then one iteration with every draw
    left to the generator (random timesteps / noise / keep masks / dropout): finite losses, every gradient present."""
    from avi_talking_b200.prior_train import PriorAdamW, PriorTrainStep
    B = 64
    inp = mg.prior_train_inputs(B, seed=5)
    sd = synth.prior_state()
    torch.set_num_threads(8)
    want = pto.train_step(sd, inp["voxel"], inp["clip_target"], inp["times"], inp["noise"], inp["keep_brain"], inp["keep_image"], 0.006,
                          dropout_masks=inp["masks"])
    ci = _cuda_inputs(inp)
    prior = build_prior("fp32").train()
    step, opt = PriorTrainStep(prior, precision="fp32"), PriorAdamW(prior, lr=3e-4)
    out = step(ci["voxel"], ci["clip_target"], 0.006, times=ci["times"], noise=ci["noise"], keep_brain=ci["keep_brain"],
               keep_image=ci["keep_image"], dropout_masks=ci["masks"], optimizer=opt)
    assert abs(float(out["loss_nce"]) - float(want["loss_nce"])) < 2e-4 * abs(float(want["loss_nce"]))
    assert abs(float(out["loss_prior_scaled"]) / 30 - float(want["loss_prior"])) < 2e-4 * float(want["loss_prior"])
    worst = 0.0
    for n, p in prior.named_parameters():
        if p.requires_grad:
            w = want["grads"][n]
            worst = max(worst, float((p.grad.cpu() - w).abs().max()) / (float(w.abs().max()) + 1e-12))
            # AdamW moves a weight by ~lr * sign(g) on its first step: compare where the gradient is not rounding noise
            sure = w.abs() > 1e-3 * w.abs().max()
            assert float(((p.detach().cpu() - want["new"][n]).abs() * sure).max()) < 5e-6, n
    print("B=64 fp32 vs oracle: worst relative gradient error", worst)
    assert worst < 1e-3
    opt.zero_grad()
    gen = torch.Generator(device="cuda").manual_seed(1)
    out = step(ci["voxel"], ci["clip_target"], 0.006, generator=gen, optimizer=opt)
    assert torch.isfinite(out["loss_nce"]).all() and torch.isfinite(out["loss_prior_scaled"]).all()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in prior.parameters() if p.requires_grad)


def test_graph_replayed_iteration_equals_eager_and_trains():
    """GraphedPriorTrainStep (forward + backward in one CUDA graph, draws copied into static buffers) gives the eager step's
    gradients (to summation order), and 30 iterations on a fixed batch lower the prior loss."""
    from avi_talking_b200.prior_train import GraphedPriorTrainStep, PriorAdamW, PriorTrainStep
    B = 32
    inp = _cuda_inputs(mg.prior_train_inputs(B, seed=9))
    prior = build_prior("fp32").train()
    kw = dict(times=inp["times"], noise=inp["noise"], keep_brain=inp["keep_brain"], keep_image=inp["keep_image"], dropout_masks=inp["masks"])
    PriorTrainStep(prior, precision="fp32")(inp["voxel"], inp["clip_target"], 0.006, **kw)
    want = {n: p.grad.clone() for n, p in _named(prior).items()}
    for p in prior.parameters():
        p.grad = None
    gstep = GraphedPriorTrainStep(prior, B, precision="fp32")
    out = gstep(inp["voxel"], inp["clip_target"], 0.006, **kw)
    for n, p in _named(prior).items():       # LayerNorm gain / bias gradients are atomic sums: equal up to the summation order
        assert float((p.grad - want[n]).abs().max()) <= 1e-5 * float(want[n].abs().max()) + 1e-12, n
    first = float(out["loss_prior_scaled"]) / 30
    opt = PriorAdamW(prior, lr=3e-4)
    for _ in range(30):
        out = gstep(inp["voxel"], inp["clip_target"], 0.006, **kw)
        opt.step()
    last = float(out["loss_prior_scaled"]) / 30
    print(f"prior loss on a fixed batch: {first:.4f} -> {last:.4f} after 30 graph-replayed AdamW iterations")
    assert last < 0.8 * first and torch.isfinite(out["loss_nce"]).all()


def test_adamw_multi_matches_torch():
    """avi_adamw_multi (one launch over a pointer table) against torch.optim.AdamW, three steps, two decay groups."""
    from avi_talking_b200 import ops
    g = torch.Generator().manual_seed(3)
    shapes = [(257,), (64, 33), (5, 7, 3)]
    ps = [torch.randn(s, generator=g) for s in shapes]
    wds = [1e-2, 0.0, 1e-2]
    ref = [p.clone().requires_grad_(True) for p in ps]
    topt = torch.optim.AdamW([{"params": [ref[0], ref[2]], "weight_decay": 1e-2}, {"params": [ref[1]], "weight_decay": 0.0}], lr=3e-3)
    mine = [p.clone().cuda() for p in ps]
    ms, vs = [torch.zeros_like(p) for p in mine], [torch.zeros_like(p) for p in mine]
    for step in range(1, 4):
        grads = [torch.randn(s, generator=g) for s in shapes]
        for r, gr in zip(ref, grads):
            r.grad = gr.clone()
        topt.step()
        cg = [gr.cuda() for gr in grads]
        tab = ops.adamw_table([(p, gr, m, v, wd) for p, gr, m, v, wd in zip(mine, cg, ms, vs, wds)])
        ops.adamw_multi(tab, len(mine), 3e-3, 0.9, 0.999, 1e-8, step)
        torch.cuda.synchronize()
    for r, p in zip(ref, mine):
        assert (r.detach() - p.cpu()).abs().max().item() < 2e-6
