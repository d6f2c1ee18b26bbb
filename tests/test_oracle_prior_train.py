"""CPU: the prior-training oracle (oracle/prior_train_oracle.py) against tests/golden/prior_train_{eval,dropout}.npz, minted by
oracle/make_golden.golden_prior_train from the REFERENCE'S OWN BrainNetwork / VersatileDiffusionPriorNetwork /
InstructDiffusionPrior.forward + p_losses / soft_clip_loss / AdamW grouping (dalle2_pytorch underneath is the un-pinned stand-in)."""
import numpy as np
import pytest
import torch

from avi_talking_b200 import synth
from oracle import make_golden as mg
from helpers import check_against_golden
from oracle import prior_train_oracle as pto


@pytest.mark.parametrize("variant", ["eval", "dropout"])
def test_oracle_train_step_matches_reference(golden, variant):
    g = golden(f"prior_train_{variant}")
    sd, inp = synth.prior_state(), mg.prior_train_inputs()
    torch.set_num_threads(8)
    out = pto.train_step(sd, inp["voxel"], inp["clip_target"], inp["times"], inp["noise"], inp["keep_brain"], inp["keep_image"],
                         float(g["temp"]), dropout_masks=inp["masks"] if variant == "dropout" else None)
    assert abs(float(out["loss_nce"]) - float(g["loss_nce"])) < 1e-4 * abs(float(g["loss_nce"]))
    assert abs(float(out["loss_prior"]) - float(g["loss_prior"])) < 1e-5 * abs(float(g["loss_prior"]))
    # the reference returns pred / image_embed_scale? no: `aligned_clip_voxels /= image_embed_scale` happens in the caller (:450)
    assert np.abs(out["pred"].numpy() - g["pred"]).max() < 2e-5
    assert set(out["grads"]) == {str(x) for x in g["names"]}
    wg, wp = check_against_golden(g, out["grads"], out["new"], 2e-4, 2e-6)
    print(f"{variant}: worst relative gradient error {wg:.2e}, worst AdamW parameter error {wp:.2e}")


def test_adamw_grouping_follows_the_reference_substring_rule():
    """'bias' in the NAME exempts a tensor from weight decay - which also catches rel_pos_bias.relative_attention_bias.weight and
    misses every LayerNorm gain (train_diffusion_prior.py:997-1003)."""
    from avi_talking_b200.prior_train import PriorAdamW
    nd = PriorAdamW.NO_DECAY
    assert any(x in "causal_transformer.rel_pos_bias.relative_attention_bias.weight" for x in nd)
    assert not any(x in "causal_transformer.layers.0.0.norm.g" for x in nd)
    assert not any(x in "lin0.1.weight" for x in nd) and any(x in "lin0.1.bias" for x in nd)
    assert tuple(nd) == tuple(pto.NO_DECAY)
