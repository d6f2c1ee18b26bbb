"""Timing of one diffusion-prior TRAINING iteration (train_diffusion_prior.py:434-486) at batch 256:
avi_talking_b200.prior_train.PriorTrainStep + PriorAdamW on the GPU (CUDA events, eager launches) next to the autograd oracle on
the host cores. Usage (GPU box): python profiles/prior_train_bench.py > gpurun_out/prior_train_bench.txt"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from avi_talking_b200 import _lib, synth  # noqa: E402
from avi_talking_b200.prior_train import GraphedPriorTrainStep, PriorAdamW, PriorTrainStep  # noqa: E402
from avi_talking_b200.smoke import build_prior  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator().manual_seed(0)
voxel, target = torch.randn(B, 768, generator=g), torch.randn(B, 1, 128, generator=g) * 0.5
res = {"batch": B}
for prec in ("fp32", "bf16"):
    prior = build_prior(prec).train()
    step, opt = PriorTrainStep(prior, precision=prec), PriorAdamW(prior, lr=3e-4)
    gen = torch.Generator(device="cuda").manual_seed(1)
    v, t = voxel.cuda(), target.cuda()
    for _ in range(3):
        opt.zero_grad()
        step(v, t, 0.006, generator=gen, optimizer=opt)
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 10
    for _ in range(K):
        opt.zero_grad()
        out = step(v, t, 0.006, generator=gen, optimizer=opt)
    e1.record()
    torch.cuda.synchronize()
    res[prec] = {"ms_per_iteration": e0.elapsed_time(e1) / K, "launches_per_iteration": (_lib.launch_count() - n0) / K,
                 "loss_nce": float(out["loss_nce"]), "loss_prior": float(out["loss_prior_scaled"]) / 30}
    # the same iteration replayed from one CUDA graph (draws made outside, one AdamW launch after it)
    prior = build_prior(prec).train()
    gstep, opt = GraphedPriorTrainStep(prior, B, precision=prec), PriorAdamW(prior, lr=3e-4)
    for _ in range(3):
        gstep(v, t, 0.006, generator=gen)
        opt.step()
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    e0.record()
    for _ in range(K):
        out = gstep(v, t, 0.006, generator=gen)
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    res[prec + "_graph"] = {"ms_per_iteration": e0.elapsed_time(e1) / K, "host_launches_per_iteration": (_lib.launch_count() - n0) / K,
                            "samples_per_s": B / (e0.elapsed_time(e1) / K) * 1e3, "loss_nce": float(out["loss_nce"]),
                            "loss_prior": float(out["loss_prior_scaled"]) / 30}
if "--no-cpu" not in sys.argv:
    from oracle import make_golden as mg  # checker / CPU baseline only
    from oracle import prior_train_oracle as pto
    torch.set_num_threads(os.cpu_count())
    inp = mg.prior_train_inputs(B, seed=5)
    sd = synth.prior_state()
    t0 = time.perf_counter()
    pto.train_step(sd, inp["voxel"], inp["clip_target"], inp["times"], inp["noise"], inp["keep_brain"], inp["keep_image"], 0.006,
                   dropout_masks=inp["masks"])
    res["cpu_oracle"] = {"ms_per_iteration": (time.perf_counter() - t0) * 1e3, "threads": os.cpu_count(), "kind": "port (torch autograd, fp32)"}
print(json.dumps(res))
