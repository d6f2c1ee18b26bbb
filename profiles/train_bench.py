"""BASELINE configs[4]: faceformer_vert teacher-forced training step (wav2vec2 encoder fwd+bwd with the conv extractor frozen,
decoder layer, 15069-wide vertex head, MSE x 10, Adam), bf16 GEMMs / fp32 master weights, data parallel over the GPUs of one node
with the bucketed NCCL all-reduce of train.GradBuckets. One VOCASET-like clip = 4 s of 16 kHz audio, 120 frames (30 fps).

  python profiles/train_bench.py [--clips C] [--steps K]                       # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/train_bench.py

Prints one JSON line (rank 0): clips/s and frames/s over all ranks (weak scaling: C clips per rank per step), ms per step as the max
over ranks (CUDA events), the per-phase split measured in a separate pass, (the CPU oracle's step is timed by `python bench.py --cpu-baseline train`)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from avi_talking_b200 import shard, synth, train  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=1)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--frames", type=int, default=120)
ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--graph", action="store_true", help="replay forward+backward from a CUDA graph (1 GPU)")
args = ap.parse_args()

rank, local, world = shard.env_rank_world()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

from transformers import Wav2Vec2Config  # noqa: E402

from avi_talking_b200.faceformer import FaceformerVert, make_args  # noqa: E402
from avi_talking_b200.wav2vec import Wav2Vec2Model  # noqa: E402

fd, B, T, N = 64, args.clips, args.frames, 64000
w2v = Wav2Vec2Model(Wav2Vec2Config())
w2v.load_state_dict(synth.wav2vec2_state(0), strict=False)
template = synth.flame_buffers()["v_template"].reshape(1, 1, 15069)
m = FaceformerVert(make_args(feature_dim=fd), audio_encoder=w2v, template=template)
m.load_state_dict(synth.faceformer_state(fd=fd, seed=264, variant="vert"), strict=False)
m.precision = w2v.precision = args.precision
m = m.to(dev)
opt = train.FlatAdam(m, lr=1e-4)
step = train.TrainStep(m, buckets=train.GradBuckets(m._flat_layout) if world > 1 else None)
m._train_step = step
rng = np.random.default_rng(7 + rank)
gt = (template + 1e-3 * torch.from_numpy(rng.normal(size=(B, T, 15069)).astype(np.float32))).to(dev)
audio = synth.audio(B, N, seed=500 + rank * B).to(dev)


gstep = train.GraphedTrainStep(m, audio.shape, gt.shape) if args.graph and world == 1 else None


def one_step():
    if gstep is not None:
        loss = gstep(audio, gt)
        opt.step()
        return loss
    opt.zero_grad()
    loss = m.training_loss(audio, gt)
    loss.backward()
    opt.step(grad_scale=getattr(step, "grad_divisor", 1.0))
    return loss


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(args.warmup):
    one_step()
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    loss = one_step()
e1.record()
barrier()
ms = e0.elapsed_time(e1) / args.steps
# phase split (separate pass)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
opt.zero_grad()
ev[0].record()
loss = m.training_loss(audio, gt)
ev[1].record()
loss.backward()
ev[2].record()
opt.step(grad_scale=getattr(step, "grad_divisor", 1.0))
ev[3].record()
torch.cuda.synchronize()
phases = {"forward_ms": ev[0].elapsed_time(ev[1]), "backward_allreduce_ms": ev[1].elapsed_time(ev[2]), "adam_ms": ev[2].elapsed_time(ev[3])}
ms = shard.max_over_ranks([ms], device=dev)[0]
in_sync = True
if world > 1:      # identical initial weights + averaged gradients => identical weights on every rank after any number of steps
    chk = torch.stack([m._flat_params.double().sum(), m._flat_params.double().abs().sum()])
    hi, lo = chk.clone(), chk.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    in_sync = bool(((hi - lo).abs() <= 1e-9 * hi.abs()).all())
if rank == 0:
    n_par = m._flat_layout.total
    line = {"metric": "faceformer_vert training steps/sec", "value": 1e3 / ms, "unit": "steps/s", "n_gpus": world, "ms_per_step": ms,
            "clips_per_sec": world * B * 1e3 / ms, "frames_per_sec": world * B * T * 1e3 / ms, "scaling": "weak",
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"BASELINE configs[4]: faceformer_vert teacher-forced train step, {B} clip(s) x 4 s ({T} frames) per GPU, "
                                   "fd=64, MSE x 10, Adam lr 1e-4, feature extractor frozen", "trainable_params": n_par,
                       "allreduce": "bucketed NCCL sum over the flat gradient buffer, overlapped with backward" if world > 1 else "none"},
            "phases": phases, "loss": float(loss.detach()), "cuda_graph": gstep is not None, "ranks_in_sync": in_sync,
            # Adam alone moves 7 fp32 words per parameter (read p, g, m, v; write p, m, v)
            "adam_hbm_gbs": 28.0 * n_par / (phases["adam_ms"] * 1e-3) / 1e9}
    print(json.dumps(line), flush=True)
if world > 1:
    dist.destroy_process_group()
