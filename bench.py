#!/usr/bin/env python
"""Benchmark of the audio -> wav2vec2 -> FaceFormer decoder -> vertices (+ FLAME LBS) hot path.

  python bench.py --gpus N --steps K --warmup W            # product arm (libavi_b200.so on the B200s)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU implementation (oracle port) on host cores

One "step" = one pass of the hot path over one batch of synthetic clips:
  Faceformer.predict (wav2vec2 encoder + autoregressive FaceFormer-disentangle decoder + vertex head, [B,T,15069])
  followed by Faceformer.convert_coeff2verts (FLAME blendshapes + LBS) on B*T frames of 53-d coefficients.
Workload at N=1 = BASELINE.json configs[1]: 64 clips x 10 s of 16 kHz audio (T = 249 frames per clip at 25 fps);
for N>1 each rank processes its own 64 clips (clips shard with no collective: "scaling": "weak", configs[2]).
Metric: generated FLAME frames per second (one frame = one [5023 x 3] fp32 vertex set of the decoder output).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=64, help="clips per GPU per step")
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--fd", type=int, default=64)
    ap.add_argument("--precision", default=os.environ.get("AVI_B200_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--cpu-clips", type=int, default=2, help="clips in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying its CUDA graph")
    ap.add_argument("--cpu-baseline", default=None, choices=["prior", "train", "clip"],
                    help="only time the CPU oracle of another BASELINE config on a bounded sample (the cpu_baseline leg of the "
                         "profiles/*_bench.py GPU scripts, which themselves never touch oracle/) and print one JSON line")
    return ap.parse_args()


def ncu_traffic(kernel_substr):
    """DRAM bytes per launch (read + write) of a kernel from the committed ncu pass over this same command
    (profiles/r1/traffic.json, written by profiles/summarise_launches.py); None when no capture is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1", "traffic.json")) as fh:
            t = json.load(fh)
        for k, v in t.items():
            if kernel_substr in k:
                return float(v["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return dict(hbm=float(p["hbm_gbs"]), tc_burst=float(p["bf16_tflops"]), tc=float(p["bf16_tflops_sustained"]), src="measured")
    except Exception:
        return dict(hbm=6650.0, tc_burst=1590.0, tc=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def workload_config(args, T, precision):
    """The `config` object: identical for the product arm and the --impl reference arm (same workload, BASELINE configs[1])."""
    return {"workload": f"BASELINE configs[1]: FaceFormer-disentangle predict (wav2vec2 + AR decoder + vertex head) + FLAME LBS, "
                        f"{args.clips} clips x {args.seconds:g} s per GPU, fd={args.fd}, random-init (seeded) weights",
            "clips_per_gpu": args.clips, "frames_per_clip": T, "precision": precision, "l2_policy": "inputs larger than L2", "launch": "eager" if getattr(args, "no_graph", False) else "cuda graph replay"}


def n_frames(n_samples):
    n = n_samples
    for k, s in zip((10, 3, 3, 3, 3, 2, 2), (5, 2, 2, 2, 2, 2, 2)):
        n = (n - k) // s + 1
    return int(n / 50.0 * 25)


def make_inputs(clips, n_samples, T, seed):
    """Synthetic z-normalised audio, emotion embeddings and FLAME coefficients for one step, in PINNED host memory."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((clips, n_samples), dtype=np.float32)
    a = (a - a.mean(1, keepdims=True)) / np.sqrt(a.var(1, keepdims=True) + 1e-7)
    host = dict(
        audio=torch.from_numpy(a.astype(np.float32)),
        emo=torch.from_numpy(rng.standard_normal((clips, T, 30), dtype=np.float32)),
        coeff=torch.from_numpy(rng.standard_normal((clips * T, 53), dtype=np.float32)),
        pose=torch.from_numpy((0.1 * rng.standard_normal((clips * T, 6))).astype(np.float32)),
        shape=torch.from_numpy(rng.standard_normal((clips, 100), dtype=np.float32)),
    )
    return host


# ----------------------------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_step(state, audio, emo, coeff, pose, shape_per_frame):
    """The reference's own CPU algorithm for one batch (oracle port: faceformer_disentangle.py predict + convert_coeff2verts)."""
    from oracle import faceformer_oracle as ffo
    from oracle import flame_oracle as fo
    v = ffo.predict(state["sd_ff"], state["sd_w2v"], state["template"], audio, emo, cached=True)   # clip by clip inside
    fv = fo.convert_coeff2verts(state["buf"], state["cmean"], state["cstd"], coeff, pose.clone(), shape_per_frame)
    return v, fv


def cpu_state(fd):
    from avi_talking_b200 import synth
    rng = np.random.default_rng(53)
    buf = synth.flame_buffers()
    return dict(sd_ff=synth.faceformer_state(fd=fd, seed=74), sd_w2v=synth.wav2vec2_state(0), buf=buf,
                template=buf["v_template"].reshape(1, 1, 15069),
                cmean=torch.from_numpy(rng.normal(0, 0.3, size=53).astype("float32")),
                cstd=torch.from_numpy((0.3 + rng.uniform(size=53)).astype("float32")))


def time_cpu(args, n_samples, T, clips, reps):
    torch.set_num_threads(os.cpu_count() or 1)
    st = cpu_state(args.fd)
    host = make_inputs(clips, n_samples, T, seed=4242)
    shape_pf = host["shape"].repeat_interleave(T, 0)
    times = []
    for r in range(reps):
        t0 = time.perf_counter()
        cpu_reference_step(st, host["audio"], host["emo"], host["coeff"], host["pose"], shape_pf)
        times.append(time.perf_counter() - t0)
    return times


def best_of(fn, reps=2):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def cpu_baseline_other(which):
    """cpu_baseline leg for the configs measured by profiles/{prior,train,clip}_bench.py: the oracle port on all host threads, bounded."""
    from avi_talking_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    if which == "prior":      # BASELINE configs[3]: DDIM-64 from 768-d instruction embeddings
        from oracle import prior_oracle as po
        nb = 16
        inp, sd = synth.prior_inputs(nb, 64), synth.prior_state()
        dt = best_of(lambda: po.voxel2style_emb(sd, inp["voxel"], inp["image_embed"], inp["noises"][:63], timesteps_prior=64))
        val, unit, sample = nb / dt, "samples/s", f"DDIM-64 prior sampling of {nb} instruction embeddings"
    elif which == "clip":     # CLIP-L text tower + 77-token mean
        from oracle import clip_oracle as co
        nb = 16
        sd, ids = synth.clip_text_state(60, 12), synth.clip_tokens(nb, seed=61)
        with torch.no_grad():
            dt = best_of(lambda: co.text_to_voxel(sd, ids))
        val, unit, sample = nb / dt, "instructions/s", f"CLIP-L text tower + token mean on {nb} instructions (77 tokens each)"
    else:                     # BASELINE configs[4]: one teacher-forced faceformer_vert step (autograd + Adam) on one 4 s clip
        from oracle import train_oracle as to
        fd, T, N = 64, 120, 64000
        sd_w2v, sd_ff = synth.wav2vec2_state(0), synth.faceformer_state(fd=fd, seed=264, variant="vert")
        template = synth.flame_buffers()["v_template"].reshape(1, 1, 15069)
        gt = template + 1e-3 * torch.from_numpy(np.random.default_rng(7).normal(size=(1, T, 15069)).astype(np.float32))
        audio = synth.audio(1, N, seed=500)
        dt = best_of(lambda: to.train_step(sd_ff, sd_w2v, template, audio, gt, lr=1e-4))
        val, unit, sample = 1.0 / dt, "steps/s", "one training step (forward, autograd backward, Adam) on one 4 s / 120-frame clip"
    print(json.dumps({"impl": "reference", "workload": which, "cpu_baseline": {"value": val, "unit": unit, "cores": torch.get_num_threads(),
                                                                               "kind": "port", "sample": sample, "seconds": dt}}), flush=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_samples = int(round(args.seconds * 16000))
    T = n_frames(n_samples)
    clips = 1  # bounded sample: one 10 s clip per step (the reference decoder is batch-1 anyway, faceformer_disentangle.py:441)
    times = time_cpu(args, n_samples, T, clips, args.warmup + args.steps)[args.warmup:]
    dt = sum(times)
    val = clips * T * args.steps / dt
    line = {
        "impl": "reference", "metric": "generated FLAME frames/sec", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, T, args.precision),
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"each step = {clips} of the {args.clips} clips ({args.seconds:g} s each), oracle restatement of the "
                                   "reference in fp32 torch-CPU on all host threads (KV-cached O(T) decoder, i.e. faster than the "
                                   "reference's O(T^2) loop; the reference itself is not installable: no package, private assets)"},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------- product arm
def run_ours(args):
    import torch.distributed as dist
    from avi_talking_b200 import _lib, ops
    from avi_talking_b200.smoke import build_models

    from avi_talking_b200 import shard
    rank, local, world = shard.env_rank_world()
    # before any pinned allocation: host staging buffers on the GPU's own socket. Only with several ranks: the single-process run also
    # times the CPU baseline, which must keep every host core
    numa = shard.bind_to_gpu_numa(local) if world > 1 else {"bound": False, "why": "single process: all host cores kept"}
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load(check_symbols=True)

    n_samples = int(round(args.seconds * 16000))
    T = n_frames(n_samples)
    B = args.clips
    model = build_models(args.precision, fd=args.fd, device=dev, flame_dir=f"/tmp/avi_flame_assets_r{rank}")
    host = {k: v.pin_memory() for k, v in make_inputs(B, n_samples, T, seed=1000 + rank).items()}
    devin = {k: v.to(dev) for k, v in host.items()}
    def base(t):   # the drop-in returns [..., 15069] views of buffers whose rows are padded to 16 bytes; copy the dense base buffer
        return t._base if t._base is not None else t


    def step_eager(inp):
        # one public call per step: wav2vec2 + AR decoder + vertex head, and FLAME on the frames' coefficients (side stream)
        return model.predict_and_convert(inp["audio"], inp["emo"], inp["coeff"], inp["pose"], inp["shape"].repeat_interleave(T, 0))

    def step(inp):
        # the device-resident `value`: the same call replayed from its CUDA graph (static outputs, overwritten by the next replay)
        if args.no_graph:
            return step_eager(inp)
        return model.graphed_predict_and_convert(inp["audio"], inp["emo"], inp["coeff"], inp["pose"],
                                                 inp["shape"].repeat_interleave(T, 0))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # inputs (41 MB audio + activations of several GB per step) are far larger than the 126 MB L2, so no explicit flush
    for _ in range(max(args.warmup, 3)):
        step(devin)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(devin)
    e1.record()
    barrier()
    dt_ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - n0
    if not args.no_graph:
        # a graph replay makes no C-ABI calls: count the kernels of one eager step (what the graph captured) x steps
        n1 = _lib.launch_count()
        step_eager(devin)
        torch.cuda.synchronize()
        launches = (_lib.launch_count() - n1) * args.steps

    # end to end through the public API with HOST buffers: H2D of the step's inputs and D2H of BOTH results inside the timed
    # region. The D2H of step i runs on a copy stream and overlaps the compute of step i+1 (double-buffered pinned outputs).
    v0, fv0 = step_eager(devin)
    out_host = [torch.empty(base(v0).shape, dtype=torch.float32).pin_memory() for _ in range(2)]
    flame_host = [torch.empty(base(fv0).shape, dtype=torch.float32).pin_memory() for _ in range(2)]
    del v0, fv0
    main = torch.cuda.current_stream()
    copy_stream = torch.cuda.Stream()

    def e2e_step(i):
        inp = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        v, fv = step_eager(inp)          # the public drop-in call (fresh outputs every step, so the D2H of step i overlaps step i+1)
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev)
            out_host[i & 1].copy_(base(v), non_blocking=True)
            flame_host[i & 1].copy_(base(fv), non_blocking=True)
        v.record_stream(copy_stream)
        fv.record_stream(copy_stream)

    for i in range(max(args.warmup, 3)):   # lets the caching allocator reach its steady state on every stream
        e2e_step(i)
    main.wait_stream(copy_stream)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        e2e_step(i)
    main.wait_stream(copy_stream)
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    # roofline pass: one extra step with CUDA events around every launch (not part of the timed numbers above)
    # (the two halves run back to back here, not overlapped, so every kernel is timed alone)
    ops.PROFILE = []
    model.predict_from_embeddings(devin["audio"], devin["emo"])
    model.convert_coeff2verts(devin["coeff"], devin["pose"], devin["shape"].repeat_interleave(T, 0))
    torch.cuda.synchronize()
    prof = ops.PROFILE
    ops.PROFILE = None
    agg = {}
    for name, a, b, work in prof:
        t = a.elapsed_time(b)
        d = agg.setdefault(name, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += t
        d[2] += work
    step_ms_prof = sum(d[1] for d in agg.values())

    dt_ms, e2e_ms = shard.max_over_ranks([dt_ms, e2e_ms], device=dev)   # identity at N=1
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    frames = world * B * T * args.steps
    value = frames / (dt_ms / 1e3)
    e2e_val = frames / (e2e_ms / 1e3)
    g = agg.get("gemm_bf16_tc") or agg.get("gemm_f32")
    gname = "gemm_bf16_tc" if "gemm_bf16_tc" in agg else "gemm_f32"
    achieved = g[2] / (g[1] / 1e3) / 1e12
    roofline = {"kernel": gname, "bound": "tensor", "achieved": achieved, "peak": pk["tc"], "unit": "TFLOP/s",
                "frac": achieved / pk["tc"], "traffic": ncu_traffic("gemm_tc2_kernel") or ncu_traffic("gemm_bf16_tc2_kernel"), "traffic_unit": "DRAM bytes per launch (ncu)", "peak_source": pk["src"] + " (sustained bf16)",
                "launches_per_step": g[0], "avg_launch_ms": g[1] / g[0], "share_of_step": g[1] / step_ms_prof}
    extra = {}
    for name in ("vertex_head", "flame_lbs", "conv0_gn_gelu", "layernorm"):     # HBM-bound kernels: algorithmic bytes / launch time
        k = agg.get(name)
        if k and k[2] > 0:
            gbs = k[2] / (k[1] / 1e3) / 1e9
            extra[name] = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                           "ms": k[1], "launches": k[0]}
    kernels = {k: {"launches": v[0], "ms": round(v[1], 4)} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}

    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = out_host[0].numel() * 4 + flame_host[0].numel() * 4
    line = {
        "metric": "generated FLAME frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": dt_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": workload_config(args, T, args.precision),
        "realtime_factor_25fps": value / 25.0,
        "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches,
        "host_affinity": numa,
        "clocks": sampler.summary(),
        "roofline": roofline,
        "roofline_other": extra,
        "kernels_ms_per_step": kernels,
    }
    if not args.no_cpu_baseline and world == 1:
        t = time_cpu(args, n_samples, T, args.cpu_clips, 2)
        cpu_val = args.cpu_clips * T / min(t)
        line["cpu_baseline"] = {"value": cpu_val, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{args.cpu_clips} of the {B} clips, best of 2 runs, oracle restatement of the reference "
                                          "(fp32 torch-CPU, KV-cached O(T) decoder)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.cpu_baseline:
        cpu_baseline_other(args.cpu_baseline)
    elif args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device for the product arm (there is no CPU fallback); "
                             "use --impl reference for the CPU reference arm")
        run_ours(args)


if __name__ == "__main__":
    main()
