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
    ap.add_argument("--inflight", type=int, default=2, choices=[1, 2],
                    help="predict workload, `value` leg: batches in flight (2 = two CUDA-graph instances replayed alternately on two streams)")
    ap.add_argument("--workload", default="predict", choices=["predict", "prior", "train"],
                    help="predict = BASELINE configs[1]/[2] (default, the headline metric); prior = configs[3] (diffusion prior, batch 256, "
                         "DDIM-64); train = configs[4] (faceformer_vert teacher-forced training step, DDP over the GPUs)")
    ap.add_argument("--clips", type=int, default=None, help="clips per GPU per step (predict: 64; train: 1)")
    ap.add_argument("--clips-total", type=int, default=None,
                    help="predict: STRONG scaling - this many clips split over the GPUs (BASELINE configs[2]: 512 over 2/4/8)")
    ap.add_argument("--prior-batch", type=int, default=256)
    ap.add_argument("--prior-timesteps", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--fd", type=int, default=64)
    ap.add_argument("--precision", default=os.environ.get("AVI_B200_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--cpu-clips", type=int, default=2, help="clips in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--regularisers", default="draw", choices=["draw", "off"],
                    help="train workload: 'draw' = the reference's TRAIN mode (dropout / SpecAugment / LayerDrop drawn on the device every "
                         "step, draws inside the timed region); 'off' = the deterministic .eval() arithmetic")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying its CUDA graph")
    ap.add_argument("--cpu-baseline", default=None, choices=["prior", "train", "clip"],
                    help="only time the CPU oracle of another BASELINE config on a bounded sample (the cpu_baseline leg of the "
                         "profiles/*_bench.py GPU scripts, which themselves never touch oracle/) and print one JSON line")
    return ap.parse_args()


def shutdown_process_group(world, timeout_s=20.0):
    """Leave the process group without ever hanging the launcher: CUDA graphs that captured NCCL kernels must be gone before the
    communicator is (the callers drop them first); the destroy itself runs on a helper thread, and if it has not returned after
    `timeout_s` (seen once on a 2-GPU box: both ranks had printed and passed their last collective) the process exits with status 0 -
    the measurement is complete and printed by then."""
    if world <= 1:
        return
    import gc
    import threading
    import torch.distributed as dist
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    sys.stdout.flush()
    t = threading.Thread(target=dist.destroy_process_group, daemon=True)
    t.start()
    t.join(timeout_s)
    if t.is_alive():
        sys.stderr.write(f"bench.py: destroy_process_group did not return within {timeout_s:g} s; exiting\n")
        sys.stderr.flush()
        os._exit(0)


def ncu_traffic(kernel_substr):
    """DRAM bytes per launch (read + write) of a kernel from the committed ncu pass over this same command
    (profiles/r2/traffic.json, else the round-1 file; written by profiles/summarise_launches.py); None when no capture is committed."""
    for rnd in ("r2", "r1"):
        try:
            with open(os.path.join(ROOT, "profiles", rnd, "traffic.json")) as fh:
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
    which = "configs[1]" if args.clips_total is None else f"configs[2] ({args.clips_total} clips sharded by clip over the GPUs, strong scaling)"
    return {"workload": f"BASELINE {which}: FaceFormer-disentangle predict (wav2vec2 + AR decoder + vertex head) + FLAME LBS, "
                        f"{args.clips} clips x {args.seconds:g} s per GPU, fd={args.fd}, random-init (seeded) weights",
            "clips_per_gpu": args.clips, "frames_per_clip": T, "precision": precision, "l2_policy": "inputs larger than L2", "launch": "eager" if getattr(args, "no_graph", False) else ("cuda graph replay" if getattr(args, "inflight", 1) == 1 else
                       f"cuda graph replay, {args.inflight} batches in flight (independent graph instances on {args.inflight} streams)")}


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
        # per-frame shape rows [clips*T, 100], as convert_coeff2verts takes them (:425): one shape per clip, repeated over its frames
        shape=torch.from_numpy(np.repeat(rng.standard_normal((clips, 100), dtype=np.float32), T, axis=0)),
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
    shape_pf = host["shape"]
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
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
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
        return model.predict_and_convert(inp["audio"], inp["emo"], inp["coeff"], inp["pose"], inp["shape"])

    lanes = [torch.cuda.Stream(device=dev) for _ in range(args.inflight)] if args.inflight > 1 and not args.no_graph else None
    counter = [0]

    def step(inp):
        # the device-resident `value`: the same call replayed from its CUDA graph (static outputs, overwritten by the next replay)
        if args.no_graph:
            return step_eager(inp)
        if lanes is None:
            return model.graphed_predict_and_convert(inp["audio"], inp["emo"], inp["coeff"], inp["pose"], inp["shape"])
        # two batches in flight: graph instance k (own activation pool, own static outputs) on stream k; the streams are joined to the
        # timing stream by fork() / join() around the timed region
        k = counter[0] % len(lanes)
        counter[0] += 1
        with torch.cuda.stream(lanes[k]):
            return model.graphed_predict_and_convert(inp["audio"], inp["emo"], inp["coeff"], inp["pose"], inp["shape"], slot=k)

    def fork():
        if lanes is not None:
            for s_ in lanes:
                s_.wait_stream(torch.cuda.current_stream())

    def join():
        if lanes is not None:
            for s_ in lanes:
                torch.cuda.current_stream().wait_stream(s_)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # inputs (41 MB audio + activations of several GB per step) are far larger than the 126 MB L2, so no explicit flush
    fork()
    for _ in range(max(args.warmup, 3) * args.inflight):
        step(devin)
    join()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    fork()
    for _ in range(args.steps):
        step(devin)
    join()
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

    # OPT-IN compact sink, reported beside the headline, never instead of it: the same end-to-end step with both vertex sets leaving
    # as fp16 displacements from the template (avi_pack_disp_f16): half the D2H bytes, outside the fp32 contract
    tpl = model.template.reshape(-1).float().to(dev)
    c_host = [[torch.empty((B * T, 15069), dtype=torch.float16).pin_memory() for _ in range(2)] for _ in range(2)]

    def e2e_compact_step(i):
        inp = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        v, fv = step_eager(inp)
        pv, pf = ops.pack_disp_f16(v.flatten(0, -2), tpl), ops.pack_disp_f16(fv.reshape(B * T, -1), tpl)
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev)
            c_host[0][i & 1].copy_(pv, non_blocking=True)
            c_host[1][i & 1].copy_(pf, non_blocking=True)
        pv.record_stream(copy_stream)
        pf.record_stream(copy_stream)

    for i in range(3):
        e2e_compact_step(i)
    main.wait_stream(copy_stream)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for i in range(args.steps):
        e2e_compact_step(i)
    main.wait_stream(copy_stream)
    c1.record()
    barrier()
    compact_ms = c0.elapsed_time(c1)
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    # roofline pass: one extra step with CUDA events around every launch (not part of the timed numbers above)
    # (the two halves run back to back here, not overlapped, so every kernel is timed alone)
    ops.PROFILE = []
    model.predict_from_embeddings(devin["audio"], devin["emo"])
    model.convert_coeff2verts(devin["coeff"], devin["pose"], devin["shape"])
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

    dt_ms, e2e_ms, compact_ms = shard.max_over_ranks([dt_ms, e2e_ms, compact_ms], device=dev)   # identity at N=1
    if rank != 0:
        shutdown_process_group(world)
        return

    pk = peaks()
    frames = world * B * T * args.steps
    value = frames / (dt_ms / 1e3)
    e2e_val = frames / (e2e_ms / 1e3)
    g = agg.get("gemm_bf16_tc") or agg.get("gemm_f32")
    gname = "gemm_bf16_tc" if "gemm_bf16_tc" in agg else "gemm_f32"
    achieved = g[2] / (g[1] / 1e3) / 1e12
    roofline = {"kernel": gname, "bound": "tensor", "achieved": achieved, "peak": pk["tc"], "unit": "TFLOP/s",
                "frac": achieved / pk["tc"], "frac_of_burst_peak": achieved / pk["tc_burst"], "peak_burst": pk["tc_burst"],
                "traffic": ncu_traffic("gemm_tc2_kernel") or ncu_traffic("gemm_bf16_tc2_kernel"), "traffic_unit": "DRAM bytes per launch (ncu)",
                "peak_source": pk["src"] + " (sustained bf16: the launches are timed inside one long step; frac_of_burst_peak uses the burst figure, "
                                           "the timed region runs at the burst clock under sw_power_cap, so the fair fraction lies between the two)",
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
        "warmup": max(args.warmup, 3), "ms_per_step": dt_ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": workload_config(args, T, args.precision),
        "realtime_factor_25fps": value / 25.0,
        "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps},
        "e2e_compact_fp16": {"value": frames / (compact_ms / 1e3), "unit": "frames/s", "ms_per_step": compact_ms / args.steps,
                             "d2h_bytes_per_step": 2 * B * T * 15069 * 2,
                             "note": "opt-in sink (frontend.CompactVertexSink): vertices leave as fp16 displacements from the template, "
                                     "2^-11 relative to the displacement; outside the fp32 contract, never the headline"},
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
    shutdown_process_group(world)


# ----------------------------------------------------------------------------------------------------------------- configs[3]: prior
PRIOR_FLOP_PER_SAMPLE_STEP = 12.8e6     # one denoiser pass over the 3 tokens of one sample
BRAIN_FLOP_PER_SAMPLE = 151e6           # BrainNetwork (768 -> 4096 x4 -> 128 + projector)


def prior_config(args):
    return {"workload": f"BASELINE configs[3]: diffusion prior DDIM {args.prior_timesteps}-step sampling from 768-d instruction embeddings "
                        f"(BrainNetwork voxel2clip -> one-launch sampler), batch {args.prior_batch} per GPU, random-init (seeded) weights; "
                        "parity of the dalle2_pytorch semantics is UNPINNED (un-vendored dependency, oracle/dalle2_standin.py)",
            "batch_per_gpu": args.prior_batch, "timesteps": args.prior_timesteps, "precision": args.precision,
            "l2_policy": "L2 flushed between timed iterations (256 MB write)", "launch": "eager (2 GEMM-side launches + 1 sampler launch)"}


def run_reference_prior(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from avi_talking_b200 import synth
    from oracle import prior_oracle as po
    torch.set_num_threads(os.cpu_count() or 1)
    nb = 16                                       # bounded sample: 16 of the 256 instruction embeddings per step
    steps_ddim = args.prior_timesteps - 1 if args.prior_timesteps < 100 else args.prior_timesteps
    inp, sd = synth.prior_inputs(nb, 100), synth.prior_state()
    fn = lambda: po.voxel2style_emb(sd, inp["voxel"], inp["image_embed"], inp["noises"][:steps_ddim], timesteps_prior=args.prior_timesteps)  # noqa: E731
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    val = nb * args.steps / dt
    print(json.dumps({"impl": "reference", "metric": "prior samples/sec", "value": val, "unit": "samples/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": prior_config(args),
                      "cpu_baseline": {"value": val, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                                       "sample": f"each step = {nb} of the {args.prior_batch} samples, oracle restatement (fp32 torch-CPU, all host threads)"},
                      "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def run_prior(args):
    import torch.distributed as dist
    from avi_talking_b200 import _lib, ops, shard, synth
    from avi_talking_b200.diffusion_prior import voxel2style_emb
    from avi_talking_b200.smoke import build_prior
    rank, local, world = shard.env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load(check_symbols=True)
    B, ts = args.prior_batch, args.prior_timesteps
    n_noise = ts - 1 if ts < 100 else ts           # DDIM draws no noise on its last pair; DDPM on every step but t = 0
    inp = synth.prior_inputs(B, 100, seed=7 + rank)
    prior = build_prior(args.precision, device=dev)
    host_voxel = inp["voxel"].pin_memory()
    voxel, x0, noise = inp["voxel"].to(dev), inp["image_embed"].to(dev), inp["noises"][:n_noise].to(dev)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)   # 256 MB > 126 MB L2

    def step(v):
        return voxel2style_emb(v, prior, timesteps_prior=ts, image_embed=x0, noise=noise)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """CUDA events around every step, the L2 flush between steps outside the events."""
        tot = 0.0
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot

    for _ in range(max(args.warmup, 3)):
        step(voxel)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    n0 = _lib.launch_count()
    dt_ms = timed(lambda: step(voxel), args.steps)
    launches = _lib.launch_count() - n0
    out_host = torch.empty(B, 1, 128).pin_memory()

    def e2e():
        v = host_voxel.to(dev, non_blocking=True)
        out_host.copy_(step(v), non_blocking=True)

    for _ in range(3):
        e2e()
    barrier()
    e2e_ms = timed(e2e, args.steps)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    ops.PROFILE = []
    step(voxel)
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    agg = {}
    for name, a, b, work in prof:
        d = agg.setdefault(name, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += a.elapsed_time(b)
        d[2] += work
    dt_ms, e2e_ms = shard.max_over_ranks([dt_ms, e2e_ms], device=dev)
    if rank == 0:
        pk = peaks()
        k = agg.get("prior_sample")
        samp_ms = k[1] if k else dt_ms / args.steps
        achieved = PRIOR_FLOP_PER_SAMPLE_STEP * B * n_noise / (samp_ms / 1e3) / 1e12
        line = {"metric": "prior samples/sec", "value": world * B * args.steps / (dt_ms / 1e3), "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dt_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": prior_config(args),
                "e2e": {"value": world * B * args.steps / (e2e_ms / 1e3), "unit": "samples/s", "h2d_bytes_per_step": host_voxel.numel() * 4,
                        "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": launches, "clocks": sampler.summary(),
                "roofline": {"kernel": "prior_sample_kernel", "bound": "tensor", "achieved": achieved, "peak": pk["tc_burst"], "unit": "TFLOP/s",
                             "frac": achieved / pk["tc_burst"], "traffic": None, "peak_source": pk["src"] + " (burst bf16: a kernel timed alone)",
                             "launches_per_step": 1, "avg_launch_ms": samp_ms,
                             "note": "the whole sampling loop is ONE launch; the denoiser is a 3-token transformer (12.8 MFLOP per "
                                     "sample-step), so the tensor roofline is nominal: the kernel is latency / L2-weight-stream bound"},
                "kernels_ms_per_step": {k2: {"launches": v[0], "ms": round(v[1], 4)} for k2, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}}
        if not args.no_cpu_baseline and world == 1:
            from oracle import prior_oracle as po
            torch.set_num_threads(os.cpu_count() or 1)
            nb = 16
            ci, sd = synth.prior_inputs(nb, 100), synth.prior_state()
            t = best_of(lambda: po.voxel2style_emb(sd, ci["voxel"], ci["image_embed"], ci["noises"][:n_noise], timesteps_prior=ts))
            line["cpu_baseline"] = {"value": nb / t, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"{nb} of the {B} samples, best of 2, oracle restatement (fp32 torch-CPU)"}
        print(json.dumps(line), flush=True)
    shutdown_process_group(world)


# ----------------------------------------------------------------------------------------------------------------- configs[4]: train
TRAIN_T, TRAIN_N = 120, 64000            # one VOCASET-like clip: 4 s of 16 kHz audio, 120 frames at 30 fps


def train_config(args, n_par, world):
    return {"workload": f"BASELINE configs[4]: faceformer_vert teacher-forced training step (wav2vec2 fwd+bwd with the conv extractor "
                        f"frozen, decoder layer, 15069-wide vertex head, MSE x 10, Adam), {args.clips} clip(s) x 4 s ({TRAIN_T} frames) per GPU, "
                        "fd=64, " + ("TRAIN mode: dropout 0.1 / SpecAugment / LayerDrop drawn on the device every step inside the timed region"
                                     if getattr(args, "regularisers", "off") == "draw" else "deterministic mode (dropout / SpecAugment / LayerDrop off)")
                        + ", random-init (seeded) weights",
            "regularisers": getattr(args, "regularisers", "off"),
            "clips_per_gpu": args.clips, "precision": args.precision, "trainable_params": n_par,
            "allreduce": "bucketed NCCL sum of the flat fp32 gradient, captured in the step's CUDA graph, overlapped with backward" if world > 1 else "none",
            "l2_policy": "working set (weights + moments + gradients = 1.5 GB) larger than L2", "launch": "eager" if args.no_graph else "cuda graph replay (forward + backward + all-reduce), Adam launched after it"}


def train_flops(clips):
    """Algorithmic FLOPs of one step: encoder (proj, pos-conv, 12 layers incl. attention) forward x 3 (dX and dW GEMMs), the frozen
    conv extractor forward only, decoder + vertex head forward x 3."""
    T50 = 199
    conv = 0.0
    L = (TRAIN_N - 10) // 5 + 1
    for k in (3, 3, 3, 3, 2, 2):
        L = (L - k) // 2 + 1
        conv += 2.0 * L * 512 * 512 * k
    T = TRAIN_T
    enc = 2.0 * T * 768 * 512 + 2.0 * T * 768 * 48 * 128 + 12 * (2.0 * T * 768 * 9216 + 4.0 * T * T * 768)
    head = 2.0 * T * 64 * 15069 * 2 + 2.0 * T * 64 * (768 + 64 * 10)
    del T50
    return clips * (conv + 3.0 * (enc + head))


def run_reference_train(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from avi_talking_b200 import synth
    from oracle import train_oracle as to
    torch.set_num_threads(os.cpu_count() or 1)
    fd = 64
    sd_w2v, sd_ff = synth.wav2vec2_state(0), synth.faceformer_state(fd=fd, seed=264, variant="vert")
    template = synth.flame_buffers()["v_template"].reshape(1, 1, 15069)
    gt = template + 1e-3 * torch.from_numpy(np.random.default_rng(7).normal(size=(1, TRAIN_T, 15069)).astype(np.float32))
    audio = synth.audio(1, TRAIN_N, seed=500)
    fn = lambda: to.train_step(sd_ff, sd_w2v, template, audio, gt, lr=1e-4)  # noqa: E731
    for _ in range(min(args.warmup, 1)):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    val = args.steps / dt
    print(json.dumps({"impl": "reference", "metric": "training clips/sec", "value": val, "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": train_config(args, None, 1),
                      "cpu_baseline": {"value": val, "unit": "clips/s", "cores": torch.get_num_threads(), "kind": "port",
                                       "sample": "each step = one training step (forward, autograd backward, Adam) on one 4 s / 120-frame clip, "
                                                 "oracle restatement under torch autograd (fp32 torch-CPU, all host threads)"},
                      "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def run_train(args):
    import torch.distributed as dist
    from transformers import Wav2Vec2Config

    from avi_talking_b200 import _lib, shard, synth, train
    from avi_talking_b200.faceformer import FaceformerVert, make_args
    from avi_talking_b200.wav2vec import Wav2Vec2Model
    rank, local, world = shard.env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load(check_symbols=True)
    fd, B, T, N = 64, args.clips, TRAIN_T, TRAIN_N
    w2v = Wav2Vec2Model(Wav2Vec2Config())
    w2v.load_state_dict(synth.wav2vec2_state(0), strict=False)
    template = synth.flame_buffers()["v_template"].reshape(1, 1, 15069)
    m = FaceformerVert(make_args(feature_dim=fd), audio_encoder=w2v, template=template)
    m.load_state_dict(synth.faceformer_state(fd=fd, seed=264, variant="vert"), strict=False)
    m.precision = w2v.precision = args.precision
    m = m.to(dev)
    opt = train.FlatAdam(m, lr=1e-4)
    buckets = train.GradBuckets(m._flat_layout) if world > 1 else None
    rng = np.random.default_rng(7 + rank)
    host_gt = (template + 1e-3 * torch.from_numpy(rng.normal(size=(B, T, 15069)).astype(np.float32))).pin_memory()
    host_audio = synth.audio(B, N, seed=500 + rank * B).pin_memory()
    gt, audio = host_gt.to(dev), host_audio.to(dev)
    cfg_w2v = w2v.config
    # the reference trains in .train() mode: every step draws its dropout masks, SpecAugment spans and LayerDrop decisions. Here the
    # draws are three launches of the library (Philox, csrc/train_draw.cu) captured inside the step's graph
    device_draws = train.DeviceDraws(B, T, fd, cfg_w2v, dev, seed=11 + rank) if args.regularisers == "draw" else None

    def draws():
        return device_draws.draw() if device_draws is not None else None

    if args.no_graph:
        step = train.TrainStep(m, buckets=buckets)
        m._train_step = step

        def one_step(a, g):
            opt.zero_grad()
            loss = m.training_loss(a, g, reg=draws())
            loss.backward()
            opt.step()
            return loss
    else:
        gstep = train.GraphedTrainStep(m, audio.shape, gt.shape, buckets=buckets, max_graphs=16)

        def one_step(a, g):
            loss = gstep(a, g, reg=device_draws)               # the draw launches live inside the captured graph
            opt.step()
            return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        one_step(audio, gt)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = one_step(audio, gt)
    e1.record()
    barrier()
    dt_ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - n0
    # end to end: the step's audio and ground-truth vertices come from pinned host memory, the loss is read back every step
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def e2e_step():
        a, g = host_audio.to(dev, non_blocking=True), host_gt.to(dev, non_blocking=True)
        loss_host.copy_(one_step(a, g), non_blocking=True)

    for _ in range(3):
        e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        e2e_step()
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)

    sampler.stop_flag.set()
    sampler.join(timeout=2)
    # exposed exchange = this step minus the same step without the all-reduce is not separable inside a graph; report the Adam kernel
    # (HBM-bound: 28 B per parameter) and the in-sync check instead
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    one_step(audio, gt)
    a0.record()
    opt.step()
    a1.record()
    torch.cuda.synchronize()
    adam_ms = a0.elapsed_time(a1)
    in_sync = True
    if world > 1:
        chk = torch.stack([m._flat_params.double().sum(), m._flat_params.double().abs().sum()])
        hi, lo = chk.clone(), chk.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        in_sync = bool((hi == lo).all())
    dt_ms, e2e_ms = shard.max_over_ranks([dt_ms, e2e_ms], device=dev)
    if rank == 0:
        pk = peaks()
        n_par = m._flat_layout.total
        ms = dt_ms / args.steps
        if not args.no_graph:      # a graph replay makes no C-ABI calls: count one eager step's launches
            st = train.TrainStep(m, buckets=None)
            n1 = _lib.launch_count()
            st.forward(audio, gt)
            st.backward()
            torch.cuda.synchronize()
            launches = (_lib.launch_count() - n1 + 1) * args.steps
        achieved = train_flops(B) / (ms / 1e3) / 1e12
        line = {"metric": "training clips/sec", "value": world * B * 1e3 / ms, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic", "config": train_config(args, n_par, world),
                "e2e": {"value": world * B * args.steps / (e2e_ms / 1e3), "unit": "clips/s", "h2d_bytes_per_step": (host_audio.numel() + host_gt.numel()) * 4,
                        "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": launches, "clocks": sampler.summary(), "loss": float(loss.detach()), "ranks_in_sync": in_sync,
                "roofline": {"kernel": "adam_step_kernel", "bound": "hbm", "achieved": 28.0 * n_par / (adam_ms / 1e3) / 1e9, "peak": pk["hbm"],
                             "unit": "GB/s", "frac": 28.0 * n_par / (adam_ms / 1e3) / 1e9 / pk["hbm"], "traffic": None,
                             "peak_source": pk["src"], "avg_launch_ms": adam_ms,
                             "note": "batch 1 x 4 s is launch / latency bound (about 700 short kernels of 120-row operands); the one "
                                     "bandwidth-bound kernel is Adam (28 B per parameter). Whole-step algorithmic rate below."},
                "step_tflops_algorithmic": achieved}
        if not args.no_cpu_baseline and world == 1:
            from oracle import train_oracle as to
            torch.set_num_threads(os.cpu_count() or 1)
            sd_w2v, sd_ff = synth.wav2vec2_state(0), synth.faceformer_state(fd=fd, seed=264, variant="vert")
            cgt = template + 1e-3 * torch.from_numpy(np.random.default_rng(7).normal(size=(1, T, 15069)).astype(np.float32))
            t = best_of(lambda: to.train_step(sd_ff, sd_w2v, template, synth.audio(1, N, seed=500), cgt, lr=1e-4))
            line["cpu_baseline"] = {"value": 1.0 / t, "unit": "clips/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "one training step on one 4 s clip, best of 2, oracle restatement under torch autograd (fp32 torch-CPU)"}
        print(json.dumps(line), flush=True)
    if not args.no_graph:          # graphs that captured NCCL kernels go before the communicator does
        gstep.graphs.clear()
        gstep.graph = None
    shutdown_process_group(world)


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.scaling = "weak"
    if args.workload == "predict":
        if args.clips_total is not None:
            if args.clips_total % world != 0:
                raise SystemExit(f"--clips-total {args.clips_total} is not divisible by the {world} ranks")
            args.clips, args.scaling = args.clips_total // world, "strong"
        elif args.clips is None:
            args.clips = 64
    elif args.clips is None:
        args.clips = 1
    if args.cpu_baseline:
        cpu_baseline_other(args.cpu_baseline)
    elif args.impl == "reference":
        {"predict": run_reference, "prior": run_reference_prior, "train": run_reference_train}[args.workload](args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device for the product arm (there is no CPU fallback); "
                             "use --impl reference for the CPU reference arm")
        {"predict": run_ours, "prior": run_prior, "train": run_train}[args.workload](args)


if __name__ == "__main__":
    main()
