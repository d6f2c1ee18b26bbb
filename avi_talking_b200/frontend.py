"""Audio front end and result sink around the EMOTE / FaceFormer inference path.

Reference (third_party/inferno/inferno_apps/TalkingHead/evaluation/evaluation_functions.py):
  process_audio :690-714          int16 wav -> `raw_audio [num_frames, samplerate // fps]` (truncate / zero-pad to whole video frames)
  create_base_sample :141-160     pad to a multiple of `smallest_unit` frames, optional silent frames, zero gt_exp / gt_jaw / gt_shape / gt_tex
  run_evalutation :624-638        per clip: predicted_exp / predicted_jaw -> .cpu().numpy() -> one pickle per clip
and AudioEncoders.py:170-178 (Wav2Vec2Processor z-normalisation), which the GPU path does in `avi_audio_znorm`.

What changes: clips are framed and batched once on the host (integer / byte work: bit-exact with the reference), the int16 -> float cast and
the z-normalisation run on the GPU, and results leave through ONE double-buffered pinned host buffer per output on a copy stream
(the D2H of batch i overlaps the compute of batch i+1) instead of a synchronising `.cpu()` per clip.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def process_audio(wavdata: np.ndarray, sampling_rate: int, video_fps: int) -> dict:
    """evaluation_functions.py:690-714, same results (vectorised: one slice copy instead of a zero buffer + conditional copy)."""
    assert sampling_rate % video_fps == 0
    wav_per_frame = sampling_rate // video_fps
    num_frames = wavdata.shape[0] // wav_per_frame
    out = np.zeros(num_frames * wav_per_frame, dtype=wavdata.dtype)
    n = min(out.size, wavdata.size)
    out[:n] = wavdata[:n]
    return {"raw_audio": out.reshape(num_frames, wav_per_frame), "samplerate": sampling_rate}


def create_base_sample(wavdata: np.ndarray, sampling_rate: int = 16000, video_fps: int = 25, smallest_unit: int = 1,
                       silent_frames_start: int = 0, silent_frames_end: int = 0, silence_all: bool = False, n_shape: int = 300,
                       n_exp: int = 50, n_tex: int = 50) -> dict:
    """evaluation_functions.py:141-160 from an already loaded int16 waveform (read_audio needs librosa and a file: host I/O, out of
    scope). Note the upstream pad `smallest_unit - T % smallest_unit` adds a full unit when T is already a multiple (for
    smallest_unit = 1 one extra, all-zero frame... and, being np.pad with a scalar pad tuple, it pads the LAST axis too); both quirks
    are reproduced because downstream frame counts depend on them."""
    sample = process_audio(wavdata, sampling_rate, video_fps)
    raw = sample["raw_audio"]
    raw = np.pad(raw, (0, smallest_unit - raw.shape[0] % smallest_unit))
    if silent_frames_start > 0:
        raw = np.concatenate([np.zeros((silent_frames_start, raw.shape[1]), dtype=raw.dtype), raw], axis=0)
    if silent_frames_end > 0:
        raw = np.concatenate([raw, np.zeros((silent_frames_end, raw.shape[1]), dtype=raw.dtype)], axis=0)
    if silence_all:
        raw = np.zeros_like(raw)
    T = raw.shape[0]
    sample["raw_audio"] = raw
    sample["gt_exp"] = np.zeros((T, n_exp), dtype=np.float32)
    sample["gt_shape"] = np.zeros((n_shape,), dtype=np.float32)
    sample["gt_jaw"] = np.zeros((T, 3), dtype=np.float32)
    sample["gt_tex"] = np.zeros((n_tex,), dtype=np.float32)
    return sample


def batch_samples(samples: list[dict], device="cuda") -> dict:
    """Stack per-clip samples of equal length into the batch dict TalkingHeadWrapper.forward consumes: raw_audio stays int16 through
    the (pinned, asynchronous) host->device copy - half the bytes of fp32 - and is cast on the GPU."""
    T = samples[0]["raw_audio"].shape[0]
    if any(s["raw_audio"].shape != samples[0]["raw_audio"].shape for s in samples):
        raise ValueError("batch_samples: clips must have the same number of frames (bucket them by length first)")
    out = {}
    raw = torch.from_numpy(np.stack([s["raw_audio"] for s in samples]))
    pinned = raw.pin_memory() if torch.cuda.is_available() else raw
    out["raw_audio"] = pinned.to(device, non_blocking=True).float()
    out["samplerate"] = [s["samplerate"] for s in samples]
    for k in ("gt_exp", "gt_shape", "gt_jaw", "gt_tex"):
        if k in samples[0]:
            out[k] = torch.from_numpy(np.stack([s[k] for s in samples])).to(device, non_blocking=True)
    out["_frames"] = T
    return out


class ResultSink:
    """Double-buffered pinned host buffers for named device results (e.g. predicted_exp [B,T,50], predicted_jaw [B,T,3]).

        sink = ResultSink(("predicted_exp", "predicted_jaw"))
        for batch in batches:
            out = wrapper(batch)
            done = sink.push(out)            # async D2H of this batch on the copy stream; returns the PREVIOUS batch (host views)
        last = sink.flush()

    `flame_dicts(host_batch, gt_shape)` yields, per clip, the dict the reference pickles (evaluation_functions.py:624-632)."""

    def __init__(self, keys, device=None):
        self.keys = tuple(keys)
        self.stream = None
        self.bufs = [dict(), dict()]
        self.events = [None, None]
        self.shapes = [None, None]
        self.i = 0
        self.pending = None

    def _collect(self, slot):
        if self.events[slot] is None:
            return None
        self.events[slot].synchronize()
        out = {}
        for k in self.keys:
            shp = self.shapes[slot][k]
            n = int(np.prod(shp)) if len(shp) else 1
            out[k] = self.bufs[slot][k][:n].view(shp)
        return out

    def push(self, results: dict):
        dev = results[self.keys[0]].device
        if self.stream is None:
            self.stream = torch.cuda.Stream(device=dev)
        slot = self.i & 1
        prev = self._collect(slot ^ 1) if self.pending is not None else None
        if self.events[slot] is not None:
            self.events[slot].synchronize()          # the host has consumed this slot two pushes ago; make sure its copy finished
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        shapes = {}
        with torch.cuda.stream(self.stream):
            for k in self.keys:
                t = results[k].contiguous()
                buf = self.bufs[slot].get(k)
                if buf is None or buf.numel() < t.numel() or buf.dtype != t.dtype:
                    buf = self.bufs[slot][k] = torch.empty((t.numel(),), dtype=t.dtype).pin_memory()
                buf[: t.numel()].copy_(t.reshape(-1), non_blocking=True)
                t.record_stream(self.stream)
                shapes[k] = (t.numel(),) if t.dim() == 0 else tuple(t.shape)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        # views are taken over the flat buffer: store (numel,) + shape for the reshaping in _collect
        self.shapes[slot] = {k: shapes[k] for k in self.keys}
        self.events[slot] = ev
        self.pending = slot
        self.i += 1
        return prev

    def flush(self):
        if self.pending is None:
            return None
        out = self._collect(self.pending)
        self.pending = None
        return out

    @staticmethod
    def flame_dicts(host_batch: dict, gt_shape) -> list[dict]:
        exp, jaw = host_batch["predicted_exp"].numpy(), host_batch["predicted_jaw"].numpy()
        gt_shape = np.asarray(gt_shape)
        return [{"shape": gt_shape[b], "expression": exp[b], "jaw_pose": jaw[b], "global_pose": np.zeros_like(jaw[b])}
                for b in range(exp.shape[0])]


class CompactVertexSink(ResultSink):
    """OPT-IN sink for vertex tensors: every pushed [..., 15069] fp32 result leaves the GPU as the fp16 displacement from a template
    (`avi_pack_disp_f16`), halving the device->host bytes that bound the end-to-end path; `unpack` restores fp32 on the host.
    Precision: 2^-11 relative to the displacement (a few 1e-6 m for speech-driven motion). Outside the fp32 contract of the path -
    ResultSink is the default and every headline number ships fp32."""

    def __init__(self, keys, templates: dict):
        super().__init__(keys)
        self.templates = {k: v.detach().reshape(-1).float() for k, v in templates.items()}

    def push(self, results: dict):
        packed = {}
        for k in self.keys:
            t = results[k]
            Cc = self.templates[k].numel()
            rows = t.reshape(-1, t.shape[-1]) if t.is_contiguous() else t.flatten(0, -2)
            packed[k] = ops.pack_disp_f16(rows, self.templates[k].to(t.device)).view(*t.shape[:-1], Cc)
        return super().push(packed)

    def unpack(self, host_batch: dict) -> dict:
        """fp16 displacements (host) -> fp32 vertices (host, numpy-side addition in the consumer's precision)."""
        return {k: v.float() + self.templates[k].cpu() for k, v in host_batch.items()}
