"""TEST INFRASTRUCTURE (checker only; never imported by the product path).

numpy restatement of the device-side draws of a TRAIN-mode training step (avi_talking_b200/csrc/train_draw.cu, include/avi_b200.h
avi_dropout_masks / avi_layerdrop_spec_draw). The generator is Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as
easy as 1, 2, 3", SC'11) - a published algorithm, pinned here by the three known-answer vectors of the Random123 distribution
(tests/test_oracle_golden.py::test_philox_known_answers); the reference itself draws through torch's generator inside its modules
(nn.Dropout; np.random.uniform for LayerDrop; _compute_mask_indices, models/lib/wav2vec.py:16-63), a stream no other implementation
can reproduce, which is why parity of the training step is pinned with INJECTED draws (tests/golden/train_reg.npz) and the draw
kernels are pinned against this file bit for bit.
"""
from __future__ import annotations

import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
STREAM_LAYERDROP, STREAM_SPEC_START, STREAM_SPEC_COUNT = 0x4C440000, 0x53500000, 0x534E0000


def philox4x32_10(ctr, key):
    """ctr: 4 uint32 arrays (broadcastable), key: 2 uint32 scalars -> 4 uint32 arrays."""
    mask = np.uint64(0xFFFFFFFF)
    c = [np.asarray(x, dtype=np.uint64) & mask for x in np.broadcast_arrays(*[np.asarray(x, dtype=np.uint64) for x in ctr])]
    k0, k1 = np.uint64(key[0]) & mask, np.uint64(key[1]) & mask
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k0) & mask, p1 & mask, ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & mask, p0 & mask]
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return [x.astype(np.uint32) for x in c]


def u01(x):
    return (np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def dropout_masks(n, p, seed, step, stream_id=0):
    """The flat mask buffer of avi_dropout_masks: float32 [n], n % 4 == 0."""
    assert n % 4 == 0
    g = np.arange(n // 4, dtype=np.uint64)
    words = philox4x32_10((g & np.uint64(0xFFFFFFFF), g >> np.uint64(32), np.uint64(step), np.uint64(stream_id)),
                          (seed & 0xFFFFFFFF, seed >> 32))
    u = u01(np.stack(words, axis=1).reshape(-1))
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return np.where(u >= np.float32(p), scale, np.float32(0.0)).astype(np.float32)


def layer_keep(n_layers, layerdrop, seed, step):
    l = np.arange(n_layers, dtype=np.uint64)
    w = philox4x32_10((l, np.uint64(0), np.uint64(step), np.uint64(STREAM_LAYERDROP)), (seed & 0xFFFFFFFF, seed >> 32))[0]
    return u01(w) >= np.float32(layerdrop)


def spec_mask(B, T, span_len, span_rate, min_spans, seed, step):
    key = (seed & 0xFFFFFFFF, seed >> 32)
    out = np.zeros((B, T), dtype=np.uint8)
    if span_len >= T:
        return out
    un = u01(philox4x32_10((0, 0, step, STREAM_SPEC_COUNT), key)[0])
    n = max(int(min_spans), int(np.floor(np.float32(span_rate) + un)))
    for b in range(B):
        for s in range(n):
            x = int(philox4x32_10((b, s, step, STREAM_SPEC_START), key)[0])
            start = (x * (T - span_len + 1)) >> 32
            out[b, start:start + span_len] = 1
    return out
