"""Ping-pong frame indexer used by Faceformer.predict (upstream loop_utils.py:4-16)."""
import torch


def calc_loop_idx(idx, loop_num):
    """Index into a clip of ``loop_num`` frames played forwards then backwards with the end frames repeated:
    0..n-1, n-1..0, 0..n-1, ... (loop_utils.py:4-7)."""
    forward = (idx // loop_num) % 2 == 0
    r = idx % loop_num
    return r if forward else loop_num - 1 - r


def loopback_frames(img, frame_num):
    """loop_utils.py:10-16: gather ``frame_num`` frames from ``img`` in ping-pong order."""
    n = img.shape[0]
    idx = torch.tensor([calc_loop_idx(i, n) for i in range(frame_num)], dtype=torch.long, device=img.device)
    return img.index_select(0, idx)
