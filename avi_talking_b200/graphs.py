"""CUDA-graph replay of a fixed-shape call (B200-first: streams and graphs, no tracing compiler).

The configs[1] step is ~125 kernels of 10-1400 us; replaying them from one graph removes the launch gaps between them and all the
host work (ctypes calls, tensor-map encodes, allocator traffic). `GraphedCall` captures `fn(*tensors)` once per input signature
and per weight version; inputs are copied into static buffers, outputs are the graph's static output tensors (valid until the
next call with the same signature).
"""
from __future__ import annotations

import torch

from . import ops


class GraphedCall:
    def __init__(self, fn, weight_modules=(), warmup=2):
        self.fn, self.mods, self.warmup = fn, tuple(weight_modules), warmup
        self.cache = {}

    def _weights_key(self):
        k = [ops.WEIGHT_EPOCH]
        for m in self.mods:
            k.extend((p.data_ptr(), p._version) for p in m.parameters())
            k.extend((b.data_ptr(), b._version) for b in m.buffers())
        return tuple(k)

    def _capture(self, tensors):
        static = [torch.empty_like(t).copy_(t) for t in tensors]
        side = torch.cuda.Stream(device=tensors[0].device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):            # packs weights, sets kernel attributes, warms the allocator
                self.fn(*[s.clone() for s in static])
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.fn(*static)
        return static, graph, out

    def __call__(self, *tensors):
        for t in tensors:
            if not t.is_cuda:
                raise RuntimeError("GraphedCall needs CUDA tensors (there is no CPU path)")
        key = tuple((tuple(t.shape), t.dtype, t.device.index) for t in tensors) + self._weights_key()
        ent = self.cache.get(key)
        if ent is None:
            self.cache.clear()                       # one live signature at a time: graphs pin their memory pools
            ent = self.cache[key] = self._capture(tensors)
        static, graph, out = ent
        for s, t in zip(static, tensors):
            s.copy_(t, non_blocking=True)
        graph.replay()
        return out
