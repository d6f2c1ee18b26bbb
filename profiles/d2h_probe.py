"""Platform ceiling of the end-to-end sink: bare pinned device->host copies, no compute.

Every rank (one process per GPU, same torchrun launch as bench.py) copies the bytes one bench step returns - two fp32 vertex
sets of [clips*T, 15072] = 1.92 GB at 64 clips - from device memory to pinned host memory with ONE cudaMemcpyAsync per
buffer (tensor.copy_(non_blocking=True)), all ranks at the same time. Reported: per-rank and aggregate GB/s, the max-over-ranks
time per "step" (what bounds bench.py's e2e), and the same for the host->device direction of the step's inputs.

  python profiles/d2h_probe.py                                   # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 profiles/d2h_probe.py

Variants: --chunks K splits each buffer into K copies (does the link care about transfer size?), --bytes overrides the size.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=2 * 64 * 249 * 15072 * 4)
    ap.add_argument("--h2d-bytes", type=int, default=46658816)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--chunks", type=int, default=1)
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_buf = 2
    per = args.bytes // n_buf // 4
    src = [torch.empty(per, dtype=torch.float32, device=dev).normal_() for _ in range(n_buf)]
    dst = [torch.empty(per, dtype=torch.float32).pin_memory() for _ in range(n_buf)]
    hin = torch.empty(args.h2d_bytes // 4, dtype=torch.float32).pin_memory()
    din = torch.empty_like(hin, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def d2h():
        for s, d in zip(src, dst):
            if args.chunks == 1:
                d.copy_(s, non_blocking=True)
            else:
                for sc, dc in zip(s.chunk(args.chunks), d.chunk(args.chunks)):
                    dc.copy_(sc, non_blocking=True)

    def timed(fn):
        fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / args.reps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    ms_d2h = timed(d2h)
    ms_h2d = timed(lambda: din.copy_(hin, non_blocking=True))
    if rank == 0:
        b = n_buf * per * 4
        print(json.dumps({"probe": "pinned D2H / H2D, no compute, all ranks concurrently", "n_gpus": world, "chunks": args.chunks,
                          "d2h_bytes_per_rank": b, "d2h_ms_max_over_ranks": ms_d2h, "d2h_gbs_per_rank": b / ms_d2h / 1e6,
                          "d2h_gbs_aggregate": world * b / ms_d2h / 1e6,
                          "h2d_bytes_per_rank": args.h2d_bytes, "h2d_ms_max_over_ranks": ms_h2d,
                          "h2d_gbs_per_rank": args.h2d_bytes / ms_h2d / 1e6}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
