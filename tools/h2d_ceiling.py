#!/usr/bin/env python3
"""What the host link gives: the bytes one bench step moves (pinned H2D of the images, pinned D2H of the results) as bare
cudaMemcpyAsync on two streams, no kernels, on 1..N GPUs AT THE SAME TIME (one process per GPU, started by torchrun, or a single
process for N = 1).  This is the ceiling of bench.py's end-to-end number (`e2e.feed_ceiling` is the same measurement inside the
bench run); VERDICT r1 item 7.

  python tools/h2d_ceiling.py                                     # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 tools/h2d_ceiling.py

Measured on the round-2 boxes (EuRoC step: 184.8 MB in, 35.7 MB out per GPU): 1 GPU 3.4 ms per step (75 k stereo frames/s, 65 GB/s),
2 GPUs 3.8 ms (136 k), 8 GPUs 11.3 ms (181 k frames/s, 156 GB/s through the one NUMA node all eight GPUs hang off)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402

FRAMES, W, H = 256, 752, 480
H2D, D2H = 2 * FRAMES * W * H, 35_663_307


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bench.pin_rank_to_numa_node(local)
    ms = bench.feed_ceiling(torch, H2D, D2H, steps=20)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    if rank == 0:
        gbs = world * (H2D + D2H) / (ms * 1e-3) / 1e9
        print(f"{world} GPU(s): {ms:.3f} ms per step of {H2D / 1e6:.1f} MB in + {D2H / 1e6:.1f} MB out per GPU -> "
              f"{world * FRAMES / (ms * 1e-3):.0f} stereo frames/s, {gbs:.1f} GB/s aggregate")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
