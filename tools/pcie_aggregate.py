"""Aggregate host <-> device copy bandwidth of the box with 1 .. N GPUs copying at once (pinned memory, both directions
together, the way the end-to-end legs of bench.py use the link): names the limiter of the e2e scaling figures.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_aggregate.py
"""
import os
import time

import torch
import torch.distributed as dist


def main():
  rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
  torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
  if world > 1:
    dist.init_process_group("nccl")
  nbytes = 2 << 30
  h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
  h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
  d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
  d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
  s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
  active = 1
  while active <= world:
    for mode in ("h2d", "d2h", "both"):
      torch.cuda.synchronize()
      if world > 1:
        dist.barrier()
      t0 = time.perf_counter()
      if rank < active:
        for _ in range(4):
          if mode in ("h2d", "both"):
            with torch.cuda.stream(s_in):
              d_in.copy_(h_in, non_blocking=True)
          if mode in ("d2h", "both"):
            with torch.cuda.stream(s_out):
              h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
      dt = torch.tensor([time.perf_counter() - t0 if rank < active else 0.0], device="cuda")
      if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
      if rank == 0:
        per_dir = 4 * nbytes / float(dt[0]) / 1e9
        print("%d GPU(s) copying, %-4s: %6.1f GB/s per GPU and direction, %7.1f GB/s aggregate" %
              (active, mode, per_dir, per_dir * active * (2 if mode == "both" else 1)), flush=True)
    active *= 2
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
