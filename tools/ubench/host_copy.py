#!/usr/bin/env python
"""Platform ceiling of the end-to-end (host-buffer) path: concurrent H2D + D2H of the configs[1] byte counts from pinned host
memory with NO kernel, one process per GPU (torchrun), every rank copying at the same time.

    python tools/ubench/host_copy.py                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/ubench/host_copy.py

Per step and rank: 1 966 080 000 B host->device (1024 clips x 480 000 f32 samples) and 1 573 388 288 B device->host
(1024 x 128 x 3001 f32), issued on two streams in chunks (default 64 MB, the library's staging granularity; also 256 MB and
whole-buffer) so that both copy engines run concurrently. Prints one JSON line: per-rank and aggregate GB/s in each direction,
ms per step (max over ranks) and the frames/s this would allow if the kernels were free -- the number bench.py's `e2e` is
compared against (`e2e.ceiling`)."""
import json
import os
import time

import torch
import torch.distributed as dist

H2D_BYTES = 1024 * 480000 * 4
D2H_BYTES = 1024 * 128 * 3001 * 4
FRAMES = 1024 * 3001


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    h_in = torch.empty(H2D_BYTES // 4, dtype=torch.float32, pin_memory=True)
    h_out = torch.empty(D2H_BYTES // 4, dtype=torch.float32, pin_memory=True)
    h_in.fill_(1.0)
    d_in = torch.empty_like(h_in, device=dev)
    d_out = torch.zeros(D2H_BYTES // 4, dtype=torch.float32, device=dev)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(chunk_bytes, directions):
        n_in = chunk_bytes // 4 if chunk_bytes else h_in.numel()
        n_out = chunk_bytes // 4 if chunk_bytes else h_out.numel()
        if "h2d" in directions:
            with torch.cuda.stream(sa):
                for o in range(0, h_in.numel(), n_in):
                    d_in[o:o + n_in].copy_(h_in[o:o + n_in], non_blocking=True)
        if "d2h" in directions:
            with torch.cuda.stream(sb):
                for o in range(0, h_out.numel(), n_out):
                    h_out[o:o + n_out].copy_(d_out[o:o + n_out], non_blocking=True)

    res = {}
    for name, chunk, dirs in (("both_64MB", 64 << 20, ("h2d", "d2h")), ("both_256MB", 256 << 20, ("h2d", "d2h")), ("both_whole", 0, ("h2d", "d2h")),
                              ("h2d_only", 64 << 20, ("h2d",)), ("d2h_only", 64 << 20, ("d2h",))):
        for _ in range(2):
            step(chunk, dirs)
        barrier()
        steps = 5
        t0 = time.perf_counter()
        for _ in range(steps):
            step(chunk, dirs)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        sec = float(dt.item()) / steps
        res[name] = {"ms_per_step": 1e3 * sec,
                     "h2d_GBps_per_rank": (H2D_BYTES / sec / 1e9) if "h2d" in dirs else 0.0,
                     "d2h_GBps_per_rank": (D2H_BYTES / sec / 1e9) if "d2h" in dirs else 0.0,
                     "aggregate_GBps": world * ((H2D_BYTES if "h2d" in dirs else 0) + (D2H_BYTES if "d2h" in dirs else 0)) / sec / 1e9,
                     "frames_per_s_ceiling": world * FRAMES / sec if len(dirs) == 2 else None}
        barrier()
    if rank == 0:
        best = max(v["frames_per_s_ceiling"] for v in res.values() if v["frames_per_s_ceiling"])
        print(json.dumps({"what": "pinned-host copy ceiling of the configs[1] step (no kernel)", "n_gpus": world, "cpus": os.cpu_count(),
                          "h2d_bytes_per_step": H2D_BYTES, "d2h_bytes_per_step": D2H_BYTES, "best_frames_per_s_ceiling": best, "cases": res}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
