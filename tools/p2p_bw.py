#!/usr/bin/env python
"""NVLink peer-copy bandwidth between the ranks of one box through symmetric memory (torchrun):
copy-engine copies (tensor.copy_ of contiguous slices) one way and all ranks at once.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/p2p_bw.py"""
import os

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 34 * 1024 * 1024 * (world - 1)  # doubles: 272 MB per peer pair at world 2
    buf = symm.empty(n, dtype=torch.float64, device=dev)
    hdl = symm.rendezvous(buf, dist.group.WORLD)
    src = torch.randn(n, dtype=torch.float64, device=dev)
    per = n // (world - 1)

    def run(active, label):
        hdl.barrier(channel=0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            if active:
                k = 0
                for peer in range(world):
                    if peer == rank:
                        continue
                    dst = hdl.get_buffer(peer, (per,), torch.float64, k * per if world == 2 else (rank - (rank > peer)) * per)
                    dst.copy_(src[k * per:(k + 1) * per], non_blocking=True)
                    k += 1
        e1.record()
        torch.cuda.synchronize()
        hdl.barrier(channel=0)
        ms = e0.elapsed_time(e1) / reps
        if active:
            print(f"[{label}] rank {rank}: {8 * n / 1e6:.0f} MB out in {ms:.3f} ms = {8 * n / ms / 1e6:.0f} GB/s", flush=True)

    run(rank == 0, "one way, copy engine")
    run(True, "all ranks at once, copy engine")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
