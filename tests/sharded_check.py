"""Multi-rank parity check of the slab-sharded ADS step (run under torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/sharded_check.py

Every rank steps its slab; the gathered state is compared on rank 0 with the CPU oracle after 1, 2
and 3 steps (both slab orientations are exercised).  Prints SHARDED_CHECK_OK on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from iga_ads_b200.sharded import ShardedHeat3d, gather_state  # noqa: E402
from oracle.oracle import Oracle, rel_l2, synthetic_state  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    # every exchange mode: copy-engine pushes behind the next chunk's compute (default; one chunk and forced
    # 4-plane chunks), peer stores fused into the sweep (one chunk, and chunked on a second stream), NCCL
    cases = [(2, 30, 1e-7, "ce", "1", "64"), (3, 21, 1e-7, "ce", "4", "4"), (2, 30, 1e-7, "ce", "4", "4"),
             (2, 30, 1e-7, "p2p", "1", "64"), (3, 21, 1e-7, "p2p", "4", "4"), (2, 30, 1e-7, "nccl", "1", "64")]
    if os.environ.get("ADSB_CHECK_QUICK"):  # default exchange only (large rank counts: keep the box time short)
        cases = cases[:2]
    if os.environ.get("ADSB_CHECK_CASES"):  # e.g. "3,4,5"
        cases = [cases[int(k)] for k in os.environ["ADSB_CHECK_CASES"].split(",")]
    for p, ne, dt, mode, chunks, min_planes in cases:
        os.environ["ADSB_SHARDED_EXCHANGE"] = mode
        os.environ["ADSB_SHARDED_CHUNKS"] = chunks
        os.environ["ADSB_SHARDED_MIN_PLANES"] = min_planes
        n = ne + p
        u0 = synthetic_state((n, n, n))                      # memory order: x fastest
        sim = ShardedHeat3d(p, ne, dt, rank, world, local)
        if rank == 0:
            print(f"exchange = {sim.exchange}, chunks = {sim.nchunk}", flush=True)
        z0, cz = sim.plan.lo(2), sim.plan.cnt(2)
        sim.set_local_state(u0.reshape(n, n, n)[z0:z0 + cz].copy())
        for steps in (1, 2, 3):
            sim.step()
            got = gather_state(sim)
            if rank == 0:
                want, _ = Oracle().run("heat_3d", p, ne, dt, steps, u0=u0)
                err = rel_l2(got.ravel(), want)
                print(f"p={p} n={ne}^3 world={world} steps={steps} rel L2 vs oracle = {err:.2e}", flush=True)
                ok = ok and err < steps * 1e-12
        if sim.use_graph:
            # the captured two-step graph: 2 eager steps (warm caches), then 4 steps as 2 replays
            sim.set_local_state(u0.reshape(n, n, n)[z0:z0 + cz].copy())
            sim.advance(2)
            sim.advance(4)
            got = gather_state(sim)
            if rank == 0:
                want, _ = Oracle().run("heat_3d", p, ne, dt, 6, u0=u0)
                err = rel_l2(got.ravel(), want)
                print(f"p={p} n={ne}^3 world={world} 6 steps, graph captured = {sim.graph is not None}: "
                      f"rel L2 vs oracle = {err:.2e}", flush=True)
                ok = ok and sim.graph is not None and err < 6e-12
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and flag.item() == 1.0:
        print("SHARDED_CHECK_OK", flush=True)
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
