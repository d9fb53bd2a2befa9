"""Multi-rank parity check of the z-slab path with the distributed z substitution (iga_ads_b200/slab.py);
run under torchrun on >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29534 tests/slab_check.py

Every rank steps its slab (symmetric memory, peer stores, signal barriers -- eagerly and through the captured
CUDA graph); the gathered state is compared on rank 0 with the CPU oracle.  Prints SLAB_CHECK_OK on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from iga_ads_b200.slab import SlabSim, gather_state  # noqa: E402
from oracle.oracle import Oracle, rel_l2, synthetic_state  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    cases = [("heat_3d", 2, 30, 1e-7), ("heat_3d", 3, 8 * world + 5, 1e-7), ("implicit_3d", 3, 30, 1e-2),
             ("scalability_3d", 2, 30, 1e-6), ("scalability_3d", 5, 12 * world, 1e-6)]
    if os.environ.get("ADSB_CHECK_QUICK"):
        cases = cases[:2]
    for problem, p, ne, dt in cases:
        n = ne + p
        u0 = synthetic_state((n, n, n))
        sim = SlabSim(problem, p, ne, dt, rank, world, local)
        sim.set_local_state(u0.reshape(n, n, n)[sim.z0:sim.z0 + sim.cz])
        sim.publish()
        for steps in (1, 2):
            sim.step()
            got = gather_state(sim)
            if rank == 0:
                want, _ = Oracle().run(problem, p, ne, dt, steps, u0=u0)
                err = rel_l2(got.ravel(), want)
                tol = steps * (1e-12 if p <= 3 else 1e-10)
                print(f"{problem} p={p} n={ne}^3 world={world} steps={steps} seg={sim.seg} rel L2 vs oracle = {err:.2e}",
                      flush=True)
                ok = ok and err < tol
        # the captured graph: restart, 2 eager steps (the loop above already warmed every cache), then 4 steps
        # as 2 replays of the captured pair of steps
        sim.set_local_state(u0.reshape(n, n, n)[sim.z0:sim.z0 + sim.cz])
        sim.publish()
        sim.advance(2, graph=False)
        sim.advance(4, graph=True)
        got = gather_state(sim)
        if rank == 0:
            want, _ = Oracle().run(problem, p, ne, dt, 6, u0=u0)
            err = rel_l2(got.ravel(), want)
            print(f"{problem} p={p}: 6 steps, graph captured = {sim.graph is not None}: rel L2 vs oracle = {err:.2e}",
                  flush=True)
            ok = ok and sim.graph is not None and err < 6 * (1e-12 if p <= 3 else 1e-10)
        del sim
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and flag.item() == 1.0:
        print("SLAB_CHECK_OK", flush=True)
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
