import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
from iga_ads_b200.sharded import ShardedHeat3d, gather_state
from oracle.oracle import Oracle, rel_l2, synthetic_state
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
p, ne, dt = 2, 30, 1e-7
n = ne + p
u0 = synthetic_state((n, n, n))
want = {}
if rank == 0:
    for st in (1, 2): want[st], _ = Oracle().run("heat_3d", p, ne, dt, st, u0=u0)
os.environ["ADSB_SHARDED_MIN_PLANES"] = "64"
for mode in ("ce", "p2p", "nccl"):
    os.environ["ADSB_SHARDED_EXCHANGE"] = mode
    sim = ShardedHeat3d(p, ne, dt, rank, world, local)
    z0, cz = sim.plan.lo(2), sim.plan.cnt(2)
    sim.set_local_state(u0.reshape(n, n, n)[z0:z0 + cz].copy())
    for st in (1, 2):
        sim.step()
        got = gather_state(sim)
        if rank == 0:
            e = rel_l2(got.ravel(), want[st])
            # which planes are wrong
            g = got.reshape(n, n, n); w = want[st].reshape(n, n, n)
            bad_z = [int(k) for k in range(n) if np.abs(g[k] - w[k]).max() > 1e-9 * np.abs(w).max()]
            bad_y = [int(k) for k in range(n) if np.abs(g[:, k] - w[:, k]).max() > 1e-9 * np.abs(w).max()]
            print(f"{os.environ.get('TAG','')} mode={mode} world={world} step {st}: err {e:.2e} bad z {bad_z[:40]} bad y {bad_y[:40]}", flush=True)
dist.barrier(); dist.destroy_process_group()
