"""2-rank debug of the fused slab path: after every step, per rank: NaNs per plane, leftover non-sentinel words in
the state arrays, poll time-outs, first plane that differs from the oracle."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iga_ads_b200.slab import SlabSim
from iga_ads_b200._lib import DIST_SENTINEL_WORD
from oracle.oracle import Oracle, synthetic_state

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
problem, p, ne, dt = "heat_3d", 2, int(os.environ.get("NE", "94")), 1e-7
n = ne + p
u0 = synthetic_state((n, n, n))
sim = SlabSim(problem, p, ne, dt, rank, world, local)
sim.set_local_state(u0.reshape(n, n, n)[sim.z0:sim.z0 + sim.cz])
sim.publish()
sent = (DIST_SENTINEL_WORD << 32) | DIST_SENTINEL_WORD
for step in (1, 2, 3):
    sim.step()
    torch.cuda.synchronize()
    dist.barrier()
    z0, a = sim.local_state()
    want = Oracle().run(problem, p, ne, dt, step, u0=u0)[0].reshape(n, n, n)[z0:z0 + a.shape[0]]
    bad = [k for k in range(a.shape[0]) if not np.isfinite(a[k]).all()]
    err = [float(np.abs(a[k] - want[k]).max()) for k in range(a.shape[0])]
    left = {nm: int((sim.sym[nm].view(torch.int64) != sent).sum().item()) for nm in ("dseg", "x")}
    H = sim.halo(sim.cur).view(-1, n, sim.pitch)
    hal = [bool(torch.isfinite(H[k]).all().item()) for k in range(H.shape[0])]
    for r in range(world):
        if r == rank:
            print(f"rank {rank} step {step}: fused={sim.fused} nl={sim.nl} timeouts={int(sim.err_flag.item())} nan planes={bad[:6]}.. ({len(bad)}) "
                  f"max err per plane first/last={err[0]:.2e}/{err[-1]:.2e} worst={max(err):.2e} at {int(np.argmax(err))} "
                  f"leftover non-sentinel words={left} halo buffer planes finite={hal[:3]}..{hal[-3:]}", flush=True)
        dist.barrier()
dist.destroy_process_group()
