#!/usr/bin/env python
"""Per-rank kernel times of the slab-sharded step on ONE GPU: rank `r` of a `world`-rank run executes alone
(peer stores land in its own arrays), CUDA events between the phases.  Shows what a rank's kernels cost
without the exchange, i.e. the compute part of the multi-GPU step.
    python tools/slab_rank_time.py [world] [rank] [problem] [p] [elements] [steps]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iga_ads_b200.slab import SlabSim, _SelfPeers  # noqa: E402


def run(world, rank, problem="heat_3d", p=2, ne=512, steps=6):
    sim = SlabSim(problem, p, ne, 1e-7 if problem == "heat_3d" else 1e-6, rank, world, 0, peers=_SelfPeers())
    if sim.fused_ok_locally():   # loop-back peers: the rank's own state arrays stand in for its neighbours'
        sim.enable_fused()
    rng = np.random.default_rng(0)
    sim.set_local_state(rng.standard_normal(sim.cz * sim.n[1] * sim.n[0]))
    for _ in range(3):
        sim.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sim.step()
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / steps
    sim.timing = True
    for _ in range(steps):
        sim.step()
    ph = {k: v / steps for k, v in sim.phase_times().items()}
    sim.timing = False
    # the same through a CUDA graph (launch gaps removed)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        sim.ctx.set_stream(s.cuda_stream)
        sim.step()
        sim.step()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            sim.step()
            sim.step()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    graph_ms = e0.elapsed_time(e1) / steps / 2
    n = sim.n
    out = {"world": world, "rank": rank, "problem": problem, "p": p, "elements": ne, "planes": sim.cz,
           "dof_rank": n[0] * n[1] * sim.cz, "ms_per_step_eager": total, "ms_per_step_graph": graph_ms,
           "phases_ms": ph, "fused": sim.fused, "nl": sim.nl, "lag": sim.lag, "flag_timeout": int(sim.err_flag.item())}
    print(json.dumps(out), flush=True)
    return out


if __name__ == "__main__":
    a = sys.argv[1:]
    run(int(a[0]) if a else 8, int(a[1]) if len(a) > 1 else 1, a[2] if len(a) > 2 else "heat_3d",
        int(a[3]) if len(a) > 3 else 2, int(a[4]) if len(a) > 4 else 512, int(a[5]) if len(a) > 5 else 6)
