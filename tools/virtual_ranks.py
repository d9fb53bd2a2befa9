#!/usr/bin/env python
"""Single-GPU emulation of the slab-sharded step: `world` virtual ranks run one after the other on one
device (same SlabPlan, same pointer-level kernels and offset tables as ShardedHeat3d's NCCL path; the
all-to-all is a block shuffle with torch).  Compares with the oracle after every step.
    python tools/virtual_ranks.py [world] [p] [ne]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iga_ads_b200._lib import Form, View  # noqa: E402
from iga_ads_b200.host import dim_config, dimension  # noqa: E402
from iga_ads_b200.sharded import SlabPlan  # noqa: E402
from iga_ads_b200.simulation import Context  # noqa: E402
from oracle.oracle import Oracle, rel_l2, synthetic_state  # noqa: E402


def run(world, p, ne, steps=2, dt=1e-7):
    n = ne + p
    N3 = (n, n, n)
    dev = torch.device("cuda", 0)
    dim = dimension(dim_config(p, ne))
    ctx = Context(N3, device=0)
    lu, ipiv = dim.factorize_matrix()
    for ax in range(3):
        ctx.set_axis(ax, dim)
        ctx.set_factor(ax, 0, lu, ipiv, p, p)
    form = Form.make(1.0, (dt, dt, dt))
    plans = [SlabPlan(N3, p, world, r) for r in range(world)]
    u0 = synthetic_state(N3)
    full = u0.reshape(n, n, n).copy()          # [z][y][x]; orientation A: slabs along A
    A = 2
    worst = 0.0
    for st in range(1, steps + 1):
        B, nx = 3 - A, n
        sends = []
        for r, pl in enumerate(plans):
            a0, c = pl.lo(A), pl.cnt(A)
            lo, hi = max(0, a0 - p), min(n, a0 + c + p)
            slab = full[lo:hi] if A == 2 else np.transpose(full[:, lo:hi, :], (1, 0, 2))
            src = torch.from_numpy(np.ascontiguousarray(slab)).to(dev)
            work = torch.zeros(c * n * n, dtype=torch.float64, device=dev)
            send = torch.zeros(world * pl.block, dtype=torch.float64, device=dev)
            v = pl.views(A)
            nin = list(v["work"][0]); nin[A] = hi - lo
            in_lo, out_lo = [0, 0, 0], [0, 0, 0]
            in_lo[A], out_lo[A] = lo, a0
            ctx.rhs_view(form, src.data_ptr(), View.make(nin, v["work"][1]), in_lo, work.data_ptr(), View.make(*v["work"]), out_lo)
            ctx.sweep_view(0, 0, work.data_ptr(), View.make(*v["work"]), work.data_ptr(), View.make(*v["work"]))
            ctx.sweep_view(B, 0, work.data_ptr(), View.make(*v["work"]), send.data_ptr(), View.make(*v["send"]),
                           off_out=pl.pack_offsets(A))
            ctx.synchronize()
            sends.append(send)
        new = np.zeros_like(full)
        for s, pl in enumerate(plans):
            blk = pl.block
            recv = torch.cat([sends[r][s * blk:(s + 1) * blk] for r in range(world)])
            v = pl.views(A)
            out = torch.zeros(pl.cnt(B) * n * n, dtype=torch.float64, device=dev)
            ctx.sweep_view(A, 0, recv.data_ptr(), View.make(*v["recv"]), out.data_ptr(), View.make(*v["new"]),
                           off_in=pl.unpack_offsets(A))
            ctx.synchronize()
            arr = out.cpu().numpy().reshape(pl.cnt(B), n, n)      # [b_local][a][x]
            b0 = pl.lo(B)
            if B == 1:   # new slabs along y: arr[y_local][z][x]
                new[:, b0:b0 + pl.cnt(B), :] = np.transpose(arr, (1, 0, 2))
            else:        # new slabs along z: arr[z_local][y][x]
                new[b0:b0 + pl.cnt(B)] = arr
        full = new
        A = B
        want, _ = Oracle().run("heat_3d", p, ne, dt, st, u0=u0)
        err = rel_l2(full.ravel(), want)
        worst = max(worst, err / st)
        print(f"virtual ranks: world={world} p={p} n={ne}^3 step {st}: rel L2 vs oracle = {err:.2e}", flush=True)
    return worst


if __name__ == "__main__":
    w = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    ne = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    sys.exit(0 if run(w, p, ne) < 1e-12 else 1)
