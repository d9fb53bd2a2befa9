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


def fused_sweep_check(rank, world, local, p=2, rows=48, nx=64, ny=36):
    """the fused one-kernel distributed z sweep over real peer memory: every rank owns `rows` planes of a
    (nx, ny, rows * world) tensor; the result must equal dgbtrs over the whole z lines (oracle, rank 0)"""
    import iga_ads_b200 as ads
    from iga_ads_b200 import host
    from iga_ads_b200._lib import DistArgs, View, fill_sentinel
    import torch.distributed._symmetric_memory as symm

    dev = torch.device("cuda", local)
    nz = rows * world
    lu, piv = ads.band_factorize(ads.matrix_1d(0, p, nz - p), p, p)
    bounds = host.segment_bounds(piv, p, world)
    ctx = ads.Context((nx, ny, nz), device=local)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    ctx.set_factor(2, 0, lu, piv, p, p)
    ctx.set_segments(2, 0, bounds, rank, 1)
    info = ctx.segment_info(2, 0)
    lines = nx * ny
    z0, cz = int(bounds[rank]), int(bounds[rank + 1] - bounds[rank])
    nd, nxs = world * info["KL"] * lines, world * info["KD"] * lines
    buf = symm.empty(nd + nxs, dtype=torch.float64, device=dev)
    hdl = symm.rendezvous(buf, dist.group.WORLD)
    fill_sentinel(buf)
    hdl.barrier(channel=0)
    base = [int(v) for v in hdl.buffer_ptrs]
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    good = True
    for launch in range(3):
        rhs = np.random.default_rng(100 + launch).standard_normal((nz, ny, nx))
        slab = torch.from_numpy(rhs[z0:z0 + cz].copy()).to(dev).reshape(-1)
        a = DistArgs()
        a.rank, a.nranks, a.nl, a.lag = rank, world, 32, 4
        a.dseg_local, a.x_local = base[rank], base[rank] + 8 * nd
        if rank + 1 < world:
            a.dseg_next = base[rank + 1]
        if rank > 0:
            a.x_prev = base[rank - 1] + 8 * nd
        a.error_flag = err.data_ptr()
        v = View.make([nx, ny, cz], [1, nx, lines])
        assert ctx.dist_sweep_check(2, 0, rank, v, 32, 4)
        ctx.dist_sweep_view(2, 0, slab.data_ptr(), v, a)
        torch.cuda.synchronize()
        hdl.barrier(channel=0)   # consecutive sweeps on the same state arrays are separated by a barrier
        pieces = [None] * world
        dist.all_gather_object(pieces, (z0, slab.cpu().numpy().reshape(cz, ny, nx)))
        if rank == 0:
            got = np.concatenate([a_ for _, a_ in sorted(pieces, key=lambda t: t[0])])
            zl = np.ascontiguousarray(np.moveaxis(rhs, 0, -1))
            want = np.moveaxis(Oracle().solve_factorized(lu, piv, p, p, zl).reshape(zl.shape), -1, 0)
            e = rel_l2(got.ravel(), want.ravel())
            print(f"fused distributed z sweep, world={world}, launch {launch}: rel L2 vs dgbtrs = {e:.2e}, "
                  f"timeouts = {int(err.item())}", flush=True)
            good = good and e < 1e-13
    good = good and int(err.item()) == 0
    return good


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True if os.environ.get("ADSB_CHECK_CASES") else fused_sweep_check(rank, world, local)
    cases = [("heat_3d", 2, 30, 1e-7), ("heat_3d", 3, 8 * world + 5, 1e-7), ("implicit_3d", 3, 30, 1e-2),
             ("scalability_3d", 2, 30, 1e-6), ("scalability_3d", 5, 12 * world, 1e-6)]
    if world == 2:
        cases.append(("heat_3d", 2, 94, 1e-7))   # 48-plane slabs: the fused one-kernel z sweep inside whole steps
    if os.environ.get("ADSB_CHECK_QUICK"):
        cases = cases[:2]
    if os.environ.get("ADSB_CHECK_CASES"):  # e.g. "4,5"
        cases = [cases[int(k)] for k in os.environ["ADSB_CHECK_CASES"].split(",")]
    for problem, p, ne, dt in cases:
        n = ne + p
        u0 = synthetic_state((n, n, n))
        sim = SlabSim(problem, p, ne, dt, rank, world, local)
        sim.set_local_state(u0.reshape(n, n, n)[sim.z0:sim.z0 + sim.cz])
        sim.publish()
        for steps in (1, 2):
            dist.barrier()   # rank 0 was busy with the oracle: enter the step together
            sim.step()
            got = gather_state(sim)
            if rank == 0:
                want, _ = Oracle().run(problem, p, ne, dt, steps, u0=u0)
                err = rel_l2(got.ravel(), want)
                tol = steps * (1e-12 if p <= 3 else 1e-10)
                print(f"{problem} p={p} n={ne}^3 world={world} steps={steps} fused={sim.fused} rel L2 vs oracle = {err:.2e}",
                      flush=True)
                ok = ok and err < tol
        # the captured graph: restart, 2 eager steps (the loop above already warmed every cache), then 4 steps
        # as 2 replays of the captured pair of steps
        dist.barrier()
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
        ok = ok and int(sim.err_flag.item()) == 0   # no boundary-value poll timed out
        del sim
        import gc

        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and flag.item() == 1.0:
        print("SLAB_CHECK_OK", flush=True)
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
