"""bench.py's leg for N > 1 GPUs: strong scaling of a 3-D BASELINE configuration over z-slabs.

Default path: iga_ads_b200/slab.py (distributed z substitution, nothing transposed).  When the z factor
cannot be cut into slabs (ADSB_EINVAL from adsb_set_axis_segments on any rank) or ADSB_MULTI=transpose,
heat_3d falls back to the transposing exchange of iga_ads_b200/sharded.py.

Before timing, the same code path runs a small problem (62 elements per axis, 36 / 43 at p = 4 / 5; two steps) and rank 0 compares
the gathered state with the CPU oracle: the `parity` record of the JSON line.  After the timed steps the sum
and the norm of the full-size state are reduced over the ranks (`checksum`): they must agree with the 1-GPU
line of the same command (same synthetic state, same number of steps)."""
import json
import os
import time

import numpy as np


def _parity(problem, p, dt, rank, world, local_rank, ne=None):
    """two steps of the sharded path on a small grid against the oracle (rank 0 holds the verdict); the grid shrinks
    with the degree because the oracle's element loop costs (p + 1)^6 per element"""
    if ne is None:
        ne = {4: 36, 5: 43}.get(p, 62)
    import torch
    import torch.distributed as dist

    from oracle.oracle import Oracle, rel_l2, synthetic_state

    from .slab import SlabSim, gather_state

    n = ne + p
    u0 = synthetic_state((n, n, n))
    sim, why = None, ""
    try:
        sim = SlabSim(problem, p, ne, dt, rank, world, local_rank)
    except Exception as e:  # noqa: BLE001 -- e.g. slabs of the small grid too thin to cut at this degree / rank count
        why = str(e)
    ok = torch.tensor([1 if sim is not None else 0], dtype=torch.int32, device=torch.device("cuda", local_rank))
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 0:
        del sim
        return {"problem": problem, "p": p, "elements": ne, "world": world,
                "skipped": "the parity grid cannot be cut into this many slabs at this degree"
                           + (f" ({why})" if why else "")}
    sim.set_local_state(u0.reshape(n, n, n)[sim.z0:sim.z0 + sim.cz])
    sim.publish()
    out = {"problem": problem, "p": p, "elements": ne, "dof": n ** 3, "world": world, "steps": 2}
    sim.step()
    sim.step()
    got = gather_state(sim)
    if rank == 0:
        want, _ = Oracle().run(problem, p, ne, dt, 2, u0=u0)
        out["rel_l2_vs_oracle"] = rel_l2(got.ravel(), want)
    torch.cuda.synchronize()
    dist.barrier()
    del sim
    return out


def run_multi_gpu_bench(args, cfg, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from bench import UNIT, ClockSampler, metric_name, peaks, synthetic_local

    from .slab import SlabSim

    p, ne, dt, problem = cfg["p"], cfg["elements"], cfg["dt"], cfg["problem"]
    n = ne + p
    N = n ** 3
    mode = os.environ.get("ADSB_MULTI", "slab")
    sim, err = None, None
    if mode == "slab":
        try:
            sim = SlabSim(problem, p, ne, dt, rank, world, local_rank)
        except Exception as e:  # noqa: BLE001 -- every rank must take the same path: agree below
            err = e
    ok = torch.tensor([1 if sim is not None else 0], dtype=torch.int32, device=torch.device("cuda", local_rank))
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 0:
        if problem != "heat_3d":
            raise SystemExit(f"slab path unavailable for {problem}: {err}")
        if rank == 0 and mode == "slab":
            print(f"[adsb] slab path unavailable ({err}); using the transposing exchange", flush=True)
        from .sharded import run_sharded_bench

        args.p, args.elements = p, ne
        run_sharded_bench(args, rank, world, local_rank)
        return

    parity = None
    if not args.no_parity:
        parity = _parity(problem, p, dt, rank, world, local_rank)

    u0 = synthetic_local((n, n, n), (0, 0, sim.z0), (n, n, sim.cz))
    host = torch.from_numpy(u0).pin_memory()
    sim.set_local_state(host.numpy())
    sim.publish()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sim.advance(args.warmup)
    torch.cuda.synchronize()
    dist.barrier()
    # clocks under load: keep stepping ~0.3 s (untimed) while nvidia-smi samples, then restore the state
    sampler.mark()
    for _ in range(60 if world >= 4 else 30):
        sim.advance(10)
    torch.cuda.synchronize()
    dist.barrier()
    sim.cur = 0
    sim.set_local_state(host.numpy())
    sim.publish()
    sim.advance(args.warmup + (args.warmup & 1))
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sim.launches = sim.exchange_bytes = 0
    e0.record()
    sim.advance(args.steps)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=sim.dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    timed_launches, timed_exchange = sim.launches, sim.exchange_bytes
    loc = sim.interior(sim.cur).view(sim.cz, n, sim.pitch)[:, :, :n]
    red = torch.stack([loc.sum(), (loc * loc).sum(), torch.isfinite(loc).all().to(torch.float64)])
    fin = red[2:3].clone()
    dist.all_reduce(red[:2], op=dist.ReduceOp.SUM)
    dist.all_reduce(fin, op=dist.ReduceOp.MIN)
    warm = args.warmup + (args.warmup & 1)
    checksum = {"steps": warm + args.steps, "sum": float(red[0].item()), "l2": float(red[1].sqrt().item())}
    errf = sim.err_flag.clone().to(torch.float64)
    dist.all_reduce(errf, op=dist.ReduceOp.MAX)
    finite = bool(fin.item() == 1.0)
    # per-phase device times of rank 0 (separate eager pass, events between the launches)
    sim.timing = True
    sim.advance(args.steps, graph=False)
    phases = {k: v / args.steps for k, v in sim.phase_times().items()}
    sim.timing = False

    # end to end: upload the slab from pinned host memory, one step, download the slab, every step
    e2e = None
    if not args.no_e2e:
        k2 = 3
        out_host = torch.empty(sim.cz * n * n, dtype=torch.float64).pin_memory()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(k2):
            sim.set_local_state(host.numpy())
            sim.publish()
            sim.step()
            dst = out_host.view(sim.cz, n, n)
            dst.copy_(sim.interior(sim.cur).view(sim.cz, n, sim.pitch)[:, :, :n], non_blocking=True)
            torch.cuda.synchronize()
        dist.barrier()
        el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=sim.dev)
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        el = float(el.item())
        e2e = {"value": N * k2 / el, "unit": UNIT, "h2d_bytes_per_step": 8 * N, "d2h_bytes_per_step": 8 * N,
               "steps": k2, "ms_per_step": 1e3 * el / k2,
               "note": "each rank: pinned slab upload + halo publish + step + slab download per step"}
    if rank == 0:
        hbm, peak_kind = peaks()
        step_s = ms * 1e-3 / args.steps
        seg = sim.seg[sim.zslots[0]]
        achieved = cfg["bytes"] * N / world / step_s / 1e9
        line = {
            "metric": metric_name(cfg), "value": N * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{problem} p={p} {ne}^3 elements ({N} DOF), ADS step ({cfg['nsub']} sub-step"
                                   f"{'s' if cfg['nsub'] > 1 else ''}), dt={dt}",
                       "rhs": "collapsed (pre-integrated sum factorisation)",
                       "l2": f"state {8 * N / world / 1e6:.0f} MB per GPU vs 126 MB L2",
                       "parallelism": f"{world} z-slabs {list(int(v) for v in np.diff(sim.bounds))}; x, y sweeps and RHS "
                                      "slab-local; z sweep = distributed substitution (pass A, boundary values to the "
                                      "neighbours, pass B); p halo planes per step; nothing is transposed",
                       "exchange": f"{seg['KL']} + {seg['KD']} doubles per z line and rank boundary (chain depth "
                                   f"{seg['DF']}/{seg['DB']}) stored by the kernels through peer pointers (symmetric "
                                   "memory), halo planes by the copy engines; " +
                                   (f"ONE fused kernel per rank for the z sweep (pass A, flag-ordered neighbour exchange, "
                                    f"pass B; {sim.nl} lines per tile, lag {sim.lag}), 1 signal barrier per sub-step"
                                    if sim.fused else "pass A / boundary kernels / pass B, 3 signal barriers per sub-step"),
                       "cuda_graph": bool(sim.graph is not None)},
            "roofline": {"bound": "hbm", "kernel": "whole step, per GPU (64 B/DOF algorithmic; the distributed z sweep "
                                                   "itself moves 32 B/DOF)", "achieved": achieved,
                         "peak": hbm, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / hbm, "traffic": None,
                         "exchange_bytes_per_gpu_step": timed_exchange / max(args.steps, 1),
                         "phase_ms_rank0": phases},
            "clocks": clocks, "gpu_launches": timed_launches * world, "finite": finite, "checksum": checksum,
            "parity": parity, "flag_wait_timeouts": bool(errf.item() > 0),
        }
        if e2e:
            line["e2e"] = e2e
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
