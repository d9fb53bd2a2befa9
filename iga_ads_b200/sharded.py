"""Multi-GPU ADS step: slabs across the ranks of one NVSwitch box, one process per GPU.

The reference has nothing to mirror here (single address space).  Design:

* The tensor is cut into contiguous slabs along its slowest axis A (z first).  The right-hand side
  (after a p-plane halo exchange with the two neighbours), the x sweep and the sweep along the
  middle axis B are slab-local.
* The sweep along A needs whole A-lines: ONE all-to-all per step turns A-slabs into B-slabs.  Its
  pack and unpack are not separate passes: the B sweep writes its result straight into the send
  blocks and the A sweep reads straight out of the receive blocks (row-offset tables of
  adsb_sweep_view), and writes the new state in the canonical layout of the NEW orientation
  (slabs along B, A in the middle).
* The next step runs in that orientation (the three axis solves commute), so orientations
  alternate z-slabs -> y-slabs -> z-slabs ... and there is exactly one exchange per step.

`SlabPlan` is pure host logic (tested on CPU with gloo, world_size 2); `ShardedHeat3d` drives the
kernels through the pointer-level C ABI and NCCL (`torch.distributed.all_to_all_single`,
`batch_isend_irecv`).
"""
import ctypes
import json
import os
import time

import numpy as np

from . import _lib
from ._lib import Form, View
from .host import dim_config, dimension
from .simulation import Context


def split(n, parts):
    """Balanced contiguous partition of range(n): (starts, sizes)."""
    sizes = [n // parts + (1 if r < n % parts else 0) for r in range(parts)]
    starts = [sum(sizes[:r]) for r in range(parts)]
    return starts, sizes


class SlabPlan:
    """Index arithmetic of the slab decomposition for one rank.

    n = (nx, ny, nz) global DOF counts; orientation A in {2, 1} is the slab (slowest) axis, B = 3 - A
    the middle axis.  Local canonical layout in orientation A: [a_local][b][x], element strides
    x: 1, B: nx, A: n[B]*nx; the halo'ed state buffer has p extra planes on both sides of A."""

    def __init__(self, n, p, world, rank):
        self.n = tuple(int(v) for v in n)
        self.p, self.world, self.rank = int(p), int(world), int(rank)
        self.starts, self.sizes = {}, {}
        for ax in (1, 2):
            self.starts[ax], self.sizes[ax] = split(self.n[ax], world)
            if min(self.sizes[ax]) < max(self.p, 1):
                raise ValueError("slabs thinner than the spline degree are not supported")
        self.cmax = {ax: max(self.sizes[ax]) for ax in (1, 2)}
        self.block = self.cmax[1] * self.cmax[2] * self.n[0]   # doubles per (src, dst) block, padded

    # ---- ownership
    def lo(self, A):
        return self.starts[A][self.rank]

    def cnt(self, A):
        return self.sizes[A][self.rank]

    def owner(self, ax, idx):
        st, sz = self.starts[ax], self.sizes[ax]
        for r in range(self.world):
            if st[r] <= idx < st[r] + sz[r]:
                return r
        raise IndexError(idx)

    def plane(self, A):
        """doubles per A-plane of the local canonical layout"""
        return self.n[3 - A] * self.n[0]

    def local_size(self, A):
        return self.cnt(A) * self.plane(A)

    def halo_size(self):
        return max((self.cmax[A] + 2 * self.p) * self.plane(A) for A in (1, 2))

    def work_size(self):
        return max(self.cmax[A] * self.plane(A) for A in (1, 2))

    # ---- exchange layout: block (src r -> dst s) = [b_local of s][a_local of r][x], padded to cmax
    def pack_offsets(self, A):
        """off_out[j] for the B sweep of orientation A writing into the send buffer; the line
        (x, a_local) adds x + a_local*nx."""
        B, nx = 3 - A, self.n[0]
        off = np.zeros(self.n[B], dtype=np.int64)
        for j in range(self.n[B]):
            s = self.owner(B, j)
            off[j] = s * self.block + (j - self.starts[B][s]) * self.cmax[A] * nx
        return off

    def unpack_offsets(self, A):
        """off_in[k] for the A sweep reading the receive buffer; the line (x, b_local) adds
        x + b_local*cmax[A]*nx."""
        nx = self.n[0]
        off = np.zeros(self.n[A], dtype=np.int64)
        for k in range(self.n[A]):
            r = self.owner(A, k)
            off[k] = r * self.block + (k - self.starts[A][r]) * nx
        return off

    def views(self, A):
        """adsb_view arguments (n[3], s[3] by global axis) of the four sweep operands."""
        B, nx = 3 - A, self.n[0]
        nloc = [nx, 0, 0]
        nloc[A], nloc[B] = self.cnt(A), self.n[B]
        s_work = [1, 0, 0]
        s_work[B], s_work[A] = nx, self.n[B] * nx
        s_send = [1, 0, 0]
        s_send[A], s_send[B] = nx, 0
        nnew = [nx, 0, 0]
        nnew[A], nnew[B] = self.n[A], self.cnt(B)
        s_recv = [1, 0, 0]
        s_recv[B], s_recv[A] = self.cmax[A] * nx, 0
        s_new = [1, 0, 0]
        s_new[A], s_new[B] = nx, self.n[A] * nx
        return dict(work=(nloc, s_work), send=(nloc, s_send), recv=(nnew, s_recv), new=(nnew, s_new))

    # ---- numpy emulation of the data movement (CPU tests): identity "sweeps"
    def emulate_pack(self, A, work):
        """work: local canonical array [cnt(A)][n[B]][nx] -> flat send buffer"""
        B, nx = 3 - A, self.n[0]
        send = np.zeros(self.world * self.block)
        off = self.pack_offsets(A)
        for a in range(self.cnt(A)):
            for j in range(self.n[B]):
                o = off[j] + a * nx
                send[o:o + nx] = work[a, j]
        return send

    def emulate_unpack(self, A, recv):
        """flat receive buffer -> local canonical array of the NEW orientation [cnt(B)][n[A]][nx]"""
        B, nx = 3 - A, self.n[0]
        out = np.zeros((self.cnt(B), self.n[A], nx))
        off = self.unpack_offsets(A)
        for b in range(self.cnt(B)):
            for k in range(self.n[A]):
                o = off[k] + b * self.cmax[A] * nx
                out[b, k] = recv[o:o + nx]
        return out


def _view(n, s):
    return View.make(n, s)


class ShardedHeat3d:
    """heat_3d (examples/heat/heat_3d.hpp) on `world` GPUs.  State lives in halo'ed buffers on the
    device; `orientation` says along which axis it is currently slabbed."""

    def __init__(self, p, elements, dt, rank, world, device):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.p, self.dt, self.rank, self.world = p, dt, rank, world
        self.dev = torch.device("cuda", device)
        n = elements + p
        self.n = (n, n, n)
        self.plan = SlabPlan(self.n, p, world, rank)
        self.dim = dimension(dim_config(p, elements))
        self.ctx = Context(self.n, device=device)
        self.ctx.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)
        lu, ipiv = self.dim.factorize_matrix()
        for ax in range(3):
            self.ctx.set_axis(ax, self.dim)
            self.ctx.set_factor(ax, 0, lu, ipiv, p, p)
        f64 = dict(dtype=torch.float64, device=self.dev)
        self.halo = [torch.zeros(self.plan.halo_size(), **f64) for _ in range(2)]
        self.work = torch.zeros(self.plan.work_size(), **f64)
        self.send = torch.zeros(world * self.plan.block, **f64)
        self.recv = torch.zeros(world * self.plan.block, **f64)
        self.cur = 0            # which halo buffer holds the state
        self.orientation = 2    # z-slabs
        self.form = Form.make(1.0, (dt, dt, dt))
        self._off = {A: (self.plan.pack_offsets(A), self.plan.unpack_offsets(A)) for A in (1, 2)}
        self.launches = 0
        self.exchange_bytes = 0

    # ---- state access (host <-> device), canonical local layout of the current orientation
    def interior(self, buf, A):
        pl = self.plan.plane(A)
        return buf[self.p * pl:(self.p + self.plan.cnt(A)) * pl]

    def set_local_state(self, host):
        """host: this rank's z-slab [cnt_z][ny][nx] (flat)"""
        self.orientation, self.cur = 2, 0
        self.interior(self.halo[0], 2).copy_(self.torch.as_tensor(host).reshape(-1), non_blocking=True)

    def local_state(self):
        """(orientation, lo, cnt, array [cnt][n_middle][nx])"""
        A = self.orientation
        arr = self.interior(self.halo[self.cur], A).cpu().numpy()
        return A, self.plan.lo(A), self.plan.cnt(A), arr.reshape(self.plan.cnt(A), self.n[3 - A], self.n[0])

    # ---- one step
    def _halo_exchange(self, buf, A):
        dist, p, pl = self.dist, self.p, self.plan.plane(A)
        c = self.plan.cnt(A)
        ops = []
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, buf[p * pl:2 * p * pl], self.rank - 1))
            ops.append(dist.P2POp(dist.irecv, buf[0:p * pl], self.rank - 1))
        if self.rank < self.world - 1:
            ops.append(dist.P2POp(dist.isend, buf[c * pl:(c + p) * pl], self.rank + 1))
            ops.append(dist.P2POp(dist.irecv, buf[(c + p) * pl:(c + 2 * p) * pl], self.rank + 1))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            self.exchange_bytes += 8 * p * pl * len(ops) // 2

    def step(self):
        plan, p, A = self.plan, self.p, self.orientation
        B, nx = 3 - A, self.n[0]
        H, Hn = self.halo[self.cur], self.halo[1 - self.cur]
        self._halo_exchange(H, A)
        # right-hand side on the slab; the input box is the slab widened by p, clipped to the domain
        a0, c = plan.lo(A), plan.cnt(A)
        in_lo_a = max(0, a0 - p)
        in_hi_a = min(self.n[A], a0 + c + p)
        pl = plan.plane(A)
        v = plan.views(A)
        nin = list(v["work"][0])
        nin[A] = in_hi_a - in_lo_a
        in_lo, out_lo = [0, 0, 0], [0, 0, 0]
        in_lo[A], out_lo[A] = in_lo_a, a0
        in_ptr = H.data_ptr() + 8 * (p - (a0 - in_lo_a)) * pl
        self.ctx.rhs_view(self.form, in_ptr, _view(nin, v["work"][1]), in_lo, self.work.data_ptr(),
                          _view(*v["work"]), out_lo)
        # x sweep in place, B sweep into the send blocks
        self.ctx.sweep_view(0, 0, self.work.data_ptr(), _view(*v["work"]), self.work.data_ptr(), _view(*v["work"]))
        self.ctx.sweep_view(B, 0, self.work.data_ptr(), _view(*v["work"]), self.send.data_ptr(), _view(*v["send"]),
                            off_out=self._off[A][0])
        if self.world > 1:
            self.dist.all_to_all_single(self.recv, self.send)
            self.exchange_bytes += 8 * plan.block * (self.world - 1)
            src = self.recv
        else:
            src = self.send
        # A sweep out of the receive blocks into the interior of the next state (orientation B)
        out_ptr = Hn.data_ptr() + 8 * p * plan.plane(B)
        self.ctx.sweep_view(A, 0, src.data_ptr(), _view(*v["recv"]), out_ptr, _view(*v["new"]),
                            off_in=self._off[A][1])
        self.launches += 4
        self.cur = 1 - self.cur
        self.orientation = B

    def advance(self, nsteps):
        for _ in range(nsteps):
            self.step()


def gather_state(sim):
    """All ranks: returns the full tensor [z][y][x] on every rank (test helper; host memory)."""
    A, lo, cnt, arr = sim.local_state()
    pieces = [None] * sim.world
    sim.dist.all_gather_object(pieces, (A, lo, cnt, arr))
    n = sim.n
    full = np.zeros((n[2], n[1], n[0]))
    for (a, l, c, x) in pieces:
        if a == 2:
            full[l:l + c] = x
        else:  # y-slabs: x is [y_local][z][x]
            full[:, l:l + c, :] = np.transpose(x, (1, 0, 2))
    return full


def run_sharded_bench(args, rank, world, local_rank):
    """bench.py leg for N > 1: strong scaling of heat_3d p=2 512^3 over z-slabs."""
    import torch
    import torch.distributed as dist

    from bench import BYTES_PER_DOF_STEP, METRIC, UNIT, ClockSampler, peaks, synthetic_local

    p, ne, dt = args.p, args.elements, 1e-7
    n = ne + p
    N = n ** 3
    sim = ShardedHeat3d(p, ne, dt, rank, world, local_rank)
    z0, cz = sim.plan.lo(2), sim.plan.cnt(2)
    u0 = synthetic_local((n, n, n), (0, 0, z0), (n, n, cz))
    host = torch.from_numpy(u0).pin_memory()
    sim.set_local_state(host)
    sim.advance(args.warmup + (args.warmup % 2))  # even count: back in z-slab orientation
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
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
    finite = bool(torch.isfinite(sim.interior(sim.halo[sim.cur], sim.orientation)).all().item())
    timed_launches, timed_exchange = sim.launches, sim.exchange_bytes

    # end to end: upload the slab from pinned host memory, one step, download the slab, every step
    k2 = 2
    out_host = torch.empty(sim.plan.halo_size(), dtype=torch.float64).pin_memory()
    sim.set_local_state(host)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(k2):
        A = sim.orientation
        if A == 2:
            sim.interior(sim.halo[sim.cur], 2).copy_(host, non_blocking=True)
        else:  # y-slab orientation: same byte count
            sim.interior(sim.halo[sim.cur], 1).copy_(host[:sim.plan.local_size(1)], non_blocking=True)
        sim.step()
        loc = sim.interior(sim.halo[sim.cur], sim.orientation)
        out_host[:loc.numel()].copy_(loc, non_blocking=True)
        torch.cuda.synchronize()
    dist.barrier()
    el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=sim.dev)
    dist.all_reduce(el, op=dist.ReduceOp.MAX)
    el = float(el.item())
    if rank == 0:
        hbm, peak_kind = peaks()
        step_s = ms * 1e-3 / args.steps
        print(json.dumps({
            "metric": METRIC, "value": N * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"heat_3d p={p} {ne}^3 elements ({N} DOF), explicit ADS step, dt={dt}",
                       "rhs": "collapsed (pre-integrated sum factorisation)",
                       "l2": f"state {8 * N / world / 1e6:.0f} MB per GPU vs 126 MB L2: inputs exceed L2 for N <= 8",
                       "parallelism": f"{world} slabs, one all-to-all (NCCL) + p-plane halo per step, orientation alternates"},
            "roofline": {"bound": "hbm", "kernel": "whole step, per GPU", "achieved": BYTES_PER_DOF_STEP * N / world / step_s / 1e9,
                         "peak": hbm, "peak_kind": peak_kind, "unit": "GB/s",
                         "frac": BYTES_PER_DOF_STEP * N / world / step_s / 1e9 / hbm, "traffic": None,
                         "exchange_bytes_per_gpu_step": timed_exchange / max(args.steps, 1)},
            "clocks": clocks, "gpu_launches": timed_launches * world, "finite": finite,
            "e2e": {"value": N * k2 / el, "unit": UNIT, "h2d_bytes_per_step": 8 * N, "d2h_bytes_per_step": 8 * N,
                    "steps": k2, "ms_per_step": 1e3 * el / k2,
                    "note": "each rank: pinned slab upload + step + slab download per step"},
        }))
    dist.barrier()
    dist.destroy_process_group()
