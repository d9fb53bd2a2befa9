"""Multi-GPU ADS step on z-slabs with a DISTRIBUTED z substitution: nothing is transposed.

One process per GPU.  Rank r owns the planes [bounds[r], bounds[r+1]) of the canonical (x fastest)
tensor.  Per sub-step (same sequence as simulation.py / adsb_step):

    right-hand side on the slab (p halo planes of u_prev from the two neighbours)
    x sweep, y sweep                               slab-local
    z sweep = segmented substitution, the slab being one segment of every z line:
        pass A   the slab's own columns of the factor, zero incoming states    (local sweep kernel)
        dseg     KL forward boundary values per line  -> stored into the next ranks' state arrays
        din / X  KD backward boundary values per line -> stored into the previous ranks' state arrays
        pass B   x = xhat + Xi din + Psi tin          -> interior of the next state buffer
    boundary planes of the new state -> the neighbours' halo regions

Only KL + KD doubles per z line and rank boundary cross NVLink (8.4 MB per rank at 514^2 lines, p = 2)
instead of the 119 MB of a slab-to-slab transpose; the price is pass B, one more streaming pass over the
slab.  The state arrays and the halo'ed state buffers live in symmetric memory
(torch.distributed._symmetric_memory): kernels and copy engines store straight through peer pointers,
three signal-pad barriers per sub-step order them.  See iga_ads_b200/sharded.py for the transposing
variant (kept for factors whose segments cannot be cut: growing boundary responses).

`SlabSim` is one rank; `VirtualCluster` runs several ranks in lockstep on ONE device (same kernels,
same pointer plumbing, peers are plain tensors) -- that is how the path is tested on a single GPU.
"""
import os

import numpy as np

from . import _lib
from ._lib import Form, View
from .host import segment_bounds
from .simulation import FORCING, PROBLEMS, Context, timesteps_config


class _LocalPeers:
    """Peer access for virtual ranks on one device: every rank's arrays are ordinary tensors."""

    def __init__(self):
        self.ranks = []

    def ptr(self, rank, name):
        return self.ranks[rank].sym[name].data_ptr()

    def tensor(self, rank, name):
        return self.ranks[rank].sym[name]

    def barrier(self, channel):
        pass  # lockstep execution (VirtualCluster) orders the phases


class _SelfPeers(_LocalPeers):
    """Timing vehicle: one rank of a `world`-rank run alone on a device; every peer store lands in the rank's
    own arrays (results are meaningless, the work per kernel is that of a real rank)."""

    def ptr(self, rank, name):
        me = self.ranks[0]
        base = me.sym[name].data_ptr()
        # loop the exchange of the fused sweep back: what the rank stores for its neighbours (its own slot of their
        # state arrays) must land where it expects THEIR values (slots rank -/+ 1 of its own arrays)
        if name == "dseg":
            return base - 8 * me.KL * me.lines
        if name == "x":
            return base + 8 * me.KD * me.lines
        return base

    def tensor(self, rank, name):
        return self.ranks[0].sym[name]


class _SymmPeers:
    """Peer access through one symmetric allocation per rank: [array 0 | array 1 | ...] in a fixed order."""

    def __init__(self, sim, sizes):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        self.sim, self.torch = sim, torch
        self.offsets, off = {}, 0
        for name, count in sizes:
            self.offsets[name] = (off, count)
            off += count + (count & 1)
        self.buf = symm.empty(off, dtype=torch.float64, device=sim.dev)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, dist.group.WORLD)
        self.base = [int(v) for v in self.hdl.buffer_ptrs]
        self.total = off
        self.hdl.barrier(channel=0)

    def local(self, name):
        o, c = self.offsets[name]
        return self.buf[o:o + c]

    def ptr(self, rank, name):
        return self.base[rank] + 8 * self.offsets[name][0]

    def tensor(self, rank, name):
        o, c = self.offsets[name]
        return self.hdl.get_buffer(rank, (c,), self.torch.float64, o)

    def barrier(self, channel):
        self.hdl.barrier(channel=channel)


class SlabSim:
    """One rank of a z-slab sharded run of a 3-D problem of simulation.PROBLEMS."""

    def __init__(self, problem, p, elements, dt, rank, world, device, peers=None, steps=1):
        import torch

        self.torch = torch
        self.rank, self.world, self.p, self.dt = int(rank), int(world), int(p), float(dt)
        self.dev = torch.device("cuda", device)
        cls = PROBLEMS[problem] if isinstance(problem, str) else problem
        # the single-GPU problem object supplies dimensions, matrices and the sub-step program; its own
        # device context is never created
        self.model = cls(p, elements, timesteps_config(steps, dt), device=device)
        self.dims = self.model.dims
        if len(self.dims) != 3:
            raise ValueError("slab sharding is for the 3-D problems")
        self.n = tuple(d.dofs() for d in self.dims)
        nx, ny, nz = self.n
        self.pitch = nx + (nx & 1)           # x rows padded to 16 B (TMA), as the managed tensors are
        self.plane = ny * self.pitch         # doubles per z plane
        self.lines = self.pitch * ny         # z lines incl. the pad column (contiguous: the boundary kernels and
        #                                      pass B number them flat); state arrays are [S][K][lines]
        # ---- factors: let the model factorise exactly as on one GPU, but capture instead of uploading
        self.factors = {}

        class _Capture:
            def __init__(s, ndim):
                s.ndim = ndim

            def set_axis(s, ax, dim):
                pass

            def set_factor(s, ax, slot, lu, ipiv, kl, ku):
                self.factors[ax, slot] = (np.array(lu), np.array(ipiv), kl, ku)

            def upload(s, *a):
                pass

            def load_tensor(s, *a):
                s.want_forcing = a

            local_size = 0

        cap = _Capture(3)
        self.model.ctx = cap
        self.model.prepare_matrices()
        self.model.ctx = None
        self.substeps = self.model.substeps()
        self.zslots = sorted({int(s.slots[2]) for s in self.substeps})
        # ---- slabs = segments of the z lines; cuts must be free of row interchanges for every z factor
        lu0, ipiv0, kl, ku = self.factors[2, self.zslots[0]]
        self.bounds = segment_bounds(ipiv0, kl, world, 1) if world > 1 else np.array([0, nz], dtype=np.int32)
        self.z0, self.cz = int(self.bounds[rank]), int(self.bounds[rank + 1] - self.bounds[rank])
        if self.cz < max(p, 1):
            raise ValueError("slabs thinner than the spline degree are not supported")
        self.ctx = Context(self.n, lo=(0, 0, self.z0), cnt=(nx, ny, self.cz), device=device)
        self.ctx.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)
        for ax, d in enumerate(self.dims):
            self.ctx.set_axis(ax, d)
        for (ax, slot), (lu, ipiv, fkl, fku) in self.factors.items():
            self.ctx.set_factor(ax, slot, lu, ipiv, fkl, fku)
        self.seg = {}
        for slot in self.zslots:
            self.ctx.set_segments(2, slot, self.bounds, rank, 1)   # raises when the factor cannot be cut
            self.seg[slot] = self.ctx.segment_info(2, slot)
        KL = max(s["KL"] for s in self.seg.values())
        KD = max(s["KD"] for s in self.seg.values())
        self.KL, self.KD = KL, KD
        f64 = dict(dtype=torch.float64, device=self.dev)
        S = world
        nhalo = (max(np.diff(self.bounds)) + 2 * p) * self.plane
        sizes = [("h0", int(nhalo)), ("h1", int(nhalo)), ("dseg", S * KL * self.lines), ("x", S * KD * self.lines),
                 ("nbflags", 4)]   # neighbour barrier: 64-bit words [from prev, from next, own counter, -]
        if peers is None and world > 1:
            peers = _SymmPeers(self, sizes)
            self.sym = {name: peers.local(name) for name, _ in sizes}
        else:
            self.sym = {name: torch.zeros(count, **f64) for name, count in sizes}
            if peers is None:
                peers = _LocalPeers()
            peers.ranks.append(self)
        self.peers = peers
        self.work = torch.zeros(self.cz * self.plane, **f64)
        self.din = torch.zeros(S * KL * self.lines, **f64)
        self.tin = torch.zeros(S * KD * self.lines, **f64) if any(s["DB"] > 1 for s in self.seg.values()) else None
        self.forcing = None
        if any(float(s.form.gamma) != 0.0 for s in self.substeps):
            self.ctx.load_tensor(1, False, FORCING)          # slab of the load tensor (context lo / cnt)
            self.forcing = self.ctx.device_ptr(FORCING)
        # ---- fused distributed z sweep (one kernel per rank: pass A, exchange, pass B): every rank must agree
        self.lag = int(os.environ.get("ADSB_SLAB_LAG", "4"))
        rows = int(max(np.diff(self.bounds)))
        sc = -(-rows // 18)
        self.nl = 64
        while self.nl > 16 and (self.nl // 2) * sc > 288:
            self.nl //= 2
        nl_env = int(os.environ.get("ADSB_SLAB_NL", "0"))
        if nl_env in (16, 32, 64):
            self.nl = nl_env
        self.err_flag = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.fused = False
        self.halo_in_kernel = os.environ.get("ADSB_SLAB_HALO_IN_KERNEL", "1") != "0"   # fused sweep stores the halos itself
        self.neighbor_barrier = os.environ.get("ADSB_SLAB_NEIGHBOR_BARRIER", "1") != "0"
        self.want_fused = world > 1 and os.environ.get("ADSB_SLAB_FUSED", "1") != "0"
        self.cur = 0
        self.launches = 0
        self.exchange_bytes = 0
        self.timing, self._marks = False, []
        self.graph = None
        self.stream = None   # virtual ranks: each rank's kernels run on its own stream
        if self.want_fused and isinstance(self.peers, _SymmPeers):
            self.agree_on_fused()

    def fused_ok_locally(self):
        """can this rank run the fused kernel for every z factor (chain depth 1, shared memory, kernel variant)?"""
        if not self.want_fused:
            return False
        for slot in self.zslots:
            while self.nl >= 16 and not self.ctx.dist_sweep_check(2, slot, self.rank, self._view_lines(self.cz), self.nl, self.lag):
                self.nl //= 2
            if self.nl < 16:
                return False
        return True

    def agree_on_fused(self):
        """real run: all ranks take the fused path only if every rank can, with the smallest nl any rank needs"""
        import torch.distributed as dist

        ok = self.fused_ok_locally()
        t = self.torch.tensor([1 if ok else 0, self.nl if ok else 0], dtype=self.torch.int32, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t[0].item()) == 1:
            self.nl = int(t[1].item())
            self.enable_fused()
            self.peers.barrier(0)   # every inbox holds the sentinel before any neighbour stores into it

    def enable_fused(self):
        """the fused kernel's exchange protocol: every word of the state arrays holds a sentinel until the
        neighbour's store replaces it (see csrc/kernels_sweep_dist.cu)"""
        self.fused = True
        _lib.fill_sentinel(self.sym["dseg"])
        _lib.fill_sentinel(self.sym["x"])

    # ---- geometry helpers
    def _view(self, planes):
        return View.make([self.n[0], self.n[1], planes], [1, self.pitch, self.plane])

    def _view_lines(self, planes):
        """the same memory with the pad column counted in: every z plane is one contiguous run of lines"""
        return View.make([self.pitch, self.n[1], planes], [1, self.pitch, self.plane])

    def halo(self, k):
        return self.sym["h1" if k else "h0"]

    def interior(self, k):
        return self.halo(k)[self.p * self.plane:(self.p + self.cz) * self.plane]

    def set_local_state(self, host):
        """host: this rank's slab [cz][ny][nx] (dense); call on every rank, then `publish()`"""
        t = self.torch
        src = t.as_tensor(np.ascontiguousarray(host, dtype=np.float64).reshape(self.cz, self.n[1], self.n[0]))
        dst = self.interior(self.cur).view(self.cz, self.n[1], self.pitch)
        dst[:, :, :self.n[0]].copy_(src, non_blocking=True)

    def local_state(self):
        """(z0, array [cz][ny][nx])"""
        a = self.interior(self.cur).view(self.cz, self.n[1], self.pitch)[:, :, :self.n[0]]
        return self.z0, a.cpu().numpy()

    def _mark(self, name):
        if self.timing:
            e = self.torch.cuda.Event(enable_timing=True)
            e.record()
            self._marks.append((name, e))

    def phase_times(self):
        self.torch.cuda.synchronize()
        out = {}
        for (_, e0), (n1, e1) in zip(self._marks[:-1], self._marks[1:]):
            if n1 != "begin":
                out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        self._marks = []
        return out

    # ---- the phases of one sub-step; between two phases every rank must have finished the earlier one
    def phase_publish(self, k):
        """boundary planes of state buffer k -> the neighbours' halo regions (copy engines)"""
        p, pl, r = self.p, self.plane, self.rank
        if p == 0 or self.world == 1:
            return
        name = "h1" if k else "h0"
        mine = self.halo(k)
        if r > 0:   # my first p planes are the upper halo of rank r-1
            cn = int(self.bounds[r] - self.bounds[r - 1])
            dst = self.peers.tensor(r - 1, name)[(p + cn) * pl:(2 * p + cn) * pl]
            dst.copy_(mine[p * pl:2 * p * pl], non_blocking=True)
            self.exchange_bytes += 8 * p * pl
        if r < self.world - 1:   # my last p planes are the lower halo of rank r+1
            dst = self.peers.tensor(r + 1, name)[0:p * pl]
            dst.copy_(mine[self.cz * pl:(self.cz + p) * pl], non_blocking=True)
            self.exchange_bytes += 8 * p * pl

    def phase_fused(self, sub):
        """fused path: right-hand side straight into the interior of the other state buffer, x and y sweeps in
        place, then ONE kernel for the whole distributed z sweep (pass A, boundary exchange with the neighbours
        through peer pointers and flags, pass B).  Call phase_finish afterwards."""
        p, pl, nz = self.p, self.plane, self.n[2]
        H = self.halo(self.cur)
        out = self.interior(1 - self.cur).data_ptr()
        lo, hi = max(0, self.z0 - p), min(nz, self.z0 + self.cz + p)
        in_ptr = H.data_ptr() + 8 * (lo - self.z0 + p) * pl
        v = self._view(self.cz)
        self._mark("begin")
        self.ctx.rhs_view(sub.form, in_ptr, self._view(hi - lo), [0, 0, lo], out, v, [0, 0, self.z0],
                          forcing_ptr=self.forcing if float(sub.form.gamma) != 0.0 else None)
        self._mark("rhs")
        self.ctx.sweep_view(0, int(sub.slots[0]), out, v, out, v)
        self._mark("sweep_x")
        self.ctx.sweep_view(1, int(sub.slots[1]), out, v, out, v)
        self._mark("sweep_y")
        r, S = self.rank, self.world
        a = _lib.DistArgs()
        a.rank, a.nranks, a.nl, a.lag = r, S, self.nl, self.lag
        a.dseg_local, a.x_local = self.sym["dseg"].data_ptr(), self.sym["x"].data_ptr()
        a.dseg_next = self.peers.ptr(r + 1, "dseg") if r + 1 < S else None
        a.x_prev = self.peers.ptr(r - 1, "x") if r > 0 else None
        a.error_flag = self.err_flag.data_ptr()
        if self.halo_in_kernel and p > 0:
            # my first p planes are the upper halo of rank r-1, my last p planes the lower halo of rank r+1
            name = "h0" if (1 - self.cur) == 0 else "h1"
            if r > 0:
                cn = int(self.bounds[r] - self.bounds[r - 1])
                a.halo_prev = self.peers.ptr(r - 1, name) + 8 * (p + cn) * pl
            if r + 1 < S:
                a.halo_next = self.peers.ptr(r + 1, name)
            a.halo_planes = p
            self.exchange_bytes += 8 * p * pl * ((r > 0) + (r + 1 < S))
        slot = int(sub.slots[2])
        self.ctx.dist_sweep_view(2, slot, out, self._view_lines(self.cz), a)
        seg = self.seg[slot]
        self.exchange_bytes += 8 * self.lines * (seg["KL"] * (r + 1 < S) + seg["KD"] * (r > 0))
        self.launches += 4
        self._mark("sweep_z_fused")

    def phase_finish(self, sub=None):
        self.cur = 1 - self.cur
        if not self.halo_in_kernel:
            self.phase_publish(self.cur)
        self._mark("halo")

    def phase_local(self, sub):
        """right-hand side, x and y sweeps, pass A of the z sweep, forward boundary values"""
        p, pl, nz = self.p, self.plane, self.n[2]
        H = self.halo(self.cur)
        lo, hi = max(0, self.z0 - p), min(nz, self.z0 + self.cz + p)
        in_ptr = H.data_ptr() + 8 * (lo - self.z0 + p) * pl
        wk = self.work.data_ptr()
        v = self._view(self.cz)
        self._mark("begin")
        self.ctx.rhs_view(sub.form, in_ptr, self._view(hi - lo), [0, 0, lo], wk, v, [0, 0, self.z0],
                          forcing_ptr=self.forcing if float(sub.form.gamma) != 0.0 else None)
        self._mark("rhs")
        self.ctx.sweep_view(0, int(sub.slots[0]), wk, v, wk, v)
        self._mark("sweep_x")
        self.ctx.sweep_view(1, int(sub.slots[1]), wk, v, wk, v)
        self._mark("sweep_y")
        slot = int(sub.slots[2])
        if self.world == 1:
            self.ctx.sweep_view(2, slot, wk, v, wk, v)
            self.launches += 4
            self._mark("sweep_z")
            return
        self.ctx.seg_sweep_view(2, slot, self.rank, wk, v, wk, v)
        self._mark("sweep_z_a")
        DF = self.seg[slot]["DF"]
        dst = [self.peers.ptr(q, "dseg") for q in range(self.rank + 1, min(self.world, self.rank + DF + 1))]
        if dst:
            self.ctx.seg_dseg_view(2, slot, self.rank, self.rank + 1, self.z0, wk, self._view_lines(self.cz), dst)
            self.exchange_bytes += 8 * self.seg[slot]["KL"] * self.lines * len(dst)
            self.launches += 1
        self.launches += 4
        self._mark("dseg")

    def phase_backward(self, sub):
        """din from the previous ranks' forward values; backward boundary values to the previous ranks"""
        if self.world == 1:
            return
        slot = int(sub.slots[2])
        DB = self.seg[slot]["DB"]
        dst = [self.sym["x"].data_ptr()] + [self.peers.ptr(q, "x") for q in range(max(0, self.rank - DB), self.rank)]
        self.ctx.seg_din_view(2, slot, self.rank, self.rank + 1, self.z0, self.work.data_ptr(), self._view_lines(self.cz),
                              self.sym["dseg"].data_ptr(), self.din.data_ptr(), dst)
        self.exchange_bytes += 8 * self.seg[slot]["KD"] * self.lines * (len(dst) - 1)
        self.launches += 1
        self._mark("din")

    def phase_correct(self, sub):
        """pass B into the interior of the other state buffer; swap; publish the new boundary planes"""
        nxt = 1 - self.cur
        out_ptr = self.interior(nxt).data_ptr()
        v = self._view_lines(self.cz)
        if self.world == 1:
            self.interior(nxt).copy_(self.work, non_blocking=True)
        else:
            slot = int(sub.slots[2])
            tin = self.sym["x"]
            if self.seg[slot]["DB"] > 1:
                self.ctx.seg_tin(2, slot, self.rank, self.rank + 1, self.lines, self.sym["x"].data_ptr(), self.tin.data_ptr())
                tin = self.tin
                self.launches += 1
            self.ctx.seg_correct_view(2, slot, self.rank, self.rank + 1, self.z0, self.work.data_ptr(), v, out_ptr, v,
                                      self.din.data_ptr(), tin.data_ptr())
            self.launches += 1
        self._mark("correct")
        self.cur = nxt
        self.phase_publish(self.cur)
        self._mark("halo")

    PHASES = ("phase_local", "phase_backward", "phase_correct")

    # ---- one rank of a real multi-GPU run
    def publish(self):
        self.peers.barrier(0)
        self.phase_publish(self.cur)
        self.peers.barrier(1)

    def step(self):
        if self.fused:
            for sub in self.substeps:
                self.phase_fused(sub)
                self.phase_finish()
                # the neighbours' boundary planes have landed in my halo regions; consecutive barriers alternate
                # between two signal-pad channels
                self.step_barrier()
                self._mark("barrier")
            return
        for sub in self.substeps:
            self.phase_local(sub)
            self.peers.barrier(0)
            self._mark("barrier")
            self.phase_backward(sub)
            self.peers.barrier(1)
            self._mark("barrier")
            self.phase_correct(sub)
            self.peers.barrier(2)
            self._mark("barrier")

    def step_barrier(self):
        """end of a fused sub-step: the neighbours' halo stores have landed and they are done with my boundary values.
        Everything is between neighbours, so a two-neighbour flag barrier replaces the all-rank signal barrier."""
        if self.neighbor_barrier and isinstance(self.peers, _SymmPeers):
            r, S = self.rank, self.world
            self.ctx.neighbor_barrier(self.sym["nbflags"].data_ptr(), self.peers.ptr(r - 1, "nbflags") if r > 0 else None,
                                      self.peers.ptr(r + 1, "nbflags") if r + 1 < S else None, self.err_flag.data_ptr())
        else:
            self._bar = 1 - getattr(self, "_bar", 1)
            self.peers.barrier(self._bar)

    def advance(self, nsteps, graph=None):
        """nsteps steps; with graph=True (default on > 1 GPU, ADSB_SLAB_GRAPH=0 disables) one step is captured
        into a CUDA graph after two eager steps and replayed (the state buffers alternate, so a graph holds
        two steps)."""
        use_graph = (self.world > 1 and os.environ.get("ADSB_SLAB_GRAPH", "1") != "0") if graph is None else graph
        n = int(nsteps)
        while n > 0:
            if use_graph and n >= 2 and self.cur == 0 and not self.timing and getattr(self, "_eager", 0) >= 2:
                if self.graph is None:
                    self._capture()
                self.graph.replay()
                self.launches += self._graph_delta[0]
                self.exchange_bytes += self._graph_delta[1]
                n -= 2
            else:
                self.step()
                self._eager = getattr(self, "_eager", 0) + 1
                n -= 1

    def _capture(self):
        import gc

        torch = self.torch
        main = torch.cuda.current_stream(self.dev)
        l0, b0 = self.launches, self.exchange_bytes
        # no allocation may be freed while the stream is capturing: collect garbage now (an earlier simulation's
        # symmetric memory, say) and keep the collector off until the capture has ended
        gc.collect()
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            with torch.cuda.graph(g):
                self.ctx.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)
                self.step()
                self.step()
        finally:
            if gc_was_on:
                gc.enable()
        self.ctx.set_stream(main.cuda_stream)
        self._graph_delta = (self.launches - l0, self.exchange_bytes - b0)
        self.launches, self.exchange_bytes = l0, b0
        self.graph = g


class VirtualCluster:
    """`world` ranks of SlabSim on ONE device, run in lockstep (phase by phase): the single-GPU test vehicle
    of the distributed path.  Peer stores land in ordinary tensors of the same process."""

    def __init__(self, problem, p, elements, dt, world, device=0):
        import torch

        self.peers = _LocalPeers()
        self.ranks = [SlabSim(problem, p, elements, dt, r, world, device, peers=self.peers) for r in range(world)]
        self.n = self.ranks[0].n
        # fused z sweep: the ranks' kernels wait for one another, so they must run CONCURRENTLY: one stream per
        # rank and an SM cap that lets all of them be resident at once (the real run has a GPU per rank)
        self.fused = world > 1 and all(s.fused_ok_locally() for s in self.ranks)
        if self.fused:
            nl = min(s.nl for s in self.ranks)
            sms = torch.cuda.get_device_properties(device).multi_processor_count
            for s in self.ranks:
                s.nl = nl
                s.enable_fused()
                s.stream = torch.cuda.Stream(device=device)
                s.ctx.set_stream(s.stream.cuda_stream)
                s.ctx.set_sm_limit(max(1, sms // world))

    def set_state(self, full):
        nx, ny, nz = self.n
        a = np.asarray(full).reshape(nz, ny, nx)
        for s in self.ranks:
            s.set_local_state(a[s.z0:s.z0 + s.cz])
        for s in self.ranks:
            s.phase_publish(s.cur)

    def state(self):
        nx, ny, nz = self.n
        out = np.zeros((nz, ny, nx))
        for s in self.ranks:
            z0, a = s.local_state()
            out[z0:z0 + a.shape[0]] = a
        return out.ravel()

    def step(self, nsteps=1):
        torch = self.ranks[0].torch
        for _ in range(nsteps):
            for i in range(len(self.ranks[0].substeps)):
                if self.fused:
                    torch.cuda.synchronize()
                    for s in self.ranks:
                        with torch.cuda.stream(s.stream):
                            s.phase_fused(s.substeps[i])
                    torch.cuda.synchronize()
                    for s in self.ranks:
                        with torch.cuda.stream(s.stream):
                            s.phase_finish()
                    continue
                for phase in SlabSim.PHASES:
                    for s in self.ranks:
                        getattr(s, phase)(s.substeps[i])
        torch.cuda.synchronize()
        if self.fused and any(int(s.err_flag.item()) for s in self.ranks):
            raise RuntimeError("fused distributed sweep: a flag wait timed out")


def gather_state(sim):
    """All ranks: the full tensor [z][y][x] on every rank (test helper; host memory)."""
    import torch.distributed as dist

    z0, arr = sim.local_state()
    pieces = [None] * sim.world
    dist.all_gather_object(pieces, (z0, arr))
    nx, ny, nz = sim.n
    full = np.zeros((nz, ny, nx))
    for z, a in pieces:
        full[z:z + a.shape[0]] = a
    return full
