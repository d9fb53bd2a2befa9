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
* Exchange, fused into the sweep ("p2p", the default on > 1 GPU): the receive blocks and the halo'ed
  state buffers live in symmetric memory (torch.distributed._symmetric_memory: every rank maps every
  peer's allocation).  The B sweep's row-offset table points each destination's rows straight at that
  peer's receive block, so the sweep kernel's own TMA stores carry the all-to-all over NVLink while it
  computes -- no send buffer, no collective kernel, no local copy.  Boundary planes go to the
  neighbours' halo regions as peer copies.  Two signal-pad barriers per step order the writes.
  "nccl" (ADSB_SHARDED_EXCHANGE=nccl, or when symmetric memory cannot be set up) keeps the send
  buffer + all_to_all_single + batched isend/irecv path.
* Exchange through the copy engines ("ce", the default on > 1 GPU; same symmetric memory): measured
  on this pool's B200s, stores issued by SMs over NVLink reach ~410 GB/s per direction with both
  directions busy (the fused sweep and NCCL's all-to-all alike), the copy engines ~730 GB/s.  So the
  slab is cut into chunks of A-planes; per chunk the right-hand side, the x sweep and the B sweep (into a
  local send buffer in block layout; the rows this rank keeps go straight into its own receive block)
  run on the main stream, and the chunk's blocks are pushed to the peers' receive blocks with strided
  copy-engine copies (adsb_copy2d) on a second stream -- no SM is spent on the exchange and it hides
  behind the next chunk's compute.  ADSB_SHARDED_EXCHANGE selects ce | p2p | nccl.
* "p2p" can also run chunked (ADSB_SHARDED_CHUNKS): the B sweep of chunk c on a second stream on a few
  SMs (adsb_set_sm_limit) next to the compute of chunk c+1; kept for measurements.

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

    # ---- exchange layout of the copy-engine path: block (src r -> dst s) = [a_local of r][b_local of s][x]
    # (padded to cmax), i.e. the A-planes of the sender are the slowest index, so ANY range of planes -- a
    # chunk -- is one contiguous piece per peer (strided peer copies are slow: ~1 us per row)
    def chunks_of(self, A, fractions, rank=None):
        """[(first local plane, planes)]: the slab cut at the given cumulative fractions of cmax[A]"""
        c = self.sizes[A][self.rank if rank is None else rank]
        cuts = [0] + [min(c, int(round(f * self.cmax[A]))) for f in fractions] + [c]
        cuts = sorted(set(cuts))
        return [(lo, hi - lo) for lo, hi in zip(cuts[:-1], cuts[1:]) if hi > lo]

    def pack_offsets_t(self, A):
        """off_out[j] of the B sweep (relative to the send buffer); the line (x, a_local) adds
        x + a_local*cmax[B]*nx."""
        B, nx = 3 - A, self.n[0]
        off = np.zeros(self.n[B], dtype=np.int64)
        for j in range(self.n[B]):
            s = self.owner(B, j)
            off[j] = s * self.block + (j - self.starts[B][s]) * nx
        return off

    def unpack_offsets_t(self, A):
        """off_in[k] of the A sweep reading the receive blocks; the line (x, b_local) adds x + b_local*nx."""
        B, nx = 3 - A, self.n[0]
        off = np.zeros(self.n[A], dtype=np.int64)
        for k in range(self.n[A]):
            r = self.owner(A, k)
            off[k] = r * self.block + (k - self.starts[A][r]) * self.cmax[B] * nx
        return off

    def emulate_pack_t(self, A, work):
        """work [cnt(A)][n[B]][nx] -> flat send buffer, blocks [a][b][x]"""
        B, nx = 3 - A, self.n[0]
        send = np.zeros(self.world * self.block)
        off = self.pack_offsets_t(A)
        for a in range(self.cnt(A)):
            for j in range(self.n[B]):
                o = off[j] + a * self.cmax[B] * nx
                send[o:o + nx] = work[a, j]
        return send

    def emulate_unpack_t(self, A, recv):
        """flat receive buffer (blocks [a][b][x]) -> [cnt(B)][n[A]][nx]"""
        B, nx = 3 - A, self.n[0]
        out = np.zeros((self.cnt(B), self.n[A], nx))
        off = self.unpack_offsets_t(A)
        for b in range(self.cnt(B)):
            for k in range(self.n[A]):
                o = off[k] + b * nx
                out[b, k] = recv[o:o + nx]
        return out

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
        self.work = torch.zeros(self.plan.work_size(), **f64)
        self.exchange = "nccl"
        mode = os.environ.get("ADSB_SHARDED_EXCHANGE", "ce")
        # chunks of >= 32 planes (the right-hand side re-reads 2p planes per chunk)
        self.nchunk = 1
        if world > 1 and mode in ("p2p", "ce"):
            # measured on B200 (512^3): 2 chunks beat 1, 3 and 4 at 2, 4 and 8 GPUs -- every extra chunk costs
            # ~0.08 ms of kernel ramp-up / halo planes, and only the last chunk's exchange is exposed
            want = int(os.environ.get("ADSB_SHARDED_CHUNKS", "2" if mode == "ce" else "1"))
            min_planes = max(int(os.environ.get("ADSB_SHARDED_MIN_PLANES", "32")), max(p, 1))
            self.nchunk = max(1, min(want, min(self.plan.sizes[1] + self.plan.sizes[2]) // min_planes))
        self.blk = self.plan.block   # doubles per (src, dst) block
        # copy-engine path: cumulative plane fractions at which the slab is cut.  Only the LAST chunk's
        # exchange is exposed, so it is the small one; the first must still finish its exchange behind the
        # last one's compute (measured: exchange ~0.4-0.7 x the compute of the same planes)
        first = float(os.environ.get("ADSB_SHARDED_SPLIT", "0.62"))
        self.fractions = [first * (i + 1) / (self.nchunk - 1) for i in range(self.nchunk - 1)] if self.nchunk > 1 else []
        if world > 1 and mode in ("p2p", "ce"):
            self._setup_p2p()
            if self.exchange == "p2p" and mode == "ce":
                self.exchange = "ce"
        if self.exchange == "nccl":
            self.nchunk, self.blk = 1, self.plan.block
            self.halo = [torch.zeros(self.plan.halo_size(), **f64) for _ in range(2)]
            self.recv = torch.zeros(world * self.plan.block, **f64)
        if self.exchange in ("nccl", "ce"):
            self.send = torch.zeros(world * self.blk, **f64)
        if self.exchange == "ce":
            # B sweep: other ranks' rows into the send blocks, my own rows straight into my receive block k
            self._ce_off, self._ce_unpack = {}, {}
            for A in (1, 2):
                B = 3 - A
                self._ce_unpack[A] = self.plan.unpack_offsets_t(A)
                for k in (0, 1):
                    off = self.plan.pack_offsets_t(A)
                    j0, jn = self.plan.starts[B][rank], self.plan.sizes[B][rank]
                    off[j0:j0 + jn] += (self.recv2[k].data_ptr() - self.send.data_ptr()) // 8
                    self._ce_off[A, k] = off
        self.step_index = 0
        self.timing, self._marks = False, []
        self.graph, self._eager_steps, self._graph_delta = None, 0, (0, 0)
        self.use_graph = (world > 1 and os.environ.get("ADSB_SHARDED_GRAPH", "1") != "0"
                          and self.exchange in ("p2p", "ce"))
        if self.exchange == "ce" or self.nchunk > 1:
            self.bg_sms = int(os.environ.get("ADSB_SHARDED_BG_SMS", "32"))
            self.sm_count = torch.cuda.get_device_properties(self.dev).multi_processor_count
            self.bg_stream = torch.cuda.Stream(device=self.dev)
            # peer copies to different destinations go round-robin over a few streams (several copy engines)
            ncs = max(1, int(os.environ.get("ADSB_SHARDED_COPY_STREAMS", "3")))
            self.copy_streams = [self.bg_stream] + [torch.cuda.Stream(device=self.dev) for _ in range(ncs - 1)]
            self.ctx_bg = Context(self.n, device=device)   # same tables and factors, its own stream
            self.ctx_bg.set_stream(self.bg_stream.cuda_stream)
            for ax in range(3):
                self.ctx_bg.set_axis(ax, self.dim)
                self.ctx_bg.set_factor(ax, 0, lu, ipiv, p, p)
            self._chunk_events = [torch.cuda.Event() for _ in range(self.nchunk)]
        self.cur = 0            # which halo buffer holds the state
        self.orientation = 2    # z-slabs
        self.form = Form.make(1.0, (dt, dt, dt))
        self._off = {A: (self.plan.pack_offsets(A), self.plan.unpack_offsets(A)) for A in (1, 2)}
        self.launches = 0
        self.exchange_bytes = 0

    # ---- symmetric-memory exchange
    def _setup_p2p(self):
        """One symmetric allocation per rank: [recv 0 | recv 1 | halo 0 | halo 1].  Every rank agrees
        through an all-reduce whether the set-up worked, so all take the same path."""
        torch, dist = self.torch, self.dist
        nrecv, nhalo = self.world * self.blk, self.plan.halo_size()
        nrecv += nrecv & 1
        nhalo += nhalo & 1
        ok, err = 1, None
        try:
            import torch.distributed._symmetric_memory as symm

            buf = symm.empty(2 * nrecv + 2 * nhalo, dtype=torch.float64, device=self.dev)
            buf.zero_()
            hdl = symm.rendezvous(buf, dist.group.WORLD)
            peers = [int(v) for v in hdl.buffer_ptrs]
            if len(peers) != self.world or any((v - peers[self.rank]) % 16 for v in peers):
                ok = 0
        except Exception as e:  # noqa: BLE001 -- any failure means: fall back to NCCL, on every rank
            ok, err = 0, e
        flag = torch.tensor([ok], dtype=torch.int32, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if self.rank == 0:
                print(f"[adsb] symmetric memory unavailable ({err}); using the NCCL exchange", flush=True)
            return
        self.exchange, self.sym, self.hdl, self.peers = "p2p", buf, hdl, peers
        self.recv2_off = [0, nrecv]                       # element offsets inside the symmetric allocation
        self.halo_off = [2 * nrecv, 2 * nrecv + nhalo]
        self.recv2 = [buf[o:o + nrecv] for o in self.recv2_off]
        self.halo = [buf[o:o + nhalo] for o in self.halo_off]
        self._p2p_off = {}
        for A in (1, 2):
            # destination row j of the B sweep lives in the receive block `rank` of the peer owning j
            B, nx = 3 - A, self.n[0]
            off = np.zeros(self.n[B], dtype=np.int64)
            for j in range(self.n[B]):
                dst = self.plan.owner(B, j)
                off[j] = ((peers[dst] - peers[self.rank]) // 8 + self.rank * self.plan.block
                          + (j - self.plan.starts[B][dst]) * self.plan.cmax[A] * nx)
            self._p2p_off[A] = off
        hdl.barrier(channel=0)

    def _publish_halo(self, k, A):
        """Copy the boundary planes of halo buffer k (slabs along A) into the neighbours' halo regions."""
        p, pl, c = self.p, self.plan.plane(A), self.plan.cnt(A)
        mine = self.halo[k]
        f64 = self.torch.float64
        if self.rank > 0:  # my first p planes are the upper halo of rank - 1
            cn = self.plan.sizes[A][self.rank - 1]
            dst = self.hdl.get_buffer(self.rank - 1, (p * pl,), f64, self.halo_off[k] + (p + cn) * pl)
            dst.copy_(mine[p * pl:2 * p * pl], non_blocking=True)
            self.exchange_bytes += 8 * p * pl
        if self.rank < self.world - 1:  # my last p planes are the lower halo of rank + 1
            dst = self.hdl.get_buffer(self.rank + 1, (p * pl,), f64, self.halo_off[k])
            dst.copy_(mine[c * pl:(c + p) * pl], non_blocking=True)
            self.exchange_bytes += 8 * p * pl

    # ---- state access (host <-> device), canonical local layout of the current orientation
    def interior(self, buf, A):
        pl = self.plan.plane(A)
        return buf[self.p * pl:(self.p + self.plan.cnt(A)) * pl]

    def set_local_state(self, host):
        """host: this rank's z-slab [cnt_z][ny][nx] (flat)"""
        self.orientation, self.cur, self.step_index = 2, 0, 0   # refresh_halo below synchronises the ranks
        self.interior(self.halo[0], 2).copy_(self.torch.as_tensor(host).reshape(-1), non_blocking=True)
        self.refresh_halo()

    def overwrite_local_state(self, host):
        """Replace this rank's slab in the CURRENT orientation (same byte count either way)."""
        A = self.orientation
        dst = self.interior(self.halo[self.cur], A)
        dst.copy_(self.torch.as_tensor(host).reshape(-1)[:dst.numel()], non_blocking=True)
        self.refresh_halo()

    def refresh_halo(self):
        """After the interior was written from outside a step: re-publish the boundary planes (the NCCL
        path exchanges halos at the start of every step anyway)."""
        if self.exchange in ("p2p", "ce"):
            self.hdl.barrier(channel=0)   # nobody still reads the halo regions written next
            self._publish_halo(self.cur, self.orientation)
            self.hdl.barrier(channel=1)

    def local_state(self):
        """(orientation, lo, cnt, array [cnt][n_middle][nx])"""
        A = self.orientation
        arr = self.interior(self.halo[self.cur], A).cpu().numpy()
        return A, self.plan.lo(A), self.plan.cnt(A), arr.reshape(self.plan.cnt(A), self.n[3 - A], self.n[0])

    # ---- per-phase device times (bench only)
    def _mark(self, name):
        if self.timing:
            e = self.torch.cuda.Event(enable_timing=True)
            e.record()
            self._marks.append((name, e))

    def phase_times(self):
        """ms per phase summed over the steps since timing was switched on; clears the marks."""
        self.torch.cuda.synchronize()
        out = {}
        for (n0, e0), (n1, e1) in zip(self._marks[:-1], self._marks[1:]):
            if n1 != "begin":
                out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        self._marks = []
        return out

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
        p2p = self.exchange == "p2p"
        self._mark("begin")
        if self.exchange == "ce" or self.nchunk > 1:
            return self._step_pipelined()
        if not p2p:
            self._halo_exchange(H, A)
            self._mark("halo")
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
        self._mark("rhs")
        # x sweep in place, B sweep into the send blocks
        self.ctx.sweep_view(0, 0, self.work.data_ptr(), _view(*v["work"]), self.work.data_ptr(), _view(*v["work"]))
        self._mark("sweep_x")
        if p2p:
            # B sweep: rows of destination s go straight into receive block `rank` of peer s (TMA / plain
            # stores over NVLink); receive buffers alternate so a peer still reading last step's is safe
            src = self.recv2[self.step_index & 1]
            self.ctx.sweep_view(B, 0, self.work.data_ptr(), _view(*v["work"]), src.data_ptr(), _view(*v["send"]),
                                off_out=self._p2p_off[A])
            self.exchange_bytes += 8 * plan.cnt(A) * (self.n[B] - plan.cnt(B)) * nx
            self._mark("sweep_b+exchange")
            self.hdl.barrier(channel=0)   # every peer's rows have landed in my receive blocks
            self._mark("barrier")
        else:
            self.ctx.sweep_view(B, 0, self.work.data_ptr(), _view(*v["work"]), self.send.data_ptr(), _view(*v["send"]),
                                off_out=self._off[A][0])
            self._mark("sweep_b")
            if self.world > 1:
                self.dist.all_to_all_single(self.recv, self.send)
                self._mark("all_to_all")
                self.exchange_bytes += 8 * plan.block * (self.world - 1)
                src = self.recv
            else:
                src = self.send
        # A sweep out of the receive blocks into the interior of the next state (orientation B)
        out_ptr = Hn.data_ptr() + 8 * p * plan.plane(B)
        self.ctx.sweep_view(A, 0, src.data_ptr(), _view(*v["recv"]), out_ptr, _view(*v["new"]),
                            off_in=self._off[A][1])
        self._mark("sweep_a")
        self.launches += 4
        self.cur = 1 - self.cur
        self.orientation = B
        self.step_index += 1
        if p2p:
            self._publish_halo(self.cur, B)
            self.hdl.barrier(channel=1)   # my halo regions hold the neighbours' new boundary planes
            self._mark("halo")

    def _step_pipelined(self):
        """p2p step with the exchange hidden behind the next chunk's compute (see the module docstring)."""
        torch, plan, p, A = self.torch, self.plan, self.p, self.orientation
        B, nx = 3 - A, self.n[0]
        H, Hn = self.halo[self.cur], self.halo[1 - self.cur]
        a0, c, pl = plan.lo(A), plan.cnt(A), plan.plane(A)
        v = plan.views(A)
        k = self.step_index & 1
        recv = self.recv2[k]
        ce = self.exchange == "ce"
        blk = self.blk
        main = torch.cuda.current_stream(self.dev)
        if ce:
            chunks = plan.chunks_of(A, self.fractions)
            arow = plan.cmax[B] * nx          # doubles per A-plane of a block
            s_send = [1, 0, 0]
            s_send[A] = arow
        else:
            chunks = [(cs, cn) for cs, cn in zip(*split(c, self.nchunk)) if cn > 0]
        for ci, (cs, cn) in enumerate(chunks):
            last = ci == len(chunks) - 1
            # p2p: while a background sweep is in flight the foreground kernels plan for the remaining SMs
            if not ce:
                self.ctx.set_sm_limit(0 if ci == 0 else self.sm_count - self.bg_sms)
            g_lo, g_hi = max(0, a0 + cs - p), min(self.n[A], a0 + cs + cn + p)
            nin, nout = list(v["work"][0]), list(v["work"][0])
            nin[A], nout[A] = g_hi - g_lo, cn
            in_lo, out_lo = [0, 0, 0], [0, 0, 0]
            in_lo[A], out_lo[A] = g_lo, a0 + cs
            in_ptr = H.data_ptr() + 8 * (g_lo - a0 + p) * pl
            wk_ptr = self.work.data_ptr() + 8 * cs * pl
            self.ctx.rhs_view(self.form, in_ptr, _view(nin, v["work"][1]), in_lo, wk_ptr, _view(nout, v["work"][1]), out_lo)
            self._mark("rhs")
            self.ctx.sweep_view(0, 0, wk_ptr, _view(nout, v["work"][1]), wk_ptr, _view(nout, v["work"][1]))
            self._mark("sweep_x")
            if ce:
                # B sweep into the send blocks (own rows: my receive block); then the copy engines push this
                # chunk's piece of every block -- contiguous: cn planes of cmax[B]*nx doubles -- to the peers
                self.ctx.sweep_view(B, 0, wk_ptr, _view(nout, v["work"][1]), self.send.data_ptr() + 8 * cs * arow,
                                    _view(nout, s_send), off_out=self._ce_off[A, k])
                self._mark("sweep_b")
                self._chunk_events[ci].record(main)
                for cs_ in self.copy_streams:
                    cs_.wait_event(self._chunk_events[ci])
                for step_to in range(1, self.world):
                    dst = (self.rank + step_to) % self.world
                    nbytes = 8 * cn * arow
                    self.ctx_bg.set_stream(self.copy_streams[step_to % len(self.copy_streams)].cuda_stream)
                    self.ctx_bg.copy2d(self.peers[dst] + 8 * (self.recv2_off[k] + self.rank * blk + cs * arow), nbytes,
                                       self.send.data_ptr() + 8 * (dst * blk + cs * arow), nbytes, nbytes, 1)
                continue
            self._chunk_events[ci].record(main)
            self.bg_stream.wait_event(self._chunk_events[ci])
            # B sweep of the chunk on the background stream, rows straight into the peers' receive blocks
            self.ctx_bg.set_sm_limit(0 if last else self.bg_sms)
            self.ctx_bg.sweep_view(B, 0, wk_ptr, _view(nout, v["work"][1]), recv.data_ptr() + 8 * cs * nx,
                                   _view(nout, v["send"][1]), off_out=self._p2p_off[A])
        self.ctx.set_sm_limit(0)
        if ce:
            for cs_ in self.copy_streams:
                main.wait_stream(cs_)
            self.ctx_bg.set_stream(self.bg_stream.cuda_stream)
        else:
            main.wait_stream(self.bg_stream)
        self.exchange_bytes += 8 * c * (self.n[B] - plan.cnt(B)) * nx
        self._mark("exchange_tail")
        self.hdl.barrier(channel=0)
        self._mark("barrier")
        out_ptr = Hn.data_ptr() + 8 * p * plan.plane(B)
        if ce:
            s_recv = list(v["recv"][1])
            s_recv[B] = nx
            self.ctx.sweep_view(A, 0, recv.data_ptr(), _view(v["recv"][0], s_recv), out_ptr, _view(*v["new"]),
                                off_in=self._ce_unpack[A])
        else:
            self.ctx.sweep_view(A, 0, recv.data_ptr(), _view(*v["recv"]), out_ptr, _view(*v["new"]),
                                off_in=self._off[A][1])
        self._mark("sweep_a")
        self.launches += 1 + 3 * len(chunks)
        self.cur = 1 - self.cur
        self.orientation = B
        self.step_index += 1
        self._publish_halo(self.cur, B)
        self.hdl.barrier(channel=1)
        self._mark("halo")

    def advance(self, nsteps):
        """nsteps steps.  Two consecutive steps are one full cycle of the slab orientation and of the
        receive-buffer parity, so (p2p / ce exchange, ADSB_SHARDED_GRAPH != 0) that pair is captured
        once into a CUDA graph -- both streams, the copy-engine pushes and the signal barriers -- and
        replayed: a step is ~15-30 launches of 30-300 us kernels, which the host cannot issue fast
        enough one call at a time."""
        n = int(nsteps)
        while n > 0:
            if (self.use_graph and n >= 2 and self.orientation == 2 and self.step_index % 2 == 0
                    and self._eager_steps >= 2 and not self.timing):
                if self.graph is None:
                    self._capture()
                self.graph.replay()
                self.launches += self._graph_delta[0]
                self.exchange_bytes += self._graph_delta[1]
                n -= 2
            else:
                self.step()
                self._eager_steps += 1
                n -= 1

    def _capture(self):
        torch = self.torch
        main = torch.cuda.current_stream(self.dev)
        l0, b0, s0 = self.launches, self.exchange_bytes, self.step_index
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.ctx.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)
            self.step()
            self.step()
        self.ctx.set_stream(main.cuda_stream)
        self._graph_delta = (self.launches - l0, self.exchange_bytes - b0)
        self.launches, self.exchange_bytes, self.step_index = l0, b0, s0
        self.graph = g


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
    # per-phase device times of rank 0 (separate pass, events between the launches)
    sim.timing = True
    sim.advance(args.steps)
    phases = {k: v / args.steps for k, v in sim.phase_times().items()}
    sim.timing = False

    # end to end: upload the slab from pinned host memory, one step, download the slab, every step
    k2 = 2
    out_host = torch.empty(sim.plan.halo_size(), dtype=torch.float64).pin_memory()
    sim.set_local_state(host)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(k2):
        sim.overwrite_local_state(host)  # z- or y-slab orientation: same byte count
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
                       "parallelism": f"{world} slabs, one all-to-all + p-plane halo per step, orientation alternates",
                       "cuda_graph": bool(sim.graph is not None),
                       "exchange": {"ce": f"{sim.nchunk} chunks; blocks pushed into the peers' symmetric memory by the copy "
                                          "engines behind the next chunk's compute, 2 signal barriers per step",
                                    "p2p": "peer stores over NVLink fused into the sweep (symmetric memory), 2 signal "
                                           "barriers per step",
                                    "nccl": "NCCL all_to_all_single + batched isend/irecv"}[sim.exchange]},
            "roofline": {"bound": "hbm", "kernel": "whole step, per GPU", "achieved": BYTES_PER_DOF_STEP * N / world / step_s / 1e9,
                         "peak": hbm, "peak_kind": peak_kind, "unit": "GB/s",
                         "frac": BYTES_PER_DOF_STEP * N / world / step_s / 1e9 / hbm, "traffic": None,
                         "exchange_bytes_per_gpu_step": timed_exchange / max(args.steps, 1),
                         "phase_ms_rank0": phases},
            "clocks": clocks, "gpu_launches": timed_launches * world, "finite": finite,
            "e2e": {"value": N * k2 / el, "unit": UNIT, "h2d_bytes_per_step": 8 * N, "d2h_bytes_per_step": 8 * N,
                    "steps": k2, "ms_per_step": 1e3 * el / k2,
                    "note": "each rank: pinned slab upload + step + slab download per step"},
        }))
    dist.barrier()
    dist.destroy_process_group()
