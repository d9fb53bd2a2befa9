"""Host-side logic of the slab decomposition (iga_ads_b200/sharded.py) on CPU: the pack / unpack
offset tables that the sweeps use around the all-to-all, exercised with real collectives over gloo
(world_size 2 and 3) and identity "sweeps" in numpy."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from iga_ads_b200.sharded import SlabPlan, split


def test_split_is_balanced_and_contiguous():
    for n, parts in ((514, 8), (514, 2), (11, 3), (8, 8)):
        st, sz = split(n, parts)
        assert sum(sz) == n and max(sz) - min(sz) <= 1
        assert st == [sum(sz[:r]) for r in range(parts)]


def test_plan_rejects_slabs_thinner_than_p():
    with pytest.raises(ValueError):
        SlabPlan((10, 10, 10), 3, 8, 0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _all_to_all(send, plan, rank, world):
    """all_to_all_single with equal blocks; gloo has no alltoall, so gather everything and keep the
    blocks addressed to this rank (the product path uses NCCL's all_to_all_single)."""
    everything = [torch.zeros_like(send) for _ in range(world)]
    dist.all_gather(everything, send)
    return torch.cat([everything[r][rank * plan.block:(rank + 1) * plan.block] for r in range(world)])


def _worker(rank, world, port, n, p, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nx, ny, nz = n
        full = np.random.default_rng(5).standard_normal((nz, ny, nx))   # [z][y][x]
        plan = SlabPlan(n, p, world, rank)
        # orientation z-slabs -> y-slabs
        z0, cz = plan.lo(2), plan.cnt(2)
        send = torch.from_numpy(plan.emulate_pack(2, full[z0:z0 + cz]))
        recv = _all_to_all(send, plan, rank, world)
        got = plan.emulate_unpack(2, recv.numpy())                      # [y_local][z][x]
        y0, cy = plan.lo(1), plan.cnt(1)
        want = np.transpose(full[:, y0:y0 + cy, :], (1, 0, 2))
        ok1 = np.array_equal(got, want)
        # and back: y-slabs -> z-slabs
        send = torch.from_numpy(plan.emulate_pack(1, got))
        back = plan.emulate_unpack(1, _all_to_all(send, plan, rank, world).numpy())  # [z_local][y][x]
        ok2 = np.array_equal(back, full[z0:z0 + cz])
        # the block layout of the copy-engine exchange ([a][b][x]: any range of planes is one piece per peer)
        send = torch.from_numpy(plan.emulate_pack_t(2, full[z0:z0 + cz]))
        got = plan.emulate_unpack_t(2, _all_to_all(send, plan, rank, world).numpy())
        ok1 = ok1 and np.array_equal(got, want)
        for fr in ([0.62], [0.3, 0.6]):
            ch = plan.chunks_of(2, fr)
            ok1 = ok1 and sum(c for _, c in ch) == cz and ch[0][0] == 0 and all(c > 0 for _, c in ch)
        q.put((rank, ok1, ok2))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,p", [(2, (7, 9, 11), 2), (3, (6, 10, 8), 1)])
def test_slab_exchange_roundtrip_gloo(world, n, p):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, p, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(r[0] for r in res) == list(range(world))
    assert all(r[1] and r[2] for r in res), res


def test_offsets_cover_blocks_without_overlap():
    n, p, world = (5, 9, 7), 1, 3
    for rank in range(world):
        plan = SlabPlan(n, p, world, rank)
        for A in (1, 2):
            B = 3 - A
            off = plan.pack_offsets(A)
            touched = set()
            for a in range(plan.cnt(A)):
                for j in range(n[B]):
                    for x in range(n[0]):
                        o = int(off[j]) + a * n[0] + x
                        assert o not in touched and 0 <= o < world * plan.block
                        touched.add(o)
            assert len(touched) == plan.cnt(A) * n[B] * n[0]
