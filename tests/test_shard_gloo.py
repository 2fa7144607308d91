"""Row-band sharding logic on CPU: world_size 2 (and 3) over gloo.  Every rank
fills only the rows it owns, refreshes its halo rows through the same
exchange() the GPU path uses, and the band is then checked two ways: the held
rows equal the corresponding rows of the full image, and the oracle applied
to the held band reproduces the full-image result on the owned rows (sharding
invariance) -- which is exactly the contract of
morsi_cuda_apply_band_device with a halo of stages x reach rows."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import imscript_b200 as M
from imscript_b200.shard import BandPlan, exchange
from oracle import oracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, w, h, element, op, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o = oracle()
        e = o.element(element)
        up, down = M.halo_rows(op, e)
        plan = BandPlan(h, rank, world, up, down)
        assert plan.halo_complete()
        full = M.synth_host(w, h, seed=11, dist=2)
        held = torch.full((plan.rows_held, w), float("nan"))
        held[plan.own_offset:plan.own_offset + plan.rows_own] = torch.from_numpy(full[plan.b0:plan.b1])
        exchange(held, plan, dist)
        got = held.numpy()
        want = full[plan.i0:plan.i1]
        same = np.array_equal(got.view(np.uint32), want.view(np.uint32))
        y_band = o.apply(op, e, got)[plan.own_offset:plan.own_offset + plan.rows_own]
        y_full = o.apply(op, e, full)[plan.b0:plan.b1]
        nan = np.isnan(y_full)
        inv = np.array_equal(np.isnan(y_band), nan) and \
            np.array_equal(y_band.view(np.uint32)[~nan], y_full.view(np.uint32)[~nan])
        q.put((rank, bool(same), bool(inv)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,element,op", [(2, "disk5", "tophat"), (2, "cross", "gradient"),
                                              (3, "disk3", "oscillation")])
def test_halo_exchange_and_sharding_invariance(world, element, op):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 37, 61, element, op, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == [(r, True, True) for r in range(world)]


def test_band_plan_covers_image_and_edges():
    h, up, down = 1000, 28, 28
    for world in (1, 2, 4, 8):
        plans = [BandPlan(h, r, world, up, down) for r in range(world)]
        assert plans[0].b0 == 0 and plans[-1].b1 == h
        for a, b in zip(plans, plans[1:]):
            assert a.b1 == b.b0
        assert plans[0].i0 == 0 and plans[-1].i1 == h            # image edges: no neighbour data
        for p in plans:
            sends = [t for t in p.transfers() if t[0] == "send"]
            recvs = [t for t in p.transfers() if t[0] == "recv"]
            assert len(sends) == len(recvs) == (p.rank > 0) + (p.rank < world - 1)
            assert sum(t[3] for t in recvs) == p.rows_held - p.rows_own
    assert not BandPlan(40, 1, 4, 28, 28).halo_complete()


def _handle_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from imscript_b200 import shard
        from imscript_b200.binding import SHARD_HANDLE_BYTES
        mine = bytes([rank + 1]) * SHARD_HANDLE_BYTES          # stands in for morsi_shard_handle()'s 128 bytes
        table = shard.gather_handles(mine, dist, world)
        ok = len(table) == world * SHARD_HANDLE_BYTES and \
            all(table[r * SHARD_HANDLE_BYTES:(r + 1) * SHARD_HANDLE_BYTES] == bytes([r + 1]) * SHARD_HANDLE_BYTES
                for r in range(world))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_shard_handles_are_gathered_in_rank_order():
    """the one piece of plumbing the C-level sharded path leaves to the caller: every rank ends up with the
    nranks x 128-byte handle table in rank order (world_size 3 over gloo)"""
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_handle_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == [(r, True) for r in range(world)]


def test_band_plan_matches_the_c_bookkeeping():
    """BandPlan restates morsi_shard_create / shard_push (shard.cu): owned rows h*r/N, held rows clipped to the
    image, the rows pushed to a neighbour are the first `down` / last `up` owned rows"""
    for (h, world, halo) in [(40000, 8, 28), (3001, 2, 28), (3001, 4, 12), (100, 3, 0)]:
        prev_b1 = 0
        for r in range(world):
            p = BandPlan(h, r, world, halo, halo)
            assert p.b0 == h * r // world == prev_b1 and p.b1 == h * (r + 1) // world
            assert p.i0 == max(0, p.b0 - halo) and p.i1 == min(h, p.b1 + halo)
            sends = {peer: (r0, n) for kind, peer, r0, n in p.transfers() if kind == "send"}
            if r > 0 and halo:
                assert sends[r - 1] == (p.own_offset, min(halo, p.rows_own))
            if r < world - 1 and halo:
                assert sends[r + 1] == (p.own_offset + p.rows_own - min(halo, p.rows_own), min(halo, p.rows_own))
            prev_b1 = p.b1
        assert prev_b1 == h
