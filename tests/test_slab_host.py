"""Host side of the multi-GPU slab path on CPU: column partitioning (mphx_partition_columns through the
C-ABI: host code, no GPU), the sizing plan every rank derives, and the one collective the Python side
performs -- carrying the 64-byte mailbox handles between the ranks -- over gloo with world_size 2 and 3.
No compute calls: the exchange itself is device-side inside libmphx.so (tests/test_gpu_slab.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from particlemethod_fsi_b200 import cases, slab


def test_partition_balances_and_respects_halo():
    rng = np.random.default_rng(12345)
    hist = np.zeros(200, dtype=np.int64)
    hist[20:90] = rng.integers(500, 1500, 70)      # a water column on the left, like a dam break
    hist[150:155] = 50                              # far wall
    for nranks in (2, 3, 4, 8):
        parts = slab.partition_columns(hist, nranks, 3)
        assert parts[0][0] == 0 and parts[-1][1] == 200
        for (a, b), (c, d) in zip(parts[:-1], parts[1:]):
            assert b == c
        for lo, hi in parts:
            assert hi - lo >= 3 and 200 >= (hi - lo) + 6
        loads = [int(hist[lo:hi].sum()) for lo, hi in parts]
        assert max(loads) <= 1.0 * hist.sum() / nranks + hist.max() * 3 + 1
    assert slab.partition_columns(hist, 1, 3) == [(0, 200)]


def test_partition_degenerate_inputs():
    with pytest.raises(ValueError):
        slab.partition_columns(np.ones(10), 4, 3)       # 4 slabs of >= 3 columns do not fit in 10
    parts = slab.partition_columns(np.zeros(64), 4, 3)  # empty domain: still a valid cover
    assert parts[0][0] == 0 and parts[-1][1] == 64 and all(hi - lo >= 3 for lo, hi in parts)
    one_spike = np.zeros(64)
    one_spike[5] = 1000
    parts = slab.partition_columns(one_spike, 4, 3)
    assert all(hi - lo >= 3 for lo, hi in parts)
    with pytest.raises(ValueError):
        slab.partition_columns(np.ones(11), 2, 3)       # two slabs: each at most ncols - 2*halo wide -> >= 12 columns


def test_column_of_matches_reference_key_expression():
    x = np.array([-0.05, -0.0499999, 0.0, 0.35 - 1e-12, 0.1234567])
    c = slab.column_of(x, -0.05, 1e-3, 400)
    assert list(c) == [0, 0, 50, 399, int(np.floor((0.1234567 + 0.05) / 1e-3))]


@pytest.mark.parametrize("world", [2, 3, 8])
def test_plan_is_consistent(world):
    """every rank derives the same cuts and message capacity from the case alone; slots cover the owned particles,
    both halos and all replicated solids"""
    case = cases.fsi3d_mini()
    p = slab.plan(case, world)
    assert p["partition"][0][0] == 0 and len(p["partition"]) == world
    nf, ns, nw = case.counts()
    assert p["ns"] == ns and int(p["hist"].sum()) == nf + nw
    for (lo, hi), cap in zip(p["partition"], p["capacity"]):
        owned = int(p["hist"][lo:hi].sum())
        assert cap >= owned + 2 * p["msg_capacity"] + ns or cap >= case.n
    q = slab.plan(case, world)
    assert all(np.array_equal(p[k], q[k]) if k == "hist" else p[k] == q[k] for k in p)
    assert slab.ring_neighbours(0, world) == (world - 1, 1 % world)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, ok):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = bytes((17 * rank + i) % 256 for i in range(slab.IPC_HANDLE_BYTES))
        got = slab.gather_handles(mine)
        assert len(got) == world
        for r in range(world):
            assert got[r] == bytes((17 * r + i) % 256 for i in range(slab.IPC_HANDLE_BYTES))
        # the plan is computed independently by every rank: it must agree without communication
        p = slab.plan(cases.tiny2d(), world)
        t = torch.tensor([c for lohi in p["partition"] for c in lohi] + [p["msg_capacity"]], dtype=torch.int64)
        ref = t.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(t, ref)
        dist.barrier()
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_handles_travel_over_gloo(world):
    ok = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_gloo_worker, args=(world, _free_port(), ok), nprocs=world, join=True)
    assert list(ok) == [1] * world


def test_rebalance_cuts_move_towards_balance_by_at_most_one_halo():
    """mphx_rebalance_cuts: interior cuts move towards the balanced position of the CURRENT histogram by at most one halo
    width per call, never below the minimum slab width, the periodic seam (cuts[0], cuts[n]) stays put; repeated calls
    converge on mphx_partition_columns' cuts"""
    import ctypes as C
    from particlemethod_fsi_b200.solver import lib
    ncols, world, R = 120, 4, 3
    hist = np.zeros(ncols, dtype=np.int64)
    hist[10:50] = 1000            # all particles sit in the first half now
    old = np.array([0, 30, 60, 90, 120], dtype=np.int32)
    target = np.array(slab.partition_columns(hist, world, R)).ravel()[[0, 2, 4, 6, 7]]
    cuts = old.copy()
    for it in range(40):
        new = np.zeros(world + 1, dtype=np.int32)
        moved = C.c_int()
        rc = lib.mphx_rebalance_cuts(hist.ctypes.data, ncols, world, R, cuts.ctypes.data, new.ctypes.data, C.byref(moved))
        assert rc == 0
        assert new[0] == 0 and new[-1] == ncols
        assert np.all(np.abs(new - cuts) <= R)
        assert np.all(np.diff(new) >= R)
        assert moved.value == int((new != cuts).sum())
        if moved.value == 0:
            break
        cuts = new
    assert np.array_equal(cuts, target), (cuts, target)
    # an already balanced ring does not move
    new = np.zeros(world + 1, dtype=np.int32)
    moved = C.c_int(7)
    assert lib.mphx_rebalance_cuts(hist.ctypes.data, ncols, world, R, cuts.ctypes.data, new.ctypes.data, C.byref(moved)) == 0
    assert moved.value == 0 and np.array_equal(new, cuts)


def _rebalance_worker(rank, world, port, ok):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ncols, R = 96, 3
        full = np.zeros(ncols, dtype=np.int64)
        full[8:40] = 700                      # the fluid has moved into the first slabs' columns
        width = ncols // world
        part = [(r * width, (r + 1) * width if r < world - 1 else ncols) for r in range(world)]
        mine = np.zeros(ncols, dtype=np.int64)
        lo, hi = part[rank]
        mine[lo:hi] = full[lo:hi]             # every rank counts the particles it owns
        new, moved = slab.rebalance_collective(mine, part, R)
        # the same answer as one process with the whole histogram, on every rank
        import ctypes as C
        from particlemethod_fsi_b200.solver import lib
        old = np.array([p[0] for p in part] + [ncols], dtype=np.int32)
        want = np.zeros(world + 1, dtype=np.int32)
        m = C.c_int()
        assert lib.mphx_rebalance_cuts(full.ctypes.data, ncols, world, R, old.ctypes.data, want.ctypes.data, C.byref(m)) == 0
        assert [c for lohi in new for c in lohi] == [int(c) for r in range(world) for c in (want[r], want[r + 1])]
        assert moved == m.value and moved > 0
        assert all(abs(new[r][0] - part[r][0]) <= R and abs(new[r][1] - part[r][1]) <= R for r in range(world))
        t = torch.tensor([c for lohi in new for c in lohi], dtype=torch.int64)
        ref = t.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(t, ref)
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_rebalance_collective_over_gloo(world):
    """the collective half of DistSlab.rebalance(): per-rank histograms all-reduced over gloo, the library's rule applied by
    every rank -- identical cuts everywhere, equal to the single-process result"""
    ok = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_rebalance_worker, args=(world, _free_port(), ok), nprocs=world, join=True)
    assert list(ok) == [1] * world
