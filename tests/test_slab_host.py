"""Host side of the multi-GPU slab protocol on CPU: column partitioning and the ring transports
(LocalRing in-process, DistTransport over gloo with world_size 2 and 3).  No compute calls."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from particlemethod_fsi_b200 import slab


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


def test_column_of_matches_reference_key_expression():
    x = np.array([-0.05, -0.0499999, 0.0, 0.35 - 1e-12, 0.1234567])
    c = slab.column_of(x, -0.05, 1e-3, 400)
    assert list(c) == [0, 0, 50, 399, int(np.floor((0.1234567 + 0.05) / 1e-3))]


def _payload(rank, side, n):
    return torch.arange(n * slab.MSG_DOUBLES, dtype=torch.float64) + 1000.0 * rank + 100.0 * side


def _check_ring(tr_items_fn, world):
    """every slab sends n_left = rank+1 particles to its left and n_right = 2*rank+3 to its right"""
    sent = [(r + 1, 2 * r + 3) for r in range(world)]
    for r in range(world):
        l, rt = slab.ring_neighbours(r, world)
        got_counts, rl, rr = tr_items_fn(r)
        assert got_counts == (sent[l][1], sent[rt][0])
        assert torch.equal(rl[: got_counts[0] * slab.MSG_DOUBLES], _payload(l, 1, sent[l][1]))
        assert torch.equal(rr[: got_counts[1] * slab.MSG_DOUBLES], _payload(rt, 0, sent[rt][0]))


@pytest.mark.parametrize("world", [2, 3, 4])
def test_local_ring_exchange(world):
    tr = slab.LocalRing(world)
    cap = 64
    counts = [torch.tensor([r + 1, 2 * r + 3, 0, 0], dtype=torch.int32) for r in range(world)]
    got = [tuple(int(v) for v in c) for c in tr.exchange_counts(counts)]
    items = []
    for r in range(world):
        sl = torch.zeros(cap * slab.MSG_DOUBLES, dtype=torch.float64)
        sr = torch.zeros(cap * slab.MSG_DOUBLES, dtype=torch.float64)
        sl[: (r + 1) * slab.MSG_DOUBLES] = _payload(r, 0, r + 1)
        sr[: (2 * r + 3) * slab.MSG_DOUBLES] = _payload(r, 1, 2 * r + 3)
        items.append((sl, r + 1, sr, 2 * r + 3, torch.zeros_like(sl), got[r][0], torch.zeros_like(sr), got[r][1], slab.MSG_DOUBLES))
    tr.exchange(items)
    _check_ring(lambda r: (got[r], items[r][4], items[r][6]), world)
    ts = [torch.full((5,), float(r + 1), dtype=torch.float64) for r in range(world)]
    tr.allreduce_sum(ts)
    assert all(torch.equal(t, torch.full((5,), world * (world + 1) / 2.0, dtype=torch.float64)) for t in ts)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, ok):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tr = slab.DistTransport()
        assert tr.local_ranks() == [rank] and tr.world == world
        cap = 64
        counts = torch.tensor([rank + 1, 2 * rank + 3, 0, 0], dtype=torch.int32)
        got = tuple(int(v) for v in tr.exchange_counts([counts])[0])
        sl = torch.zeros(cap * slab.MSG_DOUBLES, dtype=torch.float64)
        sr = torch.zeros(cap * slab.MSG_DOUBLES, dtype=torch.float64)
        sl[: (rank + 1) * slab.MSG_DOUBLES] = _payload(rank, 0, rank + 1)
        sr[: (2 * rank + 3) * slab.MSG_DOUBLES] = _payload(rank, 1, 2 * rank + 3)
        rl, rr = torch.zeros_like(sl), torch.zeros_like(sr)
        tr.exchange([(sl, rank + 1, sr, 2 * rank + 3, rl, got[0], rr, got[1], slab.MSG_DOUBLES)])
        l, rt = slab.ring_neighbours(rank, world)
        assert got == (2 * l + 3, rt + 1), got
        assert torch.equal(rl[: got[0] * slab.MSG_DOUBLES], _payload(l, 1, 2 * l + 3))
        assert torch.equal(rr[: got[1] * slab.MSG_DOUBLES], _payload(rt, 0, rt + 1))
        t = torch.full((7,), float(rank + 1), dtype=torch.float64)
        tr.allreduce_sum([t])
        assert torch.equal(t, torch.full((7,), world * (world + 1) / 2.0, dtype=torch.float64))
        assert tr.allreduce_max_float(float(rank), torch.device("cpu")) == float(world - 1)
        # zero-length messages (an idle seam) must not dead-lock
        tr.exchange([(sl, 0, sr, 0, rl, 0, rr, 0, slab.MSG_DOUBLES)])
        tr.barrier()
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_dist_transport_over_gloo(world):
    ok = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_gloo_worker, args=(world, _free_port(), ok), nprocs=world, join=True)
    assert list(ok) == [1] * world
