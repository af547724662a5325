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


# ---- the step orchestration of slab.SlabSolver on CPU --------------------------------------------------------
class _ProtocolStandIn:
    """Stands in for libmphx.so's mphx_slab_* entry points: NO physics, only the contract of the five-phase
    protocol.  Every phase writes recognisable patterns with step- and rank-dependent counts into the
    buffers it is handed and checks that what arrives is exactly what the ring neighbours wrote: message
    sizes, left/right routing, the order of the phases and the two all-reduces."""

    def __init__(self, world):
        self.world, self.ctx, self.log = world, {}, []

    # -- helpers
    @staticmethod
    def _arr(ptr, n, ctype=None):
        import ctypes as C
        import numpy as np
        ctype = ctype or C.c_double
        return np.ctypeslib.as_array((ctype * max(n, 1)).from_address(ptr.value if hasattr(ptr, "value") else int(ptr)))[:n]

    @staticmethod
    def _counts(rank, step, phase):
        return (3 * step + rank + phase) % 5 + 1, (step + 2 * rank + 3 * phase) % 4 + 1     # (to_left, to_right)

    @staticmethod
    def _pattern(rank, step, phase, side, n, width):
        import numpy as np
        return np.arange(n * width, dtype=np.float64) + 1e6 * rank + 1e4 * step + 1e3 * phase + 1e2 * side

    # -- life cycle
    def mphx_create(self, ctx_ref, params_ref, device):
        h = len(self.ctx) + 1
        ctx_ref._obj.value = h
        self.ctx[h] = dict(step=0, phase="created")
        return 0

    def mphx_set_stream(self, ctx, stream): return 0

    def mphx_slab_configure(self, ctx, rank, world, lo, hi, cap, msg_cap):
        assert world == self.world and 0 <= rank < world and hi - lo >= 3
        self.ctx[ctx.value].update(rank=rank, msg_cap=msg_cap, cap=cap)
        return 0

    def mphx_upload(self, ctx, n, *a):
        self.ctx[ctx.value]["n"] = n
        return 0

    def mphx_init(self, ctx): return 0

    def mphx_destroy(self, ctx): self.ctx.pop(ctx.value, None)

    def mphx_time(self, ctx): return 1e-4 * self.ctx[ctx.value]["step"]

    def mphx_launch_count(self, ctx): return 0

    def mphx_join(self, ctx): return 0

    def mphx_slab_info(self, ctx, out_ref):
        c = self.ctx[ctx.value]
        out_ref._obj[0], out_ref._obj[1], out_ref._obj[2], out_ref._obj[3] = c["n"] // self.world, c["cap"], 0, c["msg_cap"]
        return 0

    # -- the five phases
    def _send(self, c, phase, left, right, counts, width):
        nl, nr = self._counts(c["rank"], c["step"], phase)
        assert max(nl, nr) <= c["msg_cap"]
        self._arr(left, nl * width)[:] = self._pattern(c["rank"], c["step"], phase, 0, nl, width)
        self._arr(right, nr * width)[:] = self._pattern(c["rank"], c["step"], phase, 1, nr, width)
        if counts is not None:
            import ctypes as C
            self._arr(counts, 4, C.c_int)[:] = [nl, nr, 0, 0]

    def _check_recv(self, c, phase, from_left, n_left, from_right, n_right, width):
        import numpy as np
        l, r = slab.ring_neighbours(c["rank"], self.world)
        # the left neighbour's "to_right" message arrives from the left, the right neighbour's "to_left" from the right
        assert n_left == self._counts(l, c["step"], phase)[1] and n_right == self._counts(r, c["step"], phase)[0], (phase, n_left, n_right)
        assert np.array_equal(self._arr(from_left, n_left * width), self._pattern(l, c["step"], phase, 1, n_left, width))
        assert np.array_equal(self._arr(from_right, n_right * width), self._pattern(r, c["step"], phase, 0, n_right, width))

    def mphx_slab_begin(self, ctx, left, right, counts):
        c = self.ctx[ctx.value]
        assert c["phase"] in ("created", "finish")
        self._send(c, 0, left, right, counts, slab.MSG_DOUBLES)
        c["phase"] = "begin"
        return 0

    def mphx_slab_append(self, ctx, from_left, n_left, from_right, n_right, ghost):
        c = self.ctx[ctx.value]
        assert (c["phase"], ghost) in (("begin", 0), ("halo", 1))
        self._check_recv(c, 0 if ghost == 0 else 1, from_left, n_left, from_right, n_right, slab.MSG_DOUBLES)
        if ghost:
            c["ghosts"] = (n_left, n_right)
        c["phase"] = "migrated" if ghost == 0 else "ghosts"
        return 0

    def mphx_slab_pack_halo(self, ctx, left, right, counts):
        c = self.ctx[ctx.value]
        assert c["phase"] == "migrated"
        self._send(c, 1, left, right, counts, slab.MSG_DOUBLES)
        c["phase"] = "halo"
        return 0

    def mphx_slab_build_pass1(self, ctx, n_left, n_right, left, right, solP):
        c = self.ctx[ctx.value]
        assert c["phase"] == "ghosts" and (n_left, n_right) == self._counts(c["rank"], c["step"], 1)
        # PressureP of the halo particles: same counts as the halo messages, one double each
        self._arr(left, n_left)[:] = self._pattern(c["rank"], c["step"], 1, 0, n_left, 1) + 0.5
        self._arr(right, n_right)[:] = self._pattern(c["rank"], c["step"], 1, 1, n_right, 1) + 0.5
        self._arr(solP, 4)[:] = c["rank"] + 1.0
        c["phase"] = "pass1"
        return 0

    def mphx_slab_pass2(self, ctx, from_left, from_right, solP, solbuf):
        import numpy as np
        c = self.ctx[ctx.value]
        assert c["phase"] == "pass1"
        l, r = slab.ring_neighbours(c["rank"], self.world)
        gl, gr = c["ghosts"]
        assert np.array_equal(self._arr(from_left, gl), self._pattern(l, c["step"], 1, 1, gl, 1) + 0.5)
        assert np.array_equal(self._arr(from_right, gr), self._pattern(r, c["step"], 1, 0, gr, 1) + 0.5)
        assert np.all(self._arr(solP, 4) == self.world * (self.world + 1) / 2.0)       # all-reduced over the ring
        self._arr(solbuf, 4)[:] = 10.0 * (c["rank"] + 1)
        c["phase"] = "pass2"
        return 0

    def mphx_slab_finish(self, ctx, solbuf):
        import numpy as np
        c = self.ctx[ctx.value]
        assert c["phase"] == "pass2"
        assert np.all(self._arr(solbuf, 4) == 10.0 * self.world * (self.world + 1) / 2.0)
        c["phase"] = "finish"
        c["step"] += 1
        self.log.append((c["rank"], c["step"]))
        return 0


def _orchestrate(transport, world, steps=4):
    from particlemethod_fsi_b200 import cases
    case = cases.tiny2d()                       # has solids: both all-reduces are exercised
    fake = _ProtocolStandIn(world)
    s = slab.SlabSolver(case, transport, device=torch.device("cpu"), lib=fake)
    s.step(steps)
    assert abs(s.time - steps * 1e-4) < 1e-12
    assert all(c["phase"] == "finish" and c["step"] == steps for c in fake.ctx.values())
    s.close()
    return len(fake.log)


@pytest.mark.parametrize("world", [2, 3])
def test_slab_step_orchestration_local_ring(world):
    assert _orchestrate(slab.LocalRing(world), world) == 4 * world


def _orchestration_worker(rank, world, port, ok):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert _orchestrate(slab.DistTransport(), world) == 4
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_step_orchestration_over_gloo(world):
    ok = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_orchestration_worker, args=(world, _free_port(), ok), nprocs=world, join=True)
    assert list(ok) == [1] * world
