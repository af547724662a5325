"""Multi-GPU host side: 1-D x-slab decomposition of the explicit step.

The reference has no distributed path (no MPI/NCCL anywhere in src/main.cpp); what is decomposed
here is its loop body (src/main.cpp:596-663).  One process drives one B200 and one slab context
of libmphx.so; this module only moves the packed device buffers between ranks:

    phase A  mphx_slab_begin       pre-step + emigrants packed   -> ring exchange -> mphx_slab_append(ghost=0)
    phase B  mphx_slab_pack_halo   halo layers packed            -> ring exchange -> mphx_slab_append(ghost=1)
    phase C  mphx_slab_build_pass1 buckets + pass 1              -> ring exchange of PressureP, all-reduce(solP)
    phase D  mphx_slab_pass2       pass 2 + integration          -> all-reduce(solbuf)
    phase E  mphx_slab_finish      solid sub-steps (replicated), Time += Dt

The exchanges are torch.distributed point-to-point operations (NCCL over NVLink on the GPU box,
gloo on CPU for the host-logic tests); every axis of the reference's domain is periodic
(CellId wrap src/main.cpp:123-125), so the slabs form a ring.  All compute is in the CUDA library:
there is no Python or CPU implementation of any phase here.

`LocalRing` runs all slabs of a ring inside ONE process on ONE device (the exchange is a device
copy): it is how the slab kernels are parity-tested on a single-GPU box against the single-context
result.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time

import numpy as np

from . import abi

MSG_DOUBLES = 7  # kMsgDoubles in csrc/kernels.cuh: x y z vx vy vz (type<<32|id)


# ---- host logic (no GPU needed; covered by the gloo tests) -----------------------------------------
def column_of(x, domain_min0: float, cell_width: float, ncols: int):
    """global bucket column of an x coordinate -- the reference's key expression (src/main.cpp:1671)"""
    c = np.floor((np.asarray(x, dtype=np.float64) - domain_min0) / cell_width).astype(np.int64) % ncols
    return ((c % ncols) + ncols) % ncols


def partition_columns(hist: np.ndarray, nranks: int, halo: int):
    """Cut the bucket columns [0, ncols) into `nranks` contiguous slabs of roughly equal particle
    count.  Every slab is at least `halo` columns wide (a particle's stencil must not reach past the
    neighbouring slab) and the whole ring must be wider than one slab plus its two halos.
    Returns [(lo, hi)] * nranks."""
    hist = np.asarray(hist, dtype=np.int64)
    ncols = int(hist.shape[0])
    if nranks < 1:
        raise ValueError("nranks must be positive")
    if nranks == 1:
        return [(0, ncols)]
    # every slab at most ncols - 2*halo wide <=> the other slabs together span >= 2*halo columns
    minw = halo if nranks >= 3 else 2 * halo
    if ncols < nranks * minw:
        raise ValueError(f"{ncols} bucket columns cannot hold {nranks} slabs of >= {minw} columns")
    if int(hist.sum()) == 0:
        hist = np.ones(ncols, dtype=np.int64)        # nothing to balance (e.g. a solid-only case): equal widths
    cum = np.concatenate([[0], np.cumsum(hist)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, nranks):
        target = total * r / nranks
        c = int(np.searchsorted(cum, target, side="left"))
        lo = cuts[-1] + minw                         # previous slab wide enough
        hi = ncols - (nranks - r) * minw             # room for the remaining slabs
        cuts.append(min(max(c, lo), hi))
    cuts.append(ncols)
    out = [(cuts[r], cuts[r + 1]) for r in range(nranks)]
    for lo, hi in out:
        if hi - lo < halo or ncols < (hi - lo) + 2 * halo:
            raise ValueError("slab partition violates the halo-width constraint")
    return out


def ring_neighbours(rank: int, nranks: int):
    return (rank - 1) % nranks, (rank + 1) % nranks


# ---- transports ----------------------------------------------------------------------------------
class DistTransport:
    """torch.distributed ring (one slab per process).  Works on CUDA tensors with the nccl backend and
    on CPU tensors with gloo (used by the CPU tests of this plumbing)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.left, self.right = ring_neighbours(self.rank, self.world)

    def local_ranks(self):
        return [self.rank]

    def _p2p(self, ops):
        for w in self.dist.batch_isend_irecv(ops):
            w.wait()

    def exchange_counts(self, counts_list):
        """counts_list[0]: int tensor [>=2] = (to_left, to_right) -> [(from_left, from_right)]"""
        import torch
        dist = self.dist
        c = counts_list[0]
        recv = torch.zeros(2, dtype=c.dtype, device=c.device)
        sl, sr = c[0:1].clone(), c[1:2].clone()
        # order matters when left == right (2 ranks): a message sent "to the left" arrives at its
        # receiver "from the right", so receives are posted right-first
        ops = [dist.P2POp(dist.isend, sl, self.left, self.group), dist.P2POp(dist.isend, sr, self.right, self.group),
               dist.P2POp(dist.irecv, recv[1:2], self.right, self.group), dist.P2POp(dist.irecv, recv[0:1], self.left, self.group)]
        self._p2p(ops)
        return [recv]

    def exchange(self, items):
        """items[0] = (send_left, n_left, send_right, n_right, recv_left, m_left, recv_right, m_right, width):
        send the first n*width elements of each send buffer, receive m*width into the recv buffers"""
        dist = self.dist
        sl, nl, sr, nr, rl, ml, rr, mr, w = items[0]
        ops = []
        if nl > 0:
            ops.append(dist.P2POp(dist.isend, sl[: nl * w], self.left, self.group))
        if nr > 0:
            ops.append(dist.P2POp(dist.isend, sr[: nr * w], self.right, self.group))
        if mr > 0:
            ops.append(dist.P2POp(dist.irecv, rr[: mr * w], self.right, self.group))
        if ml > 0:
            ops.append(dist.P2POp(dist.irecv, rl[: ml * w], self.left, self.group))
        if ops:
            self._p2p(ops)

    def allreduce_sum(self, tensors):
        self.dist.all_reduce(tensors[0], op=self.dist.ReduceOp.SUM, group=self.group)

    def allreduce_max_float(self, v: float, device) -> float:
        import torch
        t = torch.tensor([v], dtype=torch.float64, device=device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def barrier(self):
        self.dist.barrier(group=self.group)


class LocalRing:
    """all `world` slabs in this process on one device: the exchange is a device-to-device copy"""

    def __init__(self, world: int):
        self.world = world

    def local_ranks(self):
        return list(range(self.world))

    def exchange_counts(self, counts_list):
        import torch
        out = []
        for r in range(self.world):
            l, rt = ring_neighbours(r, self.world)
            # from_left = what the left neighbour sent to ITS right; from_right = what the right sent to its left
            out.append(torch.stack([counts_list[l][1], counts_list[rt][0]]))
        return out

    def exchange(self, items):
        for r in range(self.world):
            l, rt = ring_neighbours(r, self.world)
            _sl, _nl, _sr, _nr, rl, ml, rr, mr, w = items[r]
            if ml > 0:
                rl[: ml * w].copy_(items[l][2][: ml * w])    # left neighbour's send_right
            if mr > 0:
                rr[: mr * w].copy_(items[rt][0][: mr * w])   # right neighbour's send_left

    def allreduce_sum(self, tensors):
        import torch
        tot = torch.stack(list(tensors)).sum(dim=0)
        for t in tensors:
            t.copy_(tot)

    def allreduce_max_float(self, v: float, device) -> float:
        return v

    def barrier(self):
        pass


# ---- one slab context --------------------------------------------------------------------------------
class _Slab:
    def __init__(self, lib, case_params, rank, world, cols, capacity, msg_cap, ns, device, torch):
        self.lib, self.rank, self.world = lib, rank, world
        self.ctx = C.c_void_p()
        from .solver import _ck
        self._ck = _ck
        _ck("mphx_create", lib.mphx_create(C.byref(self.ctx), C.byref(case_params), device.index or 0))
        stream = torch.cuda.current_stream(device).cuda_stream if device.type == "cuda" else 0  # (cpu: protocol tests)
        _ck("mphx_set_stream", lib.mphx_set_stream(self.ctx, C.c_void_p(stream)))
        _ck("mphx_slab_configure", lib.mphx_slab_configure(self.ctx, rank, world, cols[0], cols[1], capacity, msg_cap))
        f64 = dict(dtype=torch.float64, device=device)
        self.send = [torch.zeros(MSG_DOUBLES * msg_cap, **f64) for _ in range(2)]
        self.recv = [torch.zeros(MSG_DOUBLES * msg_cap, **f64) for _ in range(2)]
        self.psend = [torch.zeros(msg_cap, **f64) for _ in range(2)]
        self.precv = [torch.zeros(msg_cap, **f64) for _ in range(2)]
        self.counts = torch.zeros(4, dtype=torch.int32, device=device)
        self.solP = torch.zeros(max(ns, 1), **f64)
        self.solbuf = torch.zeros(6 * max(ns, 1), **f64)
        self.halo_sent = (0, 0)
        self.halo_recv = (0, 0)

    def close(self):
        if self.ctx:
            self.lib.mphx_destroy(self.ctx)
            self.ctx = C.c_void_p()


class SlabSolver:
    """The explicit step on `world` x-slabs.  With a DistTransport every process holds one slab
    (rank = torch.distributed rank); with a LocalRing this object holds them all."""

    def __init__(self, case, transport, device=None, capacity: int | None = None, msg_capacity: int | None = None, lib=None):
        import torch
        from . import solver
        self.torch = torch
        self.lib = lib if lib is not None else solver.lib   # (lib: a stand-in with the mphx_slab_* contract, for the
        #                                                      CPU tests of this orchestration; never a compute path)
        self.tr = transport
        self.case = case
        self.n = case.n
        world = transport.world
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        k = solver.compute_constants(case.params)
        self.constants = k
        R = k.stencil_range
        ncols = k.cell_count[0]
        t = case.property
        solid = (t >= 2) & (t < 4)
        self.ns = int(solid.sum())
        col = column_of(case.position[~solid, 0], case.params.domain_min[0], k.cell_width, ncols)
        hist = np.bincount(col, minlength=ncols)
        self.partition = partition_columns(hist, world, R)
        per_col = k.cell_count[1] * k.cell_count[2]
        if msg_capacity is None:
            # a halo is R columns; allow twice the densest R-column window seen initially (+ slack)
            win = np.convolve(hist, np.ones(R, dtype=np.int64), mode="full").max()
            msg_capacity = int(min(self.n, max(2 * win + 1024, 4096)))
        self.slabs = []
        for r in transport.local_ranks():
            lo, hi = self.partition[r]
            owned = int(hist[lo:hi].sum())
            cap = capacity
            if cap is None:
                cap = int(min(self.n + 2 * msg_capacity, 1.5 * owned + 4 * msg_capacity + self.ns + 4096))
            s = _Slab(self.lib, case.params, r, world, (lo, hi), cap, msg_capacity, self.ns, device, torch)
            solver._ck("mphx_upload", self.lib.mphx_upload(
                s.ctx, case.n, np.ascontiguousarray(case.property, dtype=np.int32).ctypes.data,
                np.ascontiguousarray(case.position).ctypes.data, np.ascontiguousarray(case.initial_position).ctypes.data,
                np.ascontiguousarray(case.velocity).ctypes.data))
            solver._ck("mphx_init", self.lib.mphx_init(s.ctx))
            self.slabs.append(s)
        self._per_col = per_col
        # host mirror of the wall centres (src/main.cpp:3066-3070: advanced every step), carried across rebalance()
        self._wall_center = [[case.params.wall_center[t][d] for d in range(3)] for t in range(abi.TYPE_COUNT)]
        self._msg_capacity_arg, self._capacity_arg = None, capacity

    def close(self):
        for s in self.slabs:
            s.close()
        self.slabs = []

    # -- one step --------------------------------------------------------------------------------------
    def _ptr(self, t):
        return C.c_void_p(t.data_ptr())

    def _counts(self, what):
        """Own (to_left, to_right) and received (from_left, from_right) counts of every local slab with ONE
        device->host read per slab: the neighbours' counts are exchanged device-side first."""
        recv = self.tr.exchange_counts([s.counts for s in self.slabs])
        sent, got = [], []
        for s, r in zip(self.slabs, recv):
            c = self.torch.cat([s.counts[:3], r.to(s.counts.dtype)]).cpu().numpy()
            if c[2]:
                raise RuntimeError(f"slab {s.rank}: {what}: exchange error flags {int(c[2])} "
                                   "(1: a particle crossed more than one halo width in a step, 2: message buffer too small)")
            sent.append((int(c[0]), int(c[1])))
            got.append((int(c[3]), int(c[4])))
        return sent, got

    def step(self, nsteps: int = 1):
        lib, tr, ck = self.lib, self.tr, self.slabs[0]._ck
        for _ in range(nsteps):
            # A: pre-step + migration
            for s in self.slabs:
                ck("mphx_slab_begin", lib.mphx_slab_begin(s.ctx, self._ptr(s.send[0]), self._ptr(s.send[1]), self._ptr(s.counts)))
            sent, got = self._counts("migration")
            tr.exchange([(s.send[0], sent[i][0], s.send[1], sent[i][1], s.recv[0], got[i][0], s.recv[1], got[i][1], MSG_DOUBLES)
                         for i, s in enumerate(self.slabs)])
            for i, s in enumerate(self.slabs):
                ck("mphx_slab_append", lib.mphx_slab_append(s.ctx, self._ptr(s.recv[0]), got[i][0], self._ptr(s.recv[1]), got[i][1], 0))
            # B: halo
            for s in self.slabs:
                ck("mphx_slab_pack_halo", lib.mphx_slab_pack_halo(s.ctx, self._ptr(s.send[0]), self._ptr(s.send[1]), self._ptr(s.counts)))
            sent, got = self._counts("halo")
            tr.exchange([(s.send[0], sent[i][0], s.send[1], sent[i][1], s.recv[0], got[i][0], s.recv[1], got[i][1], MSG_DOUBLES)
                         for i, s in enumerate(self.slabs)])
            for i, s in enumerate(self.slabs):
                s.halo_sent, s.halo_recv = sent[i], got[i]
                ck("mphx_slab_append", lib.mphx_slab_append(s.ctx, self._ptr(s.recv[0]), got[i][0], self._ptr(s.recv[1]), got[i][1], 1))
            # C: buckets + pass 1, PressureP of the halo copies and of the replicated solids
            for s in self.slabs:
                ck("mphx_slab_build_pass1", lib.mphx_slab_build_pass1(s.ctx, s.halo_sent[0], s.halo_sent[1], self._ptr(s.psend[0]),
                                                                      self._ptr(s.psend[1]), self._ptr(s.solP)))
            tr.exchange([(s.psend[0], s.halo_sent[0], s.psend[1], s.halo_sent[1], s.precv[0], s.halo_recv[0], s.precv[1],
                          s.halo_recv[1], 1) for s in self.slabs])
            if self.ns > 0:
                tr.allreduce_sum([s.solP for s in self.slabs])
            # D: pass 2 + integration
            for s in self.slabs:
                ck("mphx_slab_pass2", lib.mphx_slab_pass2(s.ctx, self._ptr(s.precv[0]), self._ptr(s.precv[1]), self._ptr(s.solP),
                                                          self._ptr(s.solbuf)))
            if self.ns > 0:
                tr.allreduce_sum([s.solbuf for s in self.slabs])
            # E: solid sub-steps
            for s in self.slabs:
                ck("mphx_slab_finish", lib.mphx_slab_finish(s.ctx, self._ptr(s.solbuf)))
            p = self.case.params
            for t in range(4, abi.TYPE_COUNT):
                for d in range(3):
                    self._wall_center[t][d] += p.wall_velocity[t][d] * p.dt

    def imbalance(self) -> float:
        """largest / mean number of particle slots held by a slab (1.0 = perfectly balanced)"""
        held = [i["held"] for i in self.info()]
        if isinstance(self.tr, DistTransport) and self.tr.world > 1:
            t = self.torch.tensor([float(held[0])], dtype=self.torch.float64, device=self.device)
            mx = self.tr.allreduce_max_float(float(held[0]), self.device)
            self.tr.allreduce_sum([t])
            return mx / (t.item() / self.tr.world)
        return max(held) / (sum(held) / len(held))

    def rebalance(self):
        """Re-cut the slabs on the CURRENT particle distribution (a dam break empties some
        slabs and fills others).  The state is gathered through the hosts and the slab contexts are rebuilt
        from it: a coarse operation meant for every few hundred steps.  The continuation is bit-identical
        (same in-bucket order, same sums); Time and the wall centres carry over."""
        from . import cases
        full = self.download("position", "velocity")
        case = self.case
        p = case.params.copy()
        p.time0 = self.time
        for t in range(abi.TYPE_COUNT):
            for d in range(3):
                p.wall_center[t][d] = self._wall_center[t][d]
        new_case = cases.Case(case.name, p, case.rc, case.property, full["position"], case.initial_position, full["velocity"],
                              getattr(case, "cuboids", []))
        tr, device, cap, lib = self.tr, self.device, self._capacity_arg, self.lib
        old = self.partition
        self.close()
        self.__init__(new_case, tr, device=device, capacity=cap, lib=lib)
        return old, self.partition

    def join(self):
        """the contexts' streams wait (device-side) for the sub-steps still running on the internal streams"""
        for s in self.slabs:
            s._ck("mphx_join", self.lib.mphx_join(s.ctx))

    def sync(self):
        if self.device.type == "cuda":
            self.torch.cuda.synchronize(self.device)

    def info(self):
        out = []
        for s in self.slabs:
            a = (C.c_int * 4)()
            self.lib.mphx_slab_info(s.ctx, C.byref(a))
            out.append(dict(rank=s.rank, held=a[0], capacity=a[1], ghosts=a[2], msg_capacity=a[3], columns=self.partition[s.rank]))
        return out

    @property
    def launch_count(self) -> int:
        return sum(self.lib.mphx_launch_count(s.ctx) for s in self.slabs)

    @property
    def time(self) -> float:
        return self.lib.mphx_time(self.slabs[0].ctx)

    # -- download: every slab reports the particles it owns (zeros elsewhere); the sum is the case -----
    def download(self, *names):
        torch = self.torch
        out = {}
        for nm in names:
            shape, is_int = abi.VIEW_FIELDS[nm]
            tot = None
            for s in self.slabs:
                a = np.zeros((self.n,) + shape, dtype=np.int32 if is_int else np.float64)
                hv = abi.HostViews()
                setattr(hv, nm, a.ctypes.data_as(C.POINTER(C.c_int if is_int else C.c_double)))
                s._ck("mphx_download", self.lib.mphx_download(s.ctx, C.byref(hv)))
                tot = a if tot is None else tot + a
            if isinstance(self.tr, DistTransport) and self.tr.world > 1:
                t = torch.from_numpy(tot).to(self.device)
                self.tr.allreduce_sum([t])
                tot = t.cpu().numpy()
            out[nm] = tot
        return out


# ---- bench.py, N > 1 (launched by torchrun: one rank per GPU) -------------------------------------------
def bench_main(args, METRIC, UNIT, WORKLOAD, peaks, ClockSampler, cpu_reference_run, host_cores):
    import torch
    import torch.distributed as dist
    from . import cases

    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    tr = DistTransport()
    case = cases.fsi3d_for_count(args.particles)
    n = case.n
    nf, ns, nw = case.counts()
    s = SlabSolver(case, tr, device=device)
    K, W = args.steps, args.warmup
    s.step(W)
    s.sync()
    l0 = s.launch_count
    sampler = ClockSampler(local)
    tr.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s.step(K)
    s.join()   # the last step's solid sub-steps run on the library's second stream: the timed region ends after them
    e1.record()
    torch.cuda.synchronize()
    ms_local = e0.elapsed_time(e1)
    tr.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = tr.allreduce_max_float(ms_local, device)
    launches = s.launch_count - l0
    lt = torch.tensor([float(launches)], dtype=torch.float64, device=device)
    dist.all_reduce(lt)
    info = s.info()[0]
    held = torch.tensor([float(info["held"])], dtype=torch.float64, device=device)
    dist.all_reduce(held, op=dist.ReduceOp.MAX)

    # end to end through the C-ABI with HOST buffers, every step: mphx_upload_owned (ids, Position, Velocity
    # of the particles this rank owns + the replicated solids, from page-locked memory), one slab step,
    # mphx_download_owned back into page-locked memory.  Copies are inside the timed region; every rank
    # moves its own share over its own PCIe link.
    ctx0 = s.slabs[0].ctx
    rows_cap = s.info()[0]["capacity"]
    hid = torch.empty(rows_cap, dtype=torch.int32).pin_memory()
    hx = torch.empty((rows_cap, 3), dtype=torch.float64).pin_memory()
    hv = torch.empty((rows_cap, 3), dtype=torch.float64).pin_memory()
    ke = max(1, min(K, args.e2e_steps))
    nrow = C.c_int()

    def fetch():
        s.slabs[0]._ck("mphx_download_owned", s.lib.mphx_download_owned(
            ctx0, rows_cap, C.c_void_p(hid.data_ptr()), C.c_void_p(hx.data_ptr()), C.c_void_p(hv.data_ptr()), C.byref(nrow)))

    def e2e_step():
        s.slabs[0]._ck("mphx_upload_owned", s.lib.mphx_upload_owned(
            ctx0, nrow.value, C.c_void_p(hid.data_ptr()), C.c_void_p(hx.data_ptr()), C.c_void_p(hv.data_ptr())))
        s.step(1)
        fetch()

    fetch()
    e2e_step()
    tr.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(ke):
        e2e_step()
    torch.cuda.synchronize()
    te = tr.allreduce_max_float(time.perf_counter() - t0, device)
    rt = torch.tensor([float(nrow.value)], dtype=torch.float64, device=device)
    dist.all_reduce(rt)
    rows_total = rt.item()
    s.close()
    if rank == 0:
        pk, pk_kind = peaks()
        nsub = int(case.params.dt / case.params.elastic_dt + 0.5)
        step_bytes = 368.0 * nf + 260.0 * nw + (344.0 + 384.0 * nsub) * ns
        agg = step_bytes * K / (ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": n * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "particles": n, "fluid": nf, "solid": ns, "wall": nw, "dim": 3,
                           "particle_spacing": case.params.particle_spacing, "dt": case.params.dt, "solid_substeps": nsub,
                           "cache": "inputs larger than L2", "parallelism": f"{world} x-slabs (ring), halo + migration over NCCL",
                           "partition_columns": s.partition, "max_slots_held": int(held.item())},
                "clocks": clocks,
                "e2e": {"value": n * ke / te, "unit": UNIT, "h2d_bytes_per_step": int(rows_total) * 52,
                        "d2h_bytes_per_step": int(rows_total) * 52, "steps": ke, "ms_per_step": 1e3 * te / ke,
                        "path": "per rank: mphx_upload_owned + slab step + mphx_download_owned (ids+Position+Velocity of the owned "
                                "particles and the replicated solids, pinned host buffers); bytes summed over ranks"},
                "gpu_launches": int(lt.item()),
                "roofline": {"bound": "hbm", "kernel": "whole step (all ranks)", "achieved": agg, "peak": pk["hbm_gbs"] * world,
                             "unit": "GB/s", "frac": agg / (pk["hbm_gbs"] * world), "traffic": None,
                             "peak_source": pk_kind + " (MEASURED_PEAKS.json hbm_gbs x n_gpus)",
                             "algorithmic_bytes_per_step": step_bytes},
                "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()
