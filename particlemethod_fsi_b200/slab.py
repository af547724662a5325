"""Multi-GPU host side: 1-D x-slab decomposition of the explicit step.

The reference has no distributed path (no MPI/NCCL anywhere in src/main.cpp); what is decomposed
here is its loop body (src/main.cpp:596-663).  The exchange itself -- migration, halo, PressureP of
the halo copies, the replicated solids, the rebuild votes -- lives in libmphx.so and is device-side:
every slab context owns a mailbox that its peers write into over NVLink, and `mphx_step` on a slab
context is a fixed sequence of kernel launches without a host synchronisation (csrc/slab.inc).
This module is only the plumbing around it:

  * `rebalance()`   (both classes) re-cuts a running ring in place: column histogram on the devices, one all-reduce, the
                    library's rule (mphx_rebalance_cuts), mphx_slab_recut -- owners change through the ordinary migration;
  * `plan()`        cuts the bucket columns on the particle histogram (mphx_partition_columns) and
                    sizes the slots / messages -- the same rule as mphx_multi_upload;
  * `DistSlab`      one process per GPU (torchrun): creates its slab context and carries the 64-byte
                    CUDA IPC handles of the mailboxes between the ranks with ONE torch.distributed
                    all_gather at start-up (NCCL on the GPU box, gloo in the CPU tests of this plumbing);
                    afterwards a step is a single C call per rank;
  * `MultiSolver`   one process, N contexts (mphx_multi_*: what csrc/main.cpp uses with MPHX_NGPU).  The
                    contexts may share a device, which is how the slab kernels are parity-tested on a
                    one-GPU box against the single-context result.

There is no Python or CPU implementation of any phase here.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time

import numpy as np

from . import abi

IPC_HANDLE_BYTES = 64   # sizeof(cudaIpcMemHandle_t)


# ---- host logic (no GPU needed; covered by the CPU tests) -----------------------------------------
def column_of(x, domain_min0: float, cell_width: float, ncols: int):
    """global bucket column of an x coordinate -- the reference's key expression (src/main.cpp:1671)"""
    c = np.floor((np.asarray(x, dtype=np.float64) - domain_min0) / cell_width).astype(np.int64) % ncols
    return ((c % ncols) + ncols) % ncols


def partition_columns(hist: np.ndarray, nranks: int, halo: int):
    """Cut the bucket columns [0, ncols) into `nranks` contiguous slabs of roughly equal particle
    count (mphx_partition_columns: the rule lives in the library so that the C++ driver, this module
    and the tests agree).  Returns [(lo, hi)] * nranks; raises ValueError when the halo-width
    constraints cannot be met."""
    from .solver import lib
    hist = np.ascontiguousarray(hist, dtype=np.int64)
    if nranks < 1:
        raise ValueError("nranks must be positive")
    cuts = (C.c_int * (nranks + 1))()
    rc = lib.mphx_partition_columns(hist.ctypes.data, int(hist.shape[0]), nranks, halo, C.cast(cuts, C.c_void_p))
    if rc != abi.MPHX_OK:
        raise ValueError(f"{int(hist.shape[0])} bucket columns cannot hold {nranks} slabs: {lib.mphx_last_error().decode()}")
    return [(cuts[r], cuts[r + 1]) for r in range(nranks)]


def ring_neighbours(rank: int, nranks: int):
    return (rank - 1) % nranks, (rank + 1) % nranks


def plan(case, world: int, constants=None):
    """column cuts, message capacity (the same on every rank) and slot capacity per rank -- mirrors
    mphx_multi_upload (csrc/slab.inc)"""
    from . import solver
    k = constants or solver.compute_constants(case.params)
    R, ncols = k.stencil_range, k.cell_count[0]
    t = case.property
    solid = (t >= 2) & (t < 4)
    ns = int(solid.sum())
    col = column_of(case.position[~solid, 0], case.params.domain_min[0], k.cell_width, ncols)
    hist = np.bincount(col, minlength=ncols).astype(np.int64)
    parts = partition_columns(hist, world, R)
    # a halo is R columns; allow twice the densest R-column window seen initially (+ slack)
    ring = np.concatenate([hist, hist[:R]])
    win = int(np.convolve(ring, np.ones(R, dtype=np.int64), mode="valid").max())
    msg_cap = int(min(case.n, max(2 * win + 1024, 4096)))
    caps = []
    for lo, hi in parts:
        owned = int(hist[lo:hi].sum())
        caps.append(int(min(case.n + 2 * msg_cap, int(1.5 * owned) + 4 * msg_cap + ns + 4096)))
    return dict(partition=parts, msg_capacity=msg_cap, capacity=caps, ns=ns, hist=hist)


def gather_handles(mine: bytes, dist=None, group=None, device=None):
    """all ranks' mailbox handles, in rank order (one all_gather of 64 bytes per rank)"""
    import torch
    import torch.distributed as tdist
    dist = dist or tdist
    world = dist.get_world_size(group)
    assert len(mine) == IPC_HANDLE_BYTES
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).clone()
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return [bytes(o.cpu().numpy().tobytes()) for o in out]


def rebalance_collective(local_hist, partition, halo: int, dist=None, group=None, device=None):
    """the collective half of an in-place re-balancing: all-reduce the ranks' column histograms (each rank counts the
    particles it owns), then every rank applies the library's rule (mphx_rebalance_cuts) to the same global histogram.
    Returns (new partition [(lo, hi)] * world, number of interior cuts that moved)."""
    import torch
    import torch.distributed as tdist
    from .solver import lib
    dist = dist or tdist
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(local_hist, dtype=np.int64).copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, group=group)
    hist = np.ascontiguousarray(t.cpu().numpy(), dtype=np.int64)
    old = np.ascontiguousarray([p[0] for p in partition] + [partition[-1][1]], dtype=np.int32)
    new = np.zeros(world + 1, dtype=np.int32)
    moved = C.c_int()
    rc = lib.mphx_rebalance_cuts(hist.ctypes.data, int(hist.shape[0]), world, halo, old.ctypes.data, new.ctypes.data, C.byref(moved))
    if rc != abi.MPHX_OK:
        raise RuntimeError(f"mphx_rebalance_cuts: {lib.mphx_last_error().decode()}")
    return [(int(new[r]), int(new[r + 1])) for r in range(world)], moved.value


def _views_for(n, names):
    out, hv = {}, abi.HostViews()
    for nm in names:
        shape, is_int = abi.VIEW_FIELDS[nm]
        a = np.zeros((n,) + shape, dtype=np.int32 if is_int else np.float64)
        setattr(hv, nm, a.ctypes.data_as(C.POINTER(C.c_int if is_int else C.c_double)))
        out[nm] = a
    return hv, out


# ---- one process, N contexts ------------------------------------------------------------------------
class MultiSolver:
    """mphx_multi_*: the explicit step on `world` x-slabs driven by ONE process.  devices=None puts slab r
    on CUDA device r; a list may repeat a device (parity tests on a one-GPU box)."""

    def __init__(self, case, world: int, devices=None, list_reuse: bool | None = None, generated: bool = False):
        """generated=True: the particles come from the device-side generator (case.cuboids: mphx_multi_upload_generated) --
        no particle array leaves the host"""
        from . import solver
        self.lib, self._ck = solver.lib, solver._ck
        self.case, self.n, self.world = case, case.n, world
        self._m = C.c_void_p()
        dev = None
        if devices is not None:
            assert len(devices) == world
            dev = (C.c_int * world)(*devices)
        self._ck("mphx_multi_create", self.lib.mphx_multi_create(C.byref(self._m), C.byref(case.params), world,
                                                                  C.cast(dev, C.c_void_p) if dev is not None else None))
        if list_reuse is not None:
            for r in range(world):
                self._ck("mphx_set_list_reuse", self.lib.mphx_set_list_reuse(self.context(r), 1 if list_reuse else 0, 0.0))
        if generated:
            arr = solver._cuboid_array(case.cuboids)
            self._ck("mphx_multi_upload_generated", self.lib.mphx_multi_upload_generated(self._m, C.cast(arr, C.c_void_p), len(case.cuboids)))
        else:
            t = np.ascontiguousarray(case.property, dtype=np.int32)
            x, x0, v = (np.ascontiguousarray(a, dtype=np.float64) for a in (case.position, case.initial_position, case.velocity))
            self._ck("mphx_multi_upload", self.lib.mphx_multi_upload(self._m, case.n, t.ctypes.data, x.ctypes.data, x0.ctypes.data, v.ctypes.data))
        self._ck("mphx_multi_init", self.lib.mphx_multi_init(self._m))

    def context(self, r: int):
        return C.c_void_p(self.lib.mphx_multi_context(self._m, r))

    def close(self):
        if self._m:
            self.lib.mphx_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, nsteps: int = 1):
        self._ck("mphx_multi_step", self.lib.mphx_multi_step(self._m, nsteps))

    def sync(self):
        self._ck("mphx_multi_sync", self.lib.mphx_multi_sync(self._m))

    def timed_steps(self, nsteps: int) -> float:
        ms = C.c_double()
        self._ck("mphx_multi_timed_steps", self.lib.mphx_multi_timed_steps(self._m, nsteps, C.byref(ms)))
        return ms.value

    @property
    def time(self) -> float:
        return self.lib.mphx_multi_time(self._m)

    def download(self, *names):
        hv, out = _views_for(self.n, names)
        self._ck("mphx_multi_download", self.lib.mphx_multi_download(self._m, C.byref(hv)))
        return out

    def info(self):
        res = []
        for r in range(self.world):
            a, st = (C.c_int * 4)(), (C.c_int * 8)()
            self._ck("mphx_slab_info", self.lib.mphx_slab_info(self.context(r), C.byref(a)))
            self._ck("mphx_get_status", self.lib.mphx_get_status(self.context(r), C.byref(st)))
            res.append(dict(rank=r, held=a[0], capacity=a[1], ghosts=a[2], msg_capacity=a[3], err=st[0], builds=st[2], reuses=st[3]))
        return res

    def rebalance(self) -> int:
        """re-cut the slabs on the current particle distribution (every interior cut moves by at most one halo width; the
        particles that change owner travel with the next step's migration); returns the number of cuts that moved"""
        moved = C.c_int()
        self._ck("mphx_multi_rebalance", self.lib.mphx_multi_rebalance(self._m, C.byref(moved)))
        return moved.value

    def columns(self):
        out = []
        for r in range(self.world):
            a = (C.c_int * 2)()
            self._ck("mphx_slab_columns", self.lib.mphx_slab_columns(self.context(r), C.byref(a)))
            out.append((a[0], a[1]))
        return out

    def download_owned(self, r: int):
        """(ids, position, velocity) rows of slab r (its owned fluid/wall particles + all replicated solids)"""
        cap = self.info()[r]["capacity"]
        ids = np.empty(cap, dtype=np.int32)
        x, v = np.empty((cap, 3)), np.empty((cap, 3))
        n = C.c_int()
        self._ck("mphx_download_owned", self.lib.mphx_download_owned(self.context(r), cap, ids.ctypes.data, x.ctypes.data, v.ctypes.data, C.byref(n)))
        return n.value, ids, x, v

    def upload_owned(self, r: int, count, ids, x, v):
        return self.lib.mphx_upload_owned(self.context(r), count, ids.ctypes.data, x.ctypes.data, v.ctypes.data)


# ---- one process per GPU (torchrun) -------------------------------------------------------------------
class DistSlab:
    """This rank's slab of the ring.  torch.distributed is used once, to all_gather the mailbox handles
    (and in download() to sum the per-rank reports); stepping is `mphx_step` on the slab context."""

    def __init__(self, case, device_index: int = 0, group=None, list_reuse: bool | None = None):
        import torch
        import torch.distributed as dist
        from . import solver
        self.lib, self._ck = solver.lib, solver._ck
        self.dist, self.group, self.torch = dist, group, torch
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.case, self.n = case, case.n
        self.device = torch.device("cuda", device_index)
        self.plan = plan(case, self.world)
        self.partition = self.plan["partition"]
        lo, hi = self.partition[self.rank]
        self.ctx = C.c_void_p()
        self._ck("mphx_create", self.lib.mphx_create(C.byref(self.ctx), C.byref(case.params), device_index))
        if list_reuse is not None:
            self._ck("mphx_set_list_reuse", self.lib.mphx_set_list_reuse(self.ctx, 1 if list_reuse else 0, 0.0))
        self._ck("mphx_slab_configure", self.lib.mphx_slab_configure(self.ctx, self.rank, self.world, lo, hi, self.plan["capacity"][self.rank],
                                                                     self.plan["msg_capacity"]))
        t = np.ascontiguousarray(case.property, dtype=np.int32)
        x, x0, v = (np.ascontiguousarray(a, dtype=np.float64) for a in (case.position, case.initial_position, case.velocity))
        self._ck("mphx_upload", self.lib.mphx_upload(self.ctx, case.n, t.ctypes.data, x.ctypes.data, x0.ctypes.data, v.ctypes.data))
        h = (C.c_ubyte * IPC_HANDLE_BYTES)()
        self._ck("mphx_slab_mailbox", self.lib.mphx_slab_mailbox(self.ctx, C.cast(h, C.c_void_p), None, None))
        handles = gather_handles(bytes(h), dist, group, self.device)      # (also: every rank has zeroed its mailbox by now)
        blob = (C.c_ubyte * (IPC_HANDLE_BYTES * self.world)).from_buffer_copy(b"".join(handles))
        # the ranks' GPUs as this process sees them: torchrun leaves every device visible, rank r uses LOCAL_RANK = r
        devs = self._gather_ints(device_index)
        darr = (C.c_int * self.world)(*devs)
        self._ck("mphx_slab_connect", self.lib.mphx_slab_connect(self.ctx, C.cast(blob, C.c_void_p), None, C.cast(darr, C.c_void_p)))
        dist.barrier(group)
        self._ck("mphx_init", self.lib.mphx_init(self.ctx))               # (collective: the halos are exchanged)
        dist.barrier(group)

    def _gather_ints(self, v: int):
        t = self.torch.tensor([v], dtype=self.torch.int32, device=self.device)
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t, group=self.group)
        return [int(o.item()) for o in out]

    def close(self):
        if self.ctx:
            self.lib.mphx_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def step(self, nsteps: int = 1):
        self._ck("mphx_step", self.lib.mphx_step(self.ctx, nsteps))

    def sync(self):
        self._ck("mphx_sync", self.lib.mphx_sync(self.ctx))

    def timed_steps(self, nsteps: int) -> float:
        """device milliseconds of `nsteps` steps on this rank (CUDA events on the context's stream)"""
        ms = C.c_double()
        self._ck("mphx_timed_steps", self.lib.mphx_timed_steps(self.ctx, nsteps, C.byref(ms)))
        return ms.value

    @property
    def time(self) -> float:
        return self.lib.mphx_time(self.ctx)

    @property
    def launch_count(self) -> int:
        return self.lib.mphx_launch_count(self.ctx)

    def status(self):
        st, a = (C.c_int * 8)(), (C.c_int * 4)()
        self._ck("mphx_get_status", self.lib.mphx_get_status(self.ctx, C.byref(st)))
        self._ck("mphx_slab_info", self.lib.mphx_slab_info(self.ctx, C.byref(a)))
        return dict(err=st[0], held=st[1], builds=st[2], reuses=st[3], ghosts=st[7], capacity=a[1], msg_capacity=a[3])

    def rebalance(self) -> int:
        """In-place re-balancing: all-reduce the per-column histogram of the owned particles, move every interior cut towards
        the balanced position by at most one halo width (mphx_rebalance_cuts: the rule lives in the library), and request the
        new columns from this rank's context; the particles that change owner travel with the next step's migration.
        Collective; returns the number of cuts that moved."""
        from . import solver
        k = solver.compute_constants(self.case.params)
        ncols, R = k.cell_count[0], k.stencil_range
        hist = np.zeros(ncols, dtype=np.int64)
        self._ck("mphx_slab_column_histogram", self.lib.mphx_slab_column_histogram(self.ctx, hist.ctypes.data, ncols))
        nccl = self.dist.get_backend(self.group) == "nccl"
        new, moved = rebalance_collective(hist, self.partition, R, self.dist, self.group, self.device if nccl else None)
        if moved:
            self._ck("mphx_slab_recut", self.lib.mphx_slab_recut(self.ctx, new[self.rank][0], new[self.rank][1]))
            self.partition = new
        return moved

    def download(self, *names):
        """every rank reports the particles it owns (zeros elsewhere); the all-reduced sum is the case"""
        hv, out = _views_for(self.n, names)
        self._ck("mphx_download", self.lib.mphx_download(self.ctx, C.byref(hv)))
        for nm in names:
            t = self.torch.from_numpy(out[nm]).to(self.device)
            self.dist.all_reduce(t, group=self.group)
            out[nm] = t.cpu().numpy()
        return out


def verify_ring(device_index: int, particles: float = 2.0e5, steps: int = 12, list_reuse: bool | None = None):
    """Correctness of the data plane that is being timed: the NVLink ring of slabs against ONE context on rank 0, on a
    down-scaled replica of the same case, same steps.  Returns (on rank 0) per-field bit-equality and the largest
    difference; every rank must call it."""
    import torch.distributed as dist
    from . import cases
    from .solver import Solver
    case = cases.fsi3d_for_count(particles)
    ring = DistSlab(case, device_index, list_reuse=list_reuse)
    ring.step(steps)
    ring.sync()
    fields = ("position", "velocity", "pressure_p", "cell_index")
    got = ring.download(*fields)
    st = ring.status()
    ring.close()
    res = None
    if dist.get_rank() == 0:
        ref = Solver.from_case(case, device=device_index, list_reuse=list_reuse)
        ref.step(steps, sync=True)
        want = ref.download(*fields)
        ref.close()
        res = {"particles": case.n, "steps": steps, "world": dist.get_world_size(), "fields": {}}
        ok = True
        for f in fields:
            a, b = want[f].astype(np.float64), got[f].astype(np.float64)
            d = float(np.abs(a - b).max())
            s = float(np.abs(a).max())
            eq = bool(np.array_equal(want[f], got[f]))
            res["fields"][f] = {"bit_equal": eq, "max_abs_diff": d, "max_norm_rel": d / s if s > 0 else d}
            ok = ok and (eq or d <= 1e-12 * max(s, 1e-300))
        res["ok"] = ok
        res["rank0_status"] = st
    dist.barrier()
    return res


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

    def allmax(v: float) -> float:
        t = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v: float) -> float:
        t = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(t)
        return float(t.item())

    # correctness of the exchange that is about to be timed (ring of slabs == one context, on a replica)
    verify = None if args.no_verify else verify_ring(local, particles=args.verify_particles)

    case = cases.fsi3d_for_count(args.particles)
    n = case.n
    nf, ns, nw = case.counts()
    s = DistSlab(case, local)
    K, W = args.steps, args.warmup
    s.step(W)
    s.sync()
    l0 = s.launch_count
    sampler = ClockSampler(local)
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    ms_local = s.timed_steps(K)          # CUDA events on the context's stream; ends after the last step's sub-steps
    torch.cuda.synchronize()
    dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = allmax(ms_local)
    # where a step's time goes on each rank (a few more steps with the phase events on; not part of the timed region)
    s._ck("mphx_set_timing", s.lib.mphx_set_timing(s.ctx, 1))
    s.step(min(K, 10))
    s.sync()
    kms = (C.c_double * 5)()
    s._ck("mphx_get_kernel_timers", s.lib.mphx_get_kernel_timers(s.ctx, C.byref(kms)))
    s._ck("mphx_set_timing", s.lib.mphx_set_timing(s.ctx, 0))
    pt = torch.tensor([v / min(K, 10) for v in kms], dtype=torch.float64, device=device)
    allp = [torch.empty_like(pt) for _ in range(world)]
    dist.all_gather(allp, pt)
    phases = [[round(float(v), 4) for v in t.cpu()] for t in allp]
    if getattr(args, "trace", ""):   # device-side timeline of 3 steps (diagnostic; outside every timed region)
        s._ck("mphx_trace_enable", s.lib.mphx_trace_enable(s.ctx, 8192))
        dist.barrier()
        s.step(3)
        s.sync()
        buf = np.zeros(2 * 8192, dtype=np.uint64)
        cnt = C.c_int()
        s._ck("mphx_trace_read", s.lib.mphx_trace_read(s.ctx, buf.ctypes.data, 8192, C.byref(cnt)))
        s._ck("mphx_trace_enable", s.lib.mphx_trace_enable(s.ctx, 0))
        marks = buf[: 2 * cnt.value].reshape(-1, 2)
        with open(f"{args.trace}_rank{rank}.json", "w") as fh:
            json.dump({"rank": rank, "world": world, "marks": [[int(a), int(b)] for a, b in marks]}, fh)
    launches = allsum(float(s.launch_count - l0))
    st = s.status()
    held = allmax(float(st["held"]))
    errs = allmax(float(st["err"]))
    builds, reuses = st["builds"], st["reuses"]

    # what was timed is a valid trajectory: every particle is still owned by exactly one rank, and the state is finite
    rows_cap = st["capacity"]
    hid = torch.empty(rows_cap, dtype=torch.int32).pin_memory()
    hx = torch.empty((rows_cap, 3), dtype=torch.float64).pin_memory()
    hv = torch.empty((rows_cap, 3), dtype=torch.float64).pin_memory()
    nrow = C.c_int()

    def fetch():
        s._ck("mphx_download_owned", s.lib.mphx_download_owned(
            s.ctx, rows_cap, C.c_void_p(hid.data_ptr()), C.c_void_p(hx.data_ptr()), C.c_void_p(hv.data_ptr()), C.byref(nrow)))

    fetch()
    m = nrow.value
    ids = hid[:m].numpy()
    fluidwall = ~((case.property[ids] >= 2) & (case.property[ids] < 4))
    owned_total = allsum(float(fluidwall.sum()))
    finite = allsum(float(not (bool(torch.isfinite(hx[:m]).all()) and bool(torch.isfinite(hv[:m]).all()))))
    chk_x = allsum(float(hx[:m][torch.from_numpy(fluidwall)].sum()))
    chk_v = allsum(float(hv[:m][torch.from_numpy(fluidwall)].abs().sum()))
    state_check = {"owned_fluid_wall_particles": int(owned_total), "expected": nf + nw, "all_finite": finite == 0.0,
                   "error_flags": int(errs), "sum_position": chk_x, "sum_abs_velocity": chk_v,
                   "ok": int(owned_total) == nf + nw and finite == 0.0 and int(errs) & ~32 == 0}

    # end to end through the C-ABI with HOST buffers, every step: mphx_upload_owned (ids, Position, Velocity
    # of the particles this rank owns + the replicated solids, from page-locked memory), one slab step,
    # mphx_download_owned back into page-locked memory.  Copies are inside the timed region; every rank
    # moves its own share over its own PCIe link.
    ke = max(1, min(K, args.e2e_steps))

    def e2e_step():
        s._ck("mphx_upload_owned", s.lib.mphx_upload_owned(
            s.ctx, nrow.value, C.c_void_p(hid.data_ptr()), C.c_void_p(hx.data_ptr()), C.c_void_p(hv.data_ptr())))
        s.step(1)
        fetch()

    e2e_step()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(ke):
        e2e_step()
    torch.cuda.synchronize()
    te = allmax(time.perf_counter() - t0)
    rows_total = allsum(float(nrow.value))
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
                           "cache": "inputs larger than L2",
                           "parallelism": f"{world} x-slabs (ring); migration, halo, PressureP and solid exchange are device-side stores "
                                          "into peer mailboxes over NVLink (CUDA IPC), no host synchronisation per step; NCCL carries "
                                          "only the start-up handles and the timing/verification reductions",
                           "partition_columns": s.partition, "max_slots_held": int(held),
                           "phase_ms_per_step_by_rank": {"columns": ["buckets+exchange", "filter", "pass1+P exchange", "pass2", "solid sub-steps (2nd stream)"],
                                                         "rows": phases},
                           "list_builds": builds, "list_reuses": reuses},
                "clocks": clocks,
                "e2e": {"value": n * ke / te, "unit": UNIT, "h2d_bytes_per_step": int(rows_total) * 52,
                        "d2h_bytes_per_step": int(rows_total) * 52, "steps": ke, "ms_per_step": 1e3 * te / ke,
                        "path": "per rank: mphx_upload_owned + slab step + mphx_download_owned (ids+Position+Velocity of the owned "
                                "particles and the replicated solids, pinned host buffers); bytes summed over ranks"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": "whole step (all ranks)", "achieved": agg, "peak": pk["hbm_gbs"] * world,
                             "unit": "GB/s", "frac": agg / (pk["hbm_gbs"] * world), "traffic": None,
                             "peak_source": pk_kind + " (MEASURED_PEAKS.json hbm_gbs x n_gpus)",
                             "algorithmic_bytes_per_step": step_bytes},
                "verify": {"ring_vs_single_context": verify, "timed_state": state_check},
                "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()
