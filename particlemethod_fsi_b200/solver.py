"""Host-side binding of libmphx.so (the extern-"C" layer of include/mphx.h) -- plumbing only.

There is no Python or CPU implementation of the step here: every compute call goes to the CUDA
library, and importing this module fails loudly when the library has not been built.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPHX_LIB") or os.path.join(_HERE, "libmphx.so")   # MPHX_LIB: developer override (kernel variants)


class MphxError(RuntimeError):
    def __init__(self, what: str, code: int, lib=None):
        msg = f"{what}: error {code}"
        if lib is not None:
            msg = f"{what}: {lib.mphx_strerror(code).decode()} [{code}] {lib.mphx_last_error().decode()}"
        super().__init__(msg)
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C particlemethod_fsi_b200/csrc`). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.mphx_version.restype = C.c_int
    L.mphx_strerror.argtypes = [C.c_int]
    L.mphx_strerror.restype = C.c_char_p
    L.mphx_last_error.restype = C.c_char_p
    L.mphx_device_count.restype = C.c_int
    L.mphx_params_default.argtypes = [C.POINTER(abi.Params), C.POINTER(abi.RunControl)]
    L.mphx_params_default.restype = None
    L.mphx_read_data_file.argtypes = [C.c_char_p, C.POINTER(abi.Params), C.POINTER(abi.RunControl), vp, vp]
    L.mphx_read_grid_file.argtypes = [C.c_char_p, C.POINTER(abi.Params), ip, C.POINTER(ip), C.POINTER(dp),
                                      C.POINTER(dp), C.POINTER(dp)]
    L.mphx_free_host.argtypes = [vp]
    L.mphx_free_host.restype = None
    L.mphx_write_prof_file.argtypes = [C.c_char_p, C.c_double, C.POINTER(abi.Params), C.c_int, vp, vp, vp, vp]
    L.mphx_write_vtk_file.argtypes = [C.c_char_p, C.c_int, vp, C.POINTER(abi.HostViews)]
    L.mphx_write_checkpoint.argtypes = [C.c_char_p, C.c_double, C.POINTER(abi.Params), C.c_int, vp, vp, vp, vp]
    L.mphx_read_checkpoint.argtypes = [C.c_char_p, C.POINTER(abi.Params), ip, C.POINTER(ip), C.POINTER(dp), C.POINTER(dp), C.POINTER(dp)]
    L.mphx_get_wall_centers.argtypes = [vp, vp]
    L.mphx_class_ranges.argtypes = [C.c_int, vp, C.POINTER(C.c_int * 6)]
    L.mphx_class_ranges.restype = None
    L.mphx_compute_constants.argtypes = [C.POINTER(abi.Params), C.POINTER(abi.Constants)]
    L.mphx_create.argtypes = [C.POINTER(vp), C.POINTER(abi.Params), C.c_int]
    L.mphx_destroy.argtypes = [vp]
    L.mphx_destroy.restype = None
    L.mphx_upload.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    L.mphx_generate_count.argtypes = [vp, C.c_int]
    L.mphx_generate_count.restype = C.c_longlong
    L.mphx_upload_generated.argtypes = [vp, vp, C.c_int]
    L.mphx_read_boid_file.argtypes = [C.c_char_p, vp, C.POINTER(vp), ip]
    L.mphx_generate_column_histogram.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_int, vp]
    L.mphx_multi_upload_generated.argtypes = [vp, vp, C.c_int]
    L.mphx_upload_state.argtypes = [vp, vp, vp]
    L.mphx_init.argtypes = [vp]
    L.mphx_get_constants.argtypes = [vp, C.POINTER(abi.Constants)]
    L.mphx_step.argtypes = [vp, C.c_int]
    L.mphx_step_fluid_only.argtypes = [vp]
    L.mphx_sync.argtypes = [vp]
    L.mphx_time.argtypes = [vp]
    L.mphx_time.restype = C.c_double
    L.mphx_set_time.argtypes = [vp, C.c_double]
    L.mphx_download.argtypes = [vp, C.POINTER(abi.HostViews)]
    L.mphx_download_owned.argtypes = [vp, C.c_int, vp, vp, vp, ip]
    L.mphx_upload_owned.argtypes = [vp, C.c_int, vp, vp, vp]
    L.mphx_debug_neighbors.argtypes = [vp, vp, vp, C.c_longlong]
    L.mphx_debug_initial_structure_neighbors.argtypes = [vp, vp, vp, C.c_longlong]
    L.mphx_timed_steps.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    L.mphx_set_timing.argtypes = [vp, C.c_int]
    L.mphx_get_timers.argtypes = [vp, C.POINTER(C.c_double * 4)]
    L.mphx_get_kernel_timers.argtypes = [vp, C.POINTER(C.c_double * 5)]
    L.mphx_trace_enable.argtypes = [vp, C.c_int]
    L.mphx_trace_read.argtypes = [vp, vp, C.c_int, ip]
    L.mphx_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    L.mphx_count_pairs.argtypes = [vp, C.POINTER(C.c_ulonglong * 2)]
    L.mphx_get_virial_ms.argtypes = [vp]
    L.mphx_get_virial_ms.restype = C.c_double
    L.mphx_set_overlap.argtypes = [vp, C.c_int]
    L.mphx_join.argtypes = [vp]
    L.mphx_launch_count.argtypes = [vp]
    L.mphx_launch_count.restype = C.c_longlong
    L.mphx_algorithmic_bytes_per_step.argtypes = [vp]
    L.mphx_algorithmic_bytes_per_step.restype = C.c_double
    L.mphx_set_list_reuse.argtypes = [vp, C.c_int, C.c_double]
    L.mphx_get_status.argtypes = [vp, C.POINTER(C.c_int * 8)]
    L.mphx_set_stream.argtypes = [vp, vp]
    L.mphx_partition_columns.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    L.mphx_slab_configure.argtypes = [vp] + [C.c_int] * 6
    L.mphx_slab_mailbox.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(C.c_longlong)]
    L.mphx_slab_connect.argtypes = [vp, vp, vp, vp]
    L.mphx_slab_info.argtypes = [vp, C.POINTER(C.c_int * 4)]
    L.mphx_slab_column_histogram.argtypes = [vp, vp, C.c_int]
    L.mphx_rebalance_cuts.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, ip]
    L.mphx_slab_recut.argtypes = [vp, C.c_int, C.c_int]
    L.mphx_slab_columns.argtypes = [vp, C.POINTER(C.c_int * 2)]
    L.mphx_multi_rebalance.argtypes = [vp, ip]
    L.mphx_multi_create.argtypes = [C.POINTER(vp), C.POINTER(abi.Params), C.c_int, vp]
    L.mphx_multi_destroy.argtypes = [vp]
    L.mphx_multi_destroy.restype = None
    L.mphx_multi_count.argtypes = [vp]
    L.mphx_multi_context.argtypes = [vp, C.c_int]
    L.mphx_multi_context.restype = vp
    L.mphx_multi_upload.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    L.mphx_multi_init.argtypes = [vp]
    L.mphx_multi_step.argtypes = [vp, C.c_int]
    L.mphx_multi_sync.argtypes = [vp]
    L.mphx_multi_time.argtypes = [vp]
    L.mphx_multi_time.restype = C.c_double
    L.mphx_multi_download.argtypes = [vp, C.POINTER(abi.HostViews)]
    L.mphx_multi_timed_steps.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    return L


class _LazyLib:
    """libmphx.so is mapped on first use, not on import: `bench.py --impl reference` imports
    `cases` through this package and must not load the product library.  A missing library is still
    a loud ImportError -- at the first call instead of at import."""
    _L = None

    def __getattr__(self, name):
        L = _LazyLib._L
        if L is None:
            L = _LazyLib._L = _load()
        return getattr(L, name)


lib = _LazyLib()


def _ck(what: str, rc: int):
    if rc != abi.MPHX_OK:
        raise MphxError(what, rc, lib)


# ---- file formats / host constants (no GPU needed) ---------------------------------------------
def read_data_file(fn: str, dim: int = 2, module: int = abi.MODULE_BAR):
    """readDataFile (src/main.cpp:729-786) -> (Params, RunControl, [invalid lines])"""
    p, rc = abi.Params(), abi.RunControl()
    lib.mphx_params_default(C.byref(p), C.byref(rc))
    p.dim, p.clamp_module = dim, module
    bad = []
    CB = C.CFUNCTYPE(None, C.c_char_p, C.c_void_p)
    cb = CB(lambda line, _u: bad.append(line.decode(errors="replace")))
    _ck("mphx_read_data_file", lib.mphx_read_data_file(fn.encode(), C.byref(p), C.byref(rc), C.cast(cb, C.c_void_p), None))
    return p, rc, bad


def read_grid_file(fn: str, p: abi.Params):
    """readGridFile (src/main.cpp:788-929): fills p.time0/spacing/domain, returns (type, x, x0, v)"""
    n = C.c_int()
    t = C.POINTER(C.c_int)()
    x, x0, v = (C.POINTER(C.c_double)() for _ in range(3))
    _ck("mphx_read_grid_file", lib.mphx_read_grid_file(fn.encode(), C.byref(p), C.byref(n), C.byref(t), C.byref(x),
                                                        C.byref(x0), C.byref(v)))
    N = n.value
    try:
        T = np.ctypeslib.as_array(t, shape=(max(N, 1),))[:N].copy()
        X = np.ctypeslib.as_array(x, shape=(max(N, 1), 3))[:N].copy()
        X0 = np.ctypeslib.as_array(x0, shape=(max(N, 1), 3))[:N].copy()
        V = np.ctypeslib.as_array(v, shape=(max(N, 1), 3))[:N].copy()
    finally:
        for q in (t, x, x0, v):
            lib.mphx_free_host(C.cast(q, C.c_void_p))
    return T, X, X0, V


def write_prof_file(fn: str, time: float, p: abi.Params, property, position, initial_position, velocity):
    t = np.ascontiguousarray(property, dtype=np.int32)
    x, x0, v = (np.ascontiguousarray(a, dtype=np.float64) for a in (position, initial_position, velocity))
    _ck("mphx_write_prof_file", lib.mphx_write_prof_file(fn.encode(), time, C.byref(p), t.shape[0], t.ctypes.data,
                                                         x.ctypes.data, x0.ctypes.data, v.ctypes.data))


def write_checkpoint(fn: str, time: float, p: abi.Params, property, position, initial_position, velocity):
    """lossless binary checkpoint (mphx_write_checkpoint)"""
    t = np.ascontiguousarray(property, dtype=np.int32)
    x, x0, v = (np.ascontiguousarray(a, dtype=np.float64) for a in (position, initial_position, velocity))
    _ck("mphx_write_checkpoint", lib.mphx_write_checkpoint(fn.encode(), time, C.byref(p), t.shape[0], t.ctypes.data, x.ctypes.data,
                                                           x0.ctypes.data, v.ctypes.data))


def read_checkpoint(fn: str, p: abi.Params):
    """fills time0 / dim / spacing / domain / wall_center of p; returns (type, x, x0, v)"""
    n = C.c_int()
    t = C.POINTER(C.c_int)()
    x, x0, v = (C.POINTER(C.c_double)() for _ in range(3))
    _ck("mphx_read_checkpoint", lib.mphx_read_checkpoint(fn.encode(), C.byref(p), C.byref(n), C.byref(t), C.byref(x), C.byref(x0), C.byref(v)))
    N = n.value
    try:
        T = np.ctypeslib.as_array(t, shape=(N,)).copy()
        X, X0, V = (np.ctypeslib.as_array(q, shape=(N, 3)).copy() for q in (x, x0, v))
    finally:
        for q in (t, x, x0, v):
            lib.mphx_free_host(C.cast(q, C.c_void_p))
    return T, X, X0, V


def _views(n: int, fields: dict):
    hv = abi.HostViews()
    keep = {}
    for name, arr in fields.items():
        shape, is_int = abi.VIEW_FIELDS[name]
        a = np.ascontiguousarray(arr, dtype=np.int32 if is_int else np.float64)
        assert a.shape == (n,) + shape, (name, a.shape)
        keep[name] = a
        setattr(hv, name, a.ctypes.data_as(C.POINTER(C.c_int if is_int else C.c_double)))
    return hv, keep


def write_vtk_file(fn: str, initial_position, fields: dict):
    x0 = np.ascontiguousarray(initial_position, dtype=np.float64)
    hv, _keep = _views(x0.shape[0], fields)
    _ck("mphx_write_vtk_file", lib.mphx_write_vtk_file(fn.encode(), x0.shape[0], x0.ctypes.data, C.byref(hv)))


def compute_constants(p: abi.Params) -> abi.Constants:
    k = abi.Constants()
    _ck("mphx_compute_constants", lib.mphx_compute_constants(C.byref(p), C.byref(k)))
    return k


def class_ranges(property) -> list:
    t = np.ascontiguousarray(property, dtype=np.int32)
    r = (C.c_int * 6)()
    lib.mphx_class_ranges(t.shape[0], t.ctypes.data, C.byref(r))
    return list(r)


def _cuboid_array(cuboids):
    arr = (abi.CuboidC * len(cuboids))()
    for q, cb in enumerate(cuboids):
        arr[q].type, arr[q].spacing = int(cb.type), float(cb.spacing)
        for d in range(3):
            arr[q].lower[d], arr[q].upper[d], arr[q].velocity[d] = float(cb.lower[d]), float(cb.upper[d]), float(cb.velocity[d])
    return arr


def read_boid_file(fn: str, params):
    """the pre-processor's input: fills params' time0 / spacing / domain, returns [cases.Cuboid] (mphx_read_boid_file)"""
    from . import cases
    ptr, n = C.c_void_p(), C.c_int()
    _ck("mphx_read_boid_file", lib.mphx_read_boid_file(fn.encode(), C.byref(params), C.byref(ptr), C.byref(n)))
    arr = C.cast(ptr, C.POINTER(abi.CuboidC * n.value)).contents
    out = [cases.Cuboid(int(c.type), tuple(c.lower), tuple(c.upper), float(c.spacing), tuple(c.velocity)) for c in arr]
    lib.mphx_free_host(ptr)
    return out


def generate_count(cuboids) -> int:
    """particles a list of cases.Cuboid holds under the generator's lattice rule (host only)"""
    return int(lib.mphx_generate_count(C.cast(_cuboid_array(cuboids), C.c_void_p), len(cuboids)))


def device_count() -> int:
    return lib.mphx_device_count()


def measure_fp64_peak(device: int = 0) -> float:
    """dense FP64 FMA throughput of the device in TFLOP/s (measured, CUDA events)"""
    t = C.c_double()
    _ck("mphx_measure_fp64_peak", lib.mphx_measure_fp64_peak(device, C.byref(t)))
    return t.value


def _addr(a):
    """host address of a numpy array or a torch (pinned) tensor"""
    return C.c_void_p(a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data)


# ---- the solver context ---------------------------------------------------------------------------
VTK_FIELDS = ("property", "position", "velocity", "force", "acceleration", "stress", "strain",
              "neighbor_count", "initial_structure_neighbor_count")


class Solver:
    """One context = one B200.  Mirrors the reference's sequence: create (readDataFile/readGridFile
    results), upload (`acc update device`), init (initialize* + the calls of src/main.cpp:564-570),
    step (loop body :596-663), download (`acc update host`)."""

    def __init__(self, params: abi.Params, device: int = 0):
        self.params = params
        self._ctx = C.c_void_p()
        _ck("mphx_create", lib.mphx_create(C.byref(self._ctx), C.byref(params), device))
        self.n = 0

    @classmethod
    def from_case(cls, case, device: int = 0, init: bool = True, list_reuse: bool | None = None, skin: float = 0.0):
        s = cls(case.params, device)
        if list_reuse is not None or skin > 0.0:
            s.set_list_reuse(True if list_reuse is None else list_reuse, skin)
        s.upload(case.property, case.position, case.initial_position, case.velocity)
        if init:
            s.init()
        return s

    def close(self):
        if self._ctx:
            lib.mphx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, property, position, initial_position, velocity):
        t = np.ascontiguousarray(property, dtype=np.int32)
        x, x0, v = (np.ascontiguousarray(a, dtype=np.float64) for a in (position, initial_position, velocity))
        assert x.shape == (t.shape[0], 3) and x0.shape == x.shape and v.shape == x.shape
        _ck("mphx_upload", lib.mphx_upload(self._ctx, t.shape[0], t.ctypes.data, x.ctypes.data, x0.ctypes.data,
                                           v.ctypes.data))
        self.n = t.shape[0]

    def upload_generated(self, cuboids):
        """device-side generator: cuboids = cases.Cuboid list in file order (mphx_upload_generated)"""
        arr = _cuboid_array(cuboids)
        _ck("mphx_upload_generated", lib.mphx_upload_generated(self._ctx, C.cast(arr, C.c_void_p), len(cuboids)))
        self.n = int(lib.mphx_generate_count(C.cast(arr, C.c_void_p), len(cuboids)))

    def upload_state(self, position, velocity):
        """per-step H2D of Position/Velocity (original order); arrays must stay alive until sync"""
        assert position.dtype == np.float64 and velocity.dtype == np.float64
        assert position.flags.c_contiguous and velocity.flags.c_contiguous
        _ck("mphx_upload_state", lib.mphx_upload_state(self._ctx, position.ctypes.data, velocity.ctypes.data))

    def upload_state_ptr(self, position_ptr: int, velocity_ptr: int):
        _ck("mphx_upload_state", lib.mphx_upload_state(self._ctx, position_ptr, velocity_ptr))

    def download_ptr(self, **ptrs):
        """download into raw host pointers (e.g. torch pinned tensors): name=int address"""
        hv = abi.HostViews()
        for nm, ptr in ptrs.items():
            _shape, is_int = abi.VIEW_FIELDS[nm]
            setattr(hv, nm, C.cast(ptr, C.POINTER(C.c_int if is_int else C.c_double)))
        _ck("mphx_download", lib.mphx_download(self._ctx, C.byref(hv)))

    def download_owned(self, ids, position, velocity) -> int:
        """compact (ids, Position, Velocity) of the particles this context owns into caller arrays
        (numpy, or anything with .ctypes.data / an int address); returns the row count"""
        n = C.c_int()
        cap = int(ids.shape[0])
        _ck("mphx_download_owned", lib.mphx_download_owned(self._ctx, cap, _addr(ids), _addr(position), _addr(velocity), C.byref(n)))
        return n.value

    def upload_owned(self, count: int, ids, position, velocity):
        _ck("mphx_upload_owned", lib.mphx_upload_owned(self._ctx, count, _addr(ids), _addr(position), _addr(velocity)))

    def init(self):
        _ck("mphx_init", lib.mphx_init(self._ctx))

    def count_pairs(self):
        """(candidates in the current lists, pairs within the largest kernel radius)"""
        out = (C.c_ulonglong * 2)()
        _ck("mphx_count_pairs", lib.mphx_count_pairs(self._ctx, C.byref(out)))
        return int(out[0]), int(out[1])

    def set_list_reuse(self, on: bool, skin: float = 0.0):
        """candidate-list reuse (internal Verlet skin); skin in particle spacings, only before upload"""
        _ck("mphx_set_list_reuse", lib.mphx_set_list_reuse(self._ctx, 1 if on else 0, float(skin)))

    def status(self) -> dict:
        a = (C.c_int * 8)()
        _ck("mphx_get_status", lib.mphx_get_status(self._ctx, C.byref(a)))
        return dict(err=a[0], slots=a[1], builds=a[2], reuses=a[3], age=a[4], skin_on=a[5], solid_multi_occupancy=a[6], ghosts=a[7])

    def wall_centers(self):
        a = ((C.c_double * 3) * abi.TYPE_COUNT)()
        _ck("mphx_get_wall_centers", lib.mphx_get_wall_centers(self._ctx, C.cast(a, C.c_void_p)))
        return np.array([[a[t][d] for d in range(3)] for t in range(abi.TYPE_COUNT)])

    def checkpoint(self, fn: str, property, initial_position):
        """write the state of this context (Position, Velocity, Time, wall centres) to a lossless checkpoint"""
        g = self.download("position", "velocity")
        p = self.params.copy()
        wc = self.wall_centers()
        for t in range(abi.TYPE_COUNT):
            for d in range(3):
                p.wall_center[t][d] = wc[t][d]
        write_checkpoint(fn, self.time, p, property, g["position"], initial_position, g["velocity"])

    def constants(self) -> abi.Constants:
        k = abi.Constants()
        _ck("mphx_get_constants", lib.mphx_get_constants(self._ctx, C.byref(k)))
        return k

    def step(self, nsteps: int = 1, sync: bool = False):
        _ck("mphx_step", lib.mphx_step(self._ctx, nsteps))
        if sync:
            self.sync()

    def step_fluid_only(self):
        _ck("mphx_step_fluid_only", lib.mphx_step_fluid_only(self._ctx))

    def sync(self):
        _ck("mphx_sync", lib.mphx_sync(self._ctx))

    @property
    def time(self) -> float:
        return lib.mphx_time(self._ctx)

    @time.setter
    def time(self, t: float):
        _ck("mphx_set_time", lib.mphx_set_time(self._ctx, t))

    def download(self, *names, out: dict | None = None) -> dict:
        fields = {}
        for nm in names:
            shape, is_int = abi.VIEW_FIELDS[nm]
            if out is not None and nm in out:
                fields[nm] = out[nm]
            else:
                fields[nm] = np.empty((self.n,) + shape, dtype=np.int32 if is_int else np.float64)
        hv, keep = _views(self.n, fields)
        _ck("mphx_download", lib.mphx_download(self._ctx, C.byref(hv)))
        return keep

    def neighbors(self):
        """(offsets[N+1], ids) -- neighbour sets of calculateNeighbor, rows sorted ascending"""
        off = np.zeros(self.n + 1, dtype=np.int64)
        _ck("mphx_debug_neighbors", lib.mphx_debug_neighbors(self._ctx, off.ctypes.data, None, 0))
        ids = np.empty(max(int(off[-1]), 1), dtype=np.int32)
        _ck("mphx_debug_neighbors", lib.mphx_debug_neighbors(self._ctx, off.ctypes.data, ids.ctypes.data, ids.shape[0]))
        return off, ids[: int(off[-1])]

    def initial_structure_neighbors(self):
        off = np.zeros(self.n + 1, dtype=np.int64)
        _ck("mphx_debug_initial_structure_neighbors",
            lib.mphx_debug_initial_structure_neighbors(self._ctx, off.ctypes.data, None, 0))
        ids = np.empty(max(int(off[-1]), 1), dtype=np.int32)
        _ck("mphx_debug_initial_structure_neighbors",
            lib.mphx_debug_initial_structure_neighbors(self._ctx, off.ctypes.data, ids.ctypes.data, ids.shape[0]))
        return off, ids[: int(off[-1])]

    def timed_steps(self, nsteps: int) -> float:
        """device milliseconds (CUDA events on the context's stream) of `nsteps` steps"""
        ms = C.c_double()
        _ck("mphx_timed_steps", lib.mphx_timed_steps(self._ctx, nsteps, C.byref(ms)))
        return ms.value

    def set_timing(self, on: bool):
        _ck("mphx_set_timing", lib.mphx_set_timing(self._ctx, 1 if on else 0))

    def timers_ms(self):
        ms = (C.c_double * 4)()
        _ck("mphx_get_timers", lib.mphx_get_timers(self._ctx, C.byref(ms)))
        return list(ms)

    def set_overlap(self, on: bool):
        """solid sub-steps on the second stream (default) or serialised on the context's stream"""
        _ck("mphx_set_overlap", lib.mphx_set_overlap(self._ctx, 1 if on else 0))

    def kernel_timers_ms(self):
        """[bucket rebuild, candidate filter, pass 1, pass 2, solid sub-steps] accumulated device ms"""
        ms = (C.c_double * 5)()
        _ck("mphx_get_kernel_timers", lib.mphx_get_kernel_timers(self._ctx, C.byref(ms)))
        return list(ms)

    @property
    def launch_count(self) -> int:
        return lib.mphx_launch_count(self._ctx)

    @property
    def algorithmic_bytes_per_step(self) -> float:
        return lib.mphx_algorithmic_bytes_per_step(self._ctx)
