"""ctypes mirror of include/mphx.h (declarations only -- no compute, no fallback).

Field order and types must match the C structs exactly; tests/test_abi.py checks sizeof through
`mphx_abi_sizeof` and that every symbol declared in the header is exported by libmphx.so.
"""
from __future__ import annotations

import ctypes as C

TYPE_COUNT = 6  # src/main.cpp:68

MPHX_OK = 0
MPHX_ERR_INVALID = -1
MPHX_ERR_NO_DEVICE = -2
MPHX_ERR_CUDA = -3
MPHX_ERR_IO = -4
MPHX_ERR_NOMEM = -5
MPHX_ERR_UNSUPPORTED = -6
MPHX_ERR_OVERFLOW = -7

MODULE_NONE, MODULE_BAR, MODULE_DAM, MODULE_TUREK_HRON, MODULE_ROLLING1, MODULE_HYDROELASTIC, MODULE_ROLLING2 = 0, 1, 2, 3, 4, 5, 6
WALL_DEFAULT, WALL_ROLLING = 0, 1
COMPAT_DOUBLE_UPDATE = 1

_d = C.c_double
_T = _d * TYPE_COUNT
_V3 = _d * 3


class Params(C.Structure):
    _fields_ = [
        ("dim", C.c_int), ("clamp_module", C.c_int), ("ref_compat", C.c_int), ("wall_module", C.c_int),
        ("time0", _d), ("dt", _d), ("elastic_dt", _d), ("particle_spacing", _d),
        ("domain_min", _V3), ("domain_max", _V3),
        ("radius_ratio_a", _d), ("radius_ratio_p", _d), ("radius_ratio_v", _d),
        ("density", _T), ("bulk_modulus", _T), ("bulk_viscosity", _T), ("shear_viscosity", _T),
        ("surface_tension", _T), ("young_modulus", _T), ("poisson_ratio", _T),
        ("interaction_ratio", _T * TYPE_COUNT),
        ("gravity", _V3),
        ("wall_center", _V3 * TYPE_COUNT), ("wall_velocity", _V3 * TYPE_COUNT),
        ("wall_omega", _V3 * TYPE_COUNT),
    ]

    def copy(self) -> "Params":
        q = Params()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(Params))
        return q


class RunControl(C.Structure):
    _fields_ = [("output_interval", _d), ("vtk_output_interval", _d), ("end_time", _d)]


class Constants(C.Structure):
    _fields_ = [
        ("particle_volume", _d),
        ("radius_a", _d), ("radius_g", _d), ("radius_p", _d), ("radius_v", _d), ("max_radius", _d),
        ("swa", _d), ("swg", _d), ("swp", _d), ("swv", _d), ("r2g", _d), ("n0a", _d), ("n0p", _d),
        ("cof_k", _d), ("cof_a", _T),
        ("wall_rotation", (_V3 * 3) * TYPE_COUNT),
        ("domain_max", _V3), ("domain_width", _V3),
        ("cell_width", _d),
        ("cell_count", C.c_int * 3),
        ("cell_counts", C.c_int),
        ("n0a_count", C.c_int), ("n0p_count", C.c_int),
        ("stencil_range", C.c_int),
    ]


_pi = C.POINTER(C.c_int)
_pd = C.POINTER(C.c_double)


class CuboidC(C.Structure):
    _fields_ = [("type", C.c_int), ("reserved", C.c_int), ("lower", _V3), ("upper", _V3), ("spacing", _d), ("velocity", _V3)]


class HostViews(C.Structure):
    _fields_ = [
        ("property", _pi), ("position", _pd), ("velocity", _pd), ("force", _pd), ("acceleration", _pd),
        ("pressure_p", _pd), ("vol_strain_p", _pd), ("divergence_p", _pd), ("density_a", _pd),
        ("gravity_center", _pd), ("pressure_a", _pd),
        ("neighbor_count", _pi), ("initial_structure_neighbor_count", _pi), ("cell_index", _pi),
        ("normalizer", _pd), ("deform_gradient", _pd), ("strain", _pd), ("stress", _pd),
        ("lambda_lames", _pd), ("mu_lames", _pd),
        ("virial_stress", _pd), ("virial_pressure", _pd),
    ]


# field name -> (trailing shape, is_int)
VIEW_FIELDS = {
    "property": ((), True), "position": ((3,), False), "velocity": ((3,), False),
    "force": ((3,), False), "acceleration": ((3,), False), "pressure_p": ((), False),
    "vol_strain_p": ((), False), "divergence_p": ((), False), "density_a": ((), False),
    "gravity_center": ((3,), False), "pressure_a": ((), False), "neighbor_count": ((), True),
    "initial_structure_neighbor_count": ((), True), "cell_index": ((), True),
    "normalizer": ((3, 3), False), "deform_gradient": ((3, 3), False), "strain": ((3, 3), False),
    "stress": ((3, 3), False), "lambda_lames": ((), False), "mu_lames": ((), False),
    "virial_stress": ((3, 3), False), "virial_pressure": ((), False),
}

# every symbol include/mphx.h declares (tests/test_abi.py parses the header and compares)
EXPORTS = [
    "mphx_version", "mphx_strerror", "mphx_last_error", "mphx_device_count", "mphx_abi_sizeof",
    "mphx_params_default", "mphx_read_data_file", "mphx_read_grid_file", "mphx_free_host",
    "mphx_write_prof_file", "mphx_write_vtk_file", "mphx_write_checkpoint", "mphx_read_checkpoint", "mphx_class_ranges",
    "mphx_compute_constants",
    "mphx_create", "mphx_destroy", "mphx_upload", "mphx_generate_count", "mphx_read_boid_file", "mphx_upload_generated", "mphx_generate_column_histogram", "mphx_multi_upload_generated", "mphx_upload_state", "mphx_init", "mphx_get_constants", "mphx_get_wall_centers",
    "mphx_step", "mphx_step_fluid_only", "mphx_sync", "mphx_time", "mphx_set_time", "mphx_download",
    "mphx_download_owned", "mphx_upload_owned",
    "mphx_debug_neighbors", "mphx_debug_initial_structure_neighbors",
    "mphx_timed_steps", "mphx_set_timing", "mphx_get_timers", "mphx_get_kernel_timers", "mphx_trace_enable", "mphx_trace_read", "mphx_get_virial_ms", "mphx_measure_fp64_peak", "mphx_count_pairs", "mphx_set_overlap", "mphx_join", "mphx_launch_count", "mphx_algorithmic_bytes_per_step",
    "mphx_set_list_reuse", "mphx_get_status",
    "mphx_set_stream", "mphx_partition_columns", "mphx_slab_configure", "mphx_slab_mailbox", "mphx_slab_connect", "mphx_slab_info", "mphx_slab_column_histogram", "mphx_rebalance_cuts", "mphx_slab_recut", "mphx_slab_columns", "mphx_multi_rebalance",
    "mphx_multi_create", "mphx_multi_destroy", "mphx_multi_count", "mphx_multi_context", "mphx_multi_upload", "mphx_multi_init",
    "mphx_multi_step", "mphx_multi_sync", "mphx_multi_time", "mphx_multi_download", "mphx_multi_timed_steps",
]
