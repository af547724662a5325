"""ctypes driver for oracle/liboracle.so (mph_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Same surface as oracle/refharness.RefHarness so tests can swap the live reference and the
restatement.  Imports only the ABI declarations (struct layouts) from the product package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from particlemethod_fsi_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_VEC = {"Position", "InitialPosition", "Velocity", "Force", "Acceleration", "GravityCenter"}
_TEN = {"Normalizer", "DeformGradient", "Strain", "Stress", "VirialStressAtParticle"}
_SCAL = {"Mass", "DensityA", "PressureA", "VolStrainP", "DivergenceP", "PressureP", "Mu", "Lambda", "VirialPressureAtParticle",
         "Kappa", "LambdaLames", "MuLames"}
_INT = {"Property", "NeighborCount", "InitialStructureNeighborCount"}


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        L.mpho_create.argtypes = [ctypes.POINTER(abi.Params), ctypes.c_int]
        L.mpho_create.restype = ctypes.c_void_p
        L.mpho_destroy.argtypes = [ctypes.c_void_p]
        L.mpho_load.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                ctypes.c_void_p, ctypes.c_void_p]
        L.mpho_init.argtypes = [ctypes.c_void_p]
        L.mpho_step.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.mpho_call.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.mpho_ptr.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.mpho_ptr.restype = ctypes.c_void_p
        L.mpho_int.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.mpho_double.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.mpho_double.restype = ctypes.c_double
        L.mpho_set_time.argtypes = [ctypes.c_void_p, ctypes.c_double]
        _LIB = L
    return _LIB


class Oracle:
    def __init__(self, params: abi.Params, property, position, initial_position, velocity,
                 max_neighbor_count: int = 512):
        L = self.lib = lib()
        self.params = params
        self.ctx = L.mpho_create(ctypes.byref(params), max_neighbor_count)
        t = np.ascontiguousarray(property, dtype=np.int32)
        x = np.ascontiguousarray(position, dtype=np.float64)
        x0 = np.ascontiguousarray(initial_position, dtype=np.float64)
        v = np.ascontiguousarray(velocity, dtype=np.float64)
        rc = L.mpho_load(self.ctx, t.shape[0], t.ctypes.data, x.ctypes.data, x0.ctypes.data, v.ctypes.data)
        if rc:
            raise MemoryError("mpho_load failed")
        self.n = t.shape[0]
        self.dim = params.dim
        self.nbmax = max_neighbor_count

    @classmethod
    def from_case(cls, case, max_neighbor_count: int = 512):
        return cls(case.params, case.property, case.position, case.initial_position, case.velocity,
                   max_neighbor_count)

    def close(self):
        if self.ctx:
            self.lib.mpho_destroy(self.ctx)
            self.ctx = None

    def int(self, name):
        v = self.lib.mpho_int(self.ctx, name.encode())
        if v == -999999:
            raise KeyError(name)
        return v

    def double(self, name):
        return self.lib.mpho_double(self.ctx, name.encode())

    def set_double(self, name, v):
        assert name == "Time"
        self.lib.mpho_set_time(self.ctx, v)

    def view(self, name):
        p = self.lib.mpho_ptr(self.ctx, name.encode())
        if not p:
            raise KeyError(name)
        n = self.n
        if name in _VEC:
            shape, ct = (n, 3), ctypes.c_double
        elif name in _TEN:
            shape, ct = (n, 3, 3), ctypes.c_double
        elif name in _SCAL:
            shape, ct = (n,), ctypes.c_double
        elif name in _INT:
            shape, ct = (n,), ctypes.c_int
        elif name in ("Neighbor", "InitialStructureNeighbor"):
            shape, ct = (n, self.nbmax), ctypes.c_int
        elif name in ("CellIndex", "CellParticle"):
            shape, ct = (self.int("PowerParticleCount"),), ctypes.c_int
        elif name in ("CellParticleBegin", "CellParticleEnd"):
            shape, ct = (self.int("CellCounts"),), ctypes.c_int
        elif name in ("DomainMin", "DomainMax", "DomainWidth"):
            shape, ct = (3,), ctypes.c_double
        elif name == "CofA":
            shape, ct = (6,), ctypes.c_double
        elif name == "WallCenter":
            shape, ct = (6, 3), ctypes.c_double
        elif name == "WallRotation":
            shape, ct = (6, 3, 3), ctypes.c_double
        else:
            raise KeyError(name)
        cnt = int(np.prod(shape))
        buf = (ct * cnt).from_address(p)
        return np.ctypeslib.as_array(buf).reshape(shape)

    def get(self, name):
        return self.view(name).copy()

    def init(self):
        rc = self.lib.mpho_init(self.ctx)
        if rc:
            raise RuntimeError(f"mpho_init -> {rc}")

    def call(self, name):
        if self.lib.mpho_call(self.ctx, name.encode()) != 0:
            raise KeyError(name)

    def step(self, nsteps=1, stop_after_fluid=False):
        self.lib.mpho_step(self.ctx, nsteps, 1 if stop_after_fluid else 0)

    def neighbor_sets(self):
        cnt = self.get("NeighborCount")
        nb = self.view("Neighbor")
        return cnt, [np.sort(nb[i, :min(cnt[i], self.nbmax)]) for i in range(self.n)]

    def cell_of_particle(self):
        """CellId per particle (original order) from the sorted (CellIndex, CellParticle) pairs"""
        ci = self.view("CellIndex")[: self.n]
        cp = self.view("CellParticle")[: self.n]
        out = np.empty(self.n, dtype=np.int32)
        out[cp] = ci
        return out
