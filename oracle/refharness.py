"""ctypes driver for oracle/_ref/libref_<variant>.so -- TEST INFRASTRUCTURE ONLY.

The library is the *unmodified reference* (src/main.cpp) compiled in place by oracle/build_ref.sh
with oracle/ref_harness_tail.cpp appended to the same translation unit.  All reference state is
file-scope `static` (src/main.cpp:83-197), hence ONE case per loaded library: `RefHarness` copies
the .so to a private temp file before dlopen so several cases/variants can coexist in a process.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_VEC = {"Position", "InitialPosition", "Velocity", "Force", "Acceleration", "GravityCenter"}
_TEN = {"Normalizer", "DeformGradient", "Strain", "Stress", "VirialStressAtParticle"}
_SCAL = {"Mass", "DensityA", "PressureA", "VolStrainP", "DivergenceP", "PressureP", "Mu", "Lambda", "VirialPressureAtParticle",
         "Kappa", "LambdaLames", "MuLames"}
_INT = {"Property", "NeighborCount", "InitialStructureNeighborCount"}
_TYPE6 = {"CofA", "Density", "BulkModulus", "BulkViscosity", "ShearViscosity", "SurfaceTension",
          "YoungModulus", "PoissonRatio"}


def variant_name(dim: int, module: str, nb128: bool = False) -> str:
    v = f"{dim}d_{module.lower()}"
    return v + ("_nb128" if nb128 else "")


def available(variant: str) -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"libref_{variant}.so"))


class RefHarness:
    def __init__(self, variant: str, datafile: str, gridfile: str, nthreads: int = 1,
                 logfile: str = "/dev/null", zero_uninitialised: bool = True):
        src = os.path.join(REF_DIR, f"libref_{variant}.so")
        if not os.path.exists(src):
            raise FileNotFoundError(f"{src} missing -- run oracle/build_ref.sh where /root/reference exists")
        fd, self._tmp = tempfile.mkstemp(suffix=f"_libref_{variant}.so")
        os.close(fd)
        shutil.copyfile(src, self._tmp)
        L = self.lib = ctypes.CDLL(self._tmp)
        L.ref_open.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
        L.ref_call.argtypes = [ctypes.c_char_p]
        L.ref_step.argtypes = [ctypes.c_int, ctypes.c_int]
        L.ref_int.argtypes = [ctypes.c_char_p]
        L.ref_double.argtypes = [ctypes.c_char_p]
        L.ref_double.restype = ctypes.c_double
        L.ref_set_double.argtypes = [ctypes.c_char_p, ctypes.c_double]
        L.ref_ptr.argtypes = [ctypes.c_char_p]
        L.ref_ptr.restype = ctypes.c_void_p
        L.ref_write_prof.argtypes = [ctypes.c_char_p]
        L.ref_write_vtk.argtypes = [ctypes.c_char_p]
        self.variant = variant
        L.ref_open(datafile.encode(), gridfile.encode(), logfile.encode(), nthreads)
        if zero_uninitialised:
            L.ref_zero_uninitialised()
        self.n = self.int("ParticleCount")
        self.dim = self.int("dim")
        self.nbmax = self.int("MAX_NEIGHBOR_COUNT")

    def close(self):
        try:
            os.unlink(self._tmp)
        except OSError:
            pass

    # ---- scalars -------------------------------------------------------------------------
    def int(self, name: str) -> int:
        v = self.lib.ref_int(name.encode())
        if v == -999999:
            raise KeyError(name)
        return v

    def double(self, name: str) -> float:
        return self.lib.ref_double(name.encode())

    def set_double(self, name: str, v: float):
        self.lib.ref_set_double(name.encode(), v)

    # ---- arrays (zero-copy views onto the reference's globals) ------------------------------
    def view(self, name: str) -> np.ndarray:
        p = self.lib.ref_ptr(name.encode())
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
        elif name in ("DomainMin", "DomainMax", "DomainWidth", "Gravity"):
            shape, ct = (3,), ctypes.c_double
        elif name in _TYPE6:
            shape, ct = (6,), ctypes.c_double
        elif name == "InteractionRatio":
            shape, ct = (6, 6), ctypes.c_double
        elif name in ("WallCenter", "WallVelocity", "WallOmega"):
            shape, ct = (6, 3), ctypes.c_double
        elif name == "WallRotation":
            shape, ct = (6, 3, 3), ctypes.c_double
        else:
            raise KeyError(name)
        cnt = int(np.prod(shape))
        buf = (ct * cnt).from_address(p)
        return np.ctypeslib.as_array(buf).reshape(shape)

    def get(self, name: str) -> np.ndarray:
        return self.view(name).copy()

    # ---- procedures ---------------------------------------------------------------------
    def init(self):
        self.lib.ref_init()

    def call(self, name: str):
        if self.lib.ref_call(name.encode()) != 0:
            raise KeyError(name)

    def step(self, nsteps: int = 1, stop_after_fluid: bool = False):
        self.lib.ref_step(nsteps, 1 if stop_after_fluid else 0)

    def write_prof(self, fn: str):
        self.lib.ref_write_prof(fn.encode())

    def write_vtk(self, fn: str):
        self.lib.ref_write_vtk(fn.encode())

    def neighbor_sets(self):
        """(count, [sorted neighbour ids]) from NeighborCount/Neighbor (src/main.cpp:1731-1810)."""
        cnt = self.get("NeighborCount")
        nb = self.view("Neighbor")
        return cnt, [np.sort(nb[i, :min(cnt[i], self.nbmax)]) for i in range(self.n)]
