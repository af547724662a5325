"""Pin the oracle (oracle/mph_oracle.c): bit-exact against golden vectors dumped from the reference
itself (tests/golden, made by oracle/make_golden.py) and, where oracle/_ref is present, against the
live reference library stage by stage."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, REF_PRESENT
from oracle.oracle import Oracle
from particlemethod_fsi_b200 import cases

FIELDS = ["Position", "Velocity", "Force", "Acceleration", "PressureP", "VolStrainP", "DivergenceP",
          "NeighborCount", "InitialStructureNeighborCount", "Normalizer", "DeformGradient", "Strain",
          "Stress", "LambdaLames", "MuLames"]
SCALARS = ["N0a", "N0p", "Swa", "Swg", "Swp", "Swv", "R2g", "MaxRadius", "RadiusA", "RadiusP", "RadiusV",
           "ParticleVolume", "CofK"]


def _neighbor_sha(o):
    cnt = o.get("NeighborCount")
    nb = o.view("Neighbor")
    rows = [np.sort(nb[i, :cnt[i]]) for i in range(o.n)]
    off = np.zeros(o.n + 1, dtype=np.int64)
    off[1:] = np.cumsum(cnt)
    ids = np.concatenate(rows).astype(np.int32)
    return hashlib.sha256(off.tobytes() + ids.tobytes()).hexdigest()


@pytest.mark.parametrize("name,steps", [("tiny2d", [0, 1, 20]), ("tiny3d", [0, 1, 10])])
def test_oracle_bit_exact_vs_reference_golden(name, steps):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    c = getattr(cases, name)()
    assert c.n == int(g["n"])
    o = Oracle.from_case(c)
    o.init()
    for s in SCALARS:
        assert o.double(s) == float(g["const_" + s]), s
    assert np.array_equal(o.get("CofA"), g["const_CofA"])
    assert np.array_equal(o.get("WallRotation"), g["const_WallRotation"])
    done = 0
    for target in steps:
        if target > done:
            o.step(target - done)
            done = target
        for f in FIELDS:
            assert np.array_equal(o.get(f), g[f"s{target}_{f}"]), (target, f)
        assert np.array_equal(o.cell_of_particle(), g[f"s{target}_CellIndex"])
        assert _neighbor_sha(o) == str(g[f"s{target}_NeighborSetsSha"])
        assert o.double("Time") == float(g[f"s{target}_Time"])
    o.close()


def test_oracle_dam2d_100_steps_vs_reference_golden():
    g = np.load(os.path.join(GOLDEN, "dam2d.npz"))
    o = Oracle.from_case(cases.dam2d())
    o.init()
    o.step(100)
    for f in ["Position", "Velocity", "PressureP", "NeighborCount"]:
        assert np.array_equal(o.get(f), g[f"s100_{f}"]), f
    assert np.array_equal(o.cell_of_particle(), g["s100_CellIndex"])
    o.close()


@pytest.mark.skipif(not REF_PRESENT, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name,variant,steps", [("tiny2d", "2d_dam", 5), ("tiny3d", "3d_dam", 3),
                                                ("bar2d", "2d_bar", 3)])
def test_oracle_stage_by_stage_vs_live_reference(name, variant, steps, tmp_path):
    from oracle.refharness import RefHarness
    c = getattr(cases, name)()
    cases.write_grid_file(str(tmp_path / "c.grid"), c)
    cases.write_data_file(str(tmp_path / "c.data"), c.params, c.rc)
    h = RefHarness(variant, str(tmp_path / "c.data"), str(tmp_path / "c.grid"), nthreads=2)
    h.init()
    o = Oracle.from_case(c)
    o.init()
    stages = ["calculateWall", "calculatePeriodicBoundary", "resetForce", "resetAccel", "calculateNeighbor",
              "calculateDensityA", "calculateGravityCenter", "calculateDensityP", "calculateDivergenceP",
              "calculatePhysicalCoefficients", "calculatePressureP", "calculatePressureA",
              "calculateDiffuseInterface", "calculateViscosityV", "calculateGravity", "calculateInterfaceForce",
              "calculateAcceleration", "calculateConvection", "calculateElasticDeformationVector",
              "calculateStress", "calculateStressForce", "updateElasticPosition"]
    watch = ["Position", "Velocity", "Force", "Acceleration", "PressureP", "VolStrainP", "DivergenceP", "DensityA",
             "GravityCenter", "PressureA", "NeighborCount", "Neighbor", "DeformGradient", "Strain", "Stress"]
    for _ in range(steps):
        for st in stages:
            h.call(st)
            o.call(st)
            for f in watch:
                assert np.array_equal(h.get(f), o.get(f)), (st, f)
    h.close()
    o.close()


def test_oracle_surface_tension_path_runs_and_is_symmetric_free():
    """surface tension on: PressureA / diffuse-interface forces are exercised (all shipped data has 0)"""
    c = cases.tiny2d()
    c.params.surface_tension[0] = c.params.surface_tension[1] = 0.072
    o = Oracle.from_case(c)
    o.init()
    o.step(3)
    assert np.isfinite(o.get("Force")).all()
    assert np.abs(o.get("PressureA")).max() > 0
    o.close()
