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


@pytest.mark.skipif(not REF_PRESENT, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("module", ["turek", "rolling1", "hydro", "rolling2", "rollwall"])
def test_oracle_other_modules_vs_live_reference(module, tmp_path):
    """the reference's other compile-time variants (Turek_Hron, Rolling1, Hydroelastic, Rolling2 clamps of
    updateElasticPosition :1910-2082; the rolling wall of `#define Rolling` :2958-3031), each compiled from the untouched
    source by oracle/build_ref.sh, against the oracle's restatement: bit for bit over 5 steps."""
    from oracle.refharness import RefHarness
    c = cases.module_case(module)
    cases.write_grid_file(str(tmp_path / "c.grid"), c)
    cases.write_data_file(str(tmp_path / "c.data"), c.params, c.rc)
    h = RefHarness("2d_" + module, str(tmp_path / "c.data"), str(tmp_path / "c.grid"), nthreads=2)
    h.init()
    o = Oracle.from_case(c)
    o.init()
    x0 = c.initial_position
    solid = (c.property >= 2) & (c.property < 4)
    if module != "rollwall":      # the clamp really splits the plate
        from particlemethod_fsi_b200 import abi
        cl = {abi.MODULE_TUREK_HRON: x0[:, 0] < 0.205, abi.MODULE_ROLLING1: x0[:, 1] < 0.003,
              abi.MODULE_HYDROELASTIC: (x0[:, 0] < 0.01) | (x0[:, 0] > 1.99), abi.MODULE_ROLLING2: x0[:, 1] > 0.3420}[c.params.clamp_module]
        assert 0 < int((cl & solid).sum()) < int(solid.sum())
    for _ in range(5):
        h.step(1)
        o.step(1)
        for f in ["Position", "Velocity", "Force", "PressureP", "Stress", "NeighborCount"]:
            assert np.array_equal(h.get(f), o.get(f)), (module, f)
    if module == "rollwall":
        w = c.property >= 4
        assert np.abs(o.get("Position")[w] - c.position[w]).max() > 0   # the walls did move
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


def _through_the_file(c, tmp_path):
    """write the case as the reference reads it (%e: 7 digits) and take the parsed values back, so that the
    oracle starts from exactly what the reference parsed"""
    cases.write_grid_file(str(tmp_path / "c.grid"), c)
    cases.write_data_file(str(tmp_path / "c.data"), c.params, c.rc)
    _t, _h, t, x, x0, v = cases.read_grid_file(str(tmp_path / "c.grid"))
    return cases.Case(c.name, c.params, c.rc, t, x, x0, v)


def _jitter(c, amp, seed=12345):
    rng = np.random.default_rng(seed)
    fl = c.property < 2
    d = rng.uniform(-amp, amp, size=(int(fl.sum()), 3)) * c.params.particle_spacing
    if c.params.dim == 2:
        d[:, 2] = 0.0
    c.position[fl] += d
    return c


def _variants():
    a = _jitter(cases.tiny2d(), 0.8)                     # several particles per bucket, empty buckets
    b = _jitter(cases.tiny3d(), 0.8)
    st = _jitter(cases.tiny2d(), 0.3)                    # surface tension + asymmetric wetting (zero in all shipped data)
    st.params.surface_tension[0] = st.params.surface_tension[1] = 0.072
    st.params.interaction_ratio[1][4] = 0.6
    mw = cases.tiny2d()                                  # moving, rotating wall (Wall6 line of the .data file)
    mw.params.wall_velocity[4][0] = 0.5
    mw.params.wall_omega[4][2] = 2.0
    mw.params.wall_center[4][0] = 0.05
    return [("jitter2d", a, "2d_dam"), ("jitter3d", b, "3d_dam"), ("surface_tension", st, "2d_dam"), ("moving_wall", mw, "2d_dam")]


@pytest.mark.skipif(not REF_PRESENT, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("which", [0, 1, 2, 3])
def test_oracle_vs_live_reference_off_the_lattice(which, tmp_path):
    """the regimes the GPU parity tests rely on beyond the lattice-born cases: irregular bucket occupancy,
    surface tension switched on, a moving wall -- oracle and reference library stage by stage, bit for bit"""
    from oracle.refharness import RefHarness
    name, c, variant = _variants()[which]
    c = _through_the_file(c, tmp_path)
    h = RefHarness(variant, str(tmp_path / "c.data"), str(tmp_path / "c.grid"), nthreads=2)
    h.init()
    o = Oracle.from_case(c)
    o.init()
    for f in ("NeighborCount", "VolStrainP", "DensityA"):
        assert np.array_equal(h.get(f), o.get(f)), (name, "init", f)
    if which < 2:
        assert np.bincount(o.cell_of_particle()).max() >= 2
    stages = ["calculateWall", "calculatePeriodicBoundary", "resetForce", "resetAccel", "calculateNeighbor",
              "calculateDensityA", "calculateGravityCenter", "calculateDensityP", "calculateDivergenceP",
              "calculatePhysicalCoefficients", "calculatePressureP", "calculatePressureA",
              "calculateDiffuseInterface", "calculateViscosityV", "calculateGravity", "calculateInterfaceForce",
              "calculateAcceleration", "calculateConvection", "calculateElasticDeformationVector",
              "calculateStress", "calculateStressForce", "updateElasticPosition"]
    watch = ["Position", "Velocity", "Force", "Acceleration", "PressureP", "VolStrainP", "DivergenceP", "DensityA",
             "GravityCenter", "PressureA", "NeighborCount", "Neighbor", "DeformGradient", "Strain", "Stress"]
    for step in range(3):
        for st in stages:
            h.call(st)
            o.call(st)
            for f in watch:
                assert np.array_equal(h.get(f), o.get(f)), (name, step, st, f)
    if which == 2:
        assert np.abs(o.get("PressureA")).max() > 0
    h.close()
    o.close()


@pytest.mark.skipif(not REF_PRESENT, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("which", [0, 1, 2])
def test_oracle_virial_stress_vs_live_reference(which, tmp_path):
    """calculateVirialStressAtParticle (src/main.cpp:3077-3318): the per-output-step
    diagnostic, restated in the oracle ahead of its CUDA kernel -- bit for bit against the reference
    library after a few steps, incl. the surface-tension terms"""
    from oracle.refharness import RefHarness
    name, c, variant = _variants()[which]
    c = _through_the_file(c, tmp_path)
    h = RefHarness(variant, str(tmp_path / "c.data"), str(tmp_path / "c.grid"), nthreads=2)
    h.init()
    o = Oracle.from_case(c)
    o.init()
    h.step(3)
    o.step(3)
    h.call("calculateVirialStressAtParticle")
    o.call("calculateVirialStressAtParticle")
    for f in ("VirialStressAtParticle", "VirialPressureAtParticle"):
        a, b = h.get(f), o.get(f)
        assert np.array_equal(a, b), (name, f, float(np.abs(a - b).max()))
    assert np.abs(o.get("VirialStressAtParticle")).max() > 0
    h.close()
    o.close()
