"""Host constants of the product (mphx_compute_constants) are bit-identical to the reference's
initializeWeight/Fluid/Wall/Domain (via golden scalars and the pinned oracle)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle.oracle import Oracle
from particlemethod_fsi_b200 import cases, solver

PAIRS = dict(n0a="N0a", n0p="N0p", swa="Swa", swg="Swg", swp="Swp", swv="Swv", r2g="R2g", max_radius="MaxRadius",
             radius_a="RadiusA", radius_p="RadiusP", radius_v="RadiusV", particle_volume="ParticleVolume", cof_k="CofK")


@pytest.mark.parametrize("name", ["tiny2d", "tiny3d"])
def test_constants_equal_reference_golden(name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    k = solver.compute_constants(getattr(cases, name)().params)
    for mine, ref in PAIRS.items():
        assert getattr(k, mine) == float(g["const_" + ref]), mine
    assert np.array_equal(np.array(k.cof_a), g["const_CofA"])
    assert np.array_equal(np.array(k.wall_rotation).reshape(6, 3, 3)[4:], g["const_WallRotation"][4:])


@pytest.mark.parametrize("mk", [cases.dam2d, cases.bar2d, cases.fsi2d, cases.fsi3d_mini])
def test_constants_equal_oracle(mk):
    c = mk()
    c.params.wall_omega[4][2] = 3.0      # exercise the quaternion path (Q9: theta = |omega|^2)
    c.params.wall_velocity[5][0] = 0.25
    c.params.surface_tension[1] = 0.072
    k = solver.compute_constants(c.params)
    o = Oracle.from_case(c)
    o.init()
    for mine, ref in PAIRS.items():
        assert getattr(k, mine) == o.double(ref), mine
    assert np.array_equal(np.array(k.cof_a), o.get("CofA"))
    assert np.array_equal(np.array(k.wall_rotation).reshape(6, 3, 3)[4:], o.get("WallRotation")[4:])
    assert list(k.cell_count) == [o.int("CellCount0"), o.int("CellCount1"), o.int("CellCount2")]
    assert k.cell_counts == o.int("CellCounts")
    assert (k.n0a_count, k.n0p_count) == (o.int("n0a_count"), o.int("n0p_count"))
    o.close()
