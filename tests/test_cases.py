"""The input builder restates the reference pre-processor's lattice rule; pinned on results/Dam."""
import hashlib
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT
from particlemethod_fsi_b200 import cases, solver

# md5 of /root/reference/results/Dam/dam.grid (also reproduced by the reference generator)
DAM_GRID_MD5 = "3a7893d123c75de04197480fa161eca4"


def test_dam2d_reproduces_shipped_grid_bytes(tmp_path):
    c = cases.dam2d()
    assert c.n == 6650 and c.counts() == (4850, 0, 1800)
    fn = str(tmp_path / "dam.grid")
    cases.write_grid_file(fn, c)
    md5 = hashlib.md5(open(fn, "rb").read()).hexdigest()
    assert md5 == DAM_GRID_MD5, md5
    ref = "/root/reference/results/Dam/dam.grid"
    if os.path.exists(ref):
        assert open(fn, "rb").read() == open(ref, "rb").read()


def test_grid_round_trip_through_c_reader(tmp_path):
    c = cases.tiny3d()
    fn = str(tmp_path / "c.grid")
    cases.write_grid_file(fn, c)
    p = c.params.copy()
    t, x, x0, v = solver.read_grid_file(fn, p)
    assert np.array_equal(t, c.property) and np.array_equal(x, c.position)
    assert np.array_equal(x0, c.initial_position) and np.array_equal(v, c.velocity)
    assert list(p.domain_min) == list(c.params.domain_min) and list(p.domain_max) == list(c.params.domain_max)
    assert p.particle_spacing == c.params.particle_spacing


def test_classes_contiguous_and_sizes():
    for mk, n in ((cases.tiny2d, 684), (cases.tiny3d, 2544), (cases.fsi2d, 23338)):
        c = mk()
        assert c.n == n
        r = solver.class_ranges(c.property)
        f, s, w = c.counts()
        assert r[1] - r[0] == f and r[3] - r[2] == s and r[5] - r[4] == w


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "GeneratorForMph")),
                    reason="reference generator not built")
def test_lattice_rule_matches_reference_generator(tmp_path):
    c = cases.tiny3d()
    cases.write_boid_file(str(tmp_path / "t.boid"), c)   # (the same writer feeds mphx_read_boid_file: tests/test_io.py)
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "GeneratorForMph"), "t"], cwd=tmp_path, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    mine = tmp_path / "mine.grid"
    cases.write_grid_file(str(mine), c)
    assert open(mine, "rb").read() == open(tmp_path / "t.grid", "rb").read()


def test_generator_count_matches_the_lattice_rule():
    """mphx_generate_count (host part of the device-side generator) counts what cases.py builds"""
    from particlemethod_fsi_b200 import solver
    for c in (cases.dam2d(), cases.fsi3d_mini(), cases.tiny2d(), cases.bar2d()):
        assert solver.generate_count(c.cuboids) == c.n, c.name
    assert solver.generate_count([cases.Cuboid(9, (0, 0, 0), (1, 1, 1), 0.1)]) == -1     # invalid type


def test_generated_column_histogram_equals_the_particle_histogram():
    """mphx_generate_column_histogram (axis tables only) against the histogram of the generated particles (slab.plan)"""
    import ctypes as C
    from particlemethod_fsi_b200 import slab, solver
    for case in (cases.fsi3d_mini(), cases.fsi2d(), cases.dam2d()):
        k = solver.compute_constants(case.params)
        ncols = k.cell_count[0]
        want = slab.plan(case, 2, k)["hist"] if ncols >= 4 * k.stencil_range else None
        arr = solver._cuboid_array(case.cuboids)
        hist = np.zeros(ncols, dtype=np.int64)
        rc = solver.lib.mphx_generate_column_histogram(C.cast(arr, C.c_void_p), len(case.cuboids), case.params.domain_min[0], k.cell_width, ncols,
                                                       hist.ctypes.data)
        assert rc == 0
        t = case.property
        assert hist.sum() == int(((t < 2) | (t >= 4)).sum())
        if want is not None:
            assert np.array_equal(hist, want)
