"""File formats at the drop-in boundary: the C readers/writers against text produced by the
reference executable itself (tests/golden/*.gz, made by oracle/make_golden.py)."""
import gzip
import os

import numpy as np
import pytest

from conftest import GOLDEN
from particlemethod_fsi_b200 import abi, cases, solver


def _gold(name):
    return gzip.open(os.path.join(GOLDEN, name + ".gz"), "rb").read()


def test_data_file_reader_round_trip(tmp_path):
    c = cases.tiny3d()
    fn = str(tmp_path / "c.data")
    cases.write_data_file(fn, c.params, c.rc)
    p, rc, bad = solver.read_data_file(fn, dim=3, module=abi.MODULE_DAM)
    assert bad == []
    for k in ("dt", "elastic_dt", "radius_ratio_a", "radius_ratio_p", "radius_ratio_v"):
        assert getattr(p, k) == getattr(c.params, k), k
    for k in ("density", "bulk_modulus", "bulk_viscosity", "shear_viscosity", "surface_tension", "young_modulus",
              "poisson_ratio", "gravity"):
        assert list(getattr(p, k)) == list(getattr(c.params, k)), k
    assert [list(r) for r in p.interaction_ratio] == [list(r) for r in c.params.interaction_ratio]
    assert (rc.end_time, rc.output_interval, rc.vtk_output_interval) == \
        (c.rc.end_time, c.rc.output_interval, c.rc.vtk_output_interval)


def test_data_file_reader_reports_unknown_lines_like_reference(tmp_path):
    fn = tmp_path / "d.data"
    fn.write_text("#######\nDt 1.0e-4\nCohesion 1 0 0 0\nWall2 Center 0 0 0 Velocity 0 0 0 Omega 0 0 0\n"
                  "SurfaceTension 0.1 0.2 0.3 0.4\nYoungModulus 1 2 3 4\n\nGravity 0 -9.8 0\n"
                  "Wall6  Center 1 2 3 Velocity 4 5 6 Omega 7 8 9\n")
    p, rc, bad = solver.read_data_file(str(fn))
    assert p.dt == 1.0e-4 and list(p.gravity) == [0.0, -9.8, 0.0]
    # src/main.cpp:756-757: four values go to types (0,1,4,5) and (2,3,4,5)
    assert list(p.surface_tension) == [0.1, 0.2, 0.0, 0.0, 0.3, 0.4]
    assert list(p.young_modulus) == [0.0, 0.0, 1.0, 2.0, 3.0, 4.0]
    assert list(p.wall_center[4]) == [1.0, 2.0, 3.0] and list(p.wall_omega[4]) == [7.0, 8.0, 9.0]
    assert [b.split()[0] for b in bad if b.strip()] == ["Cohesion", "Wall2"]
    assert sum(1 for b in bad if not b.strip()) == 1  # the blank line is "invalid" too (:768-770)


def _parse_prof(txt):
    lines = txt.decode().splitlines()
    time = float(lines[0])
    n = int(lines[1].split()[0])
    data = np.array([[float(v) for v in ln.split()] for ln in lines[2:2 + n]])
    return time, lines[1], data


def test_prof_writer_is_byte_identical_to_reference_text(tmp_path):
    for name in ("tiny2d", "tiny3d"):
        c = getattr(cases, name)()
        for tag in ("t000.prof", "t002.prof"):
            gold = _gold(f"{name}_{tag}")
            time, _hdr, data = _parse_prof(gold)
            # re-emit through our writer from the 7-digit values: must reproduce the bytes when the
            # values survive the %e round trip (they do: they were printed with %e)
            fn = str(tmp_path / "o.prof")
            solver.write_prof_file(fn, time, c.params, data[:, 0].astype(np.int32), data[:, 1:4], data[:, 4:7], data[:, 7:10])
            assert open(fn, "rb").read() == gold, (name, tag)
        # and the step-0 file is exactly the input state
        _t, _h, data = _parse_prof(_gold(f"{name}_t000.prof"))
        assert np.array_equal(data[:, 1:4], c.position) and np.array_equal(data[:, 0].astype(np.int32), c.property)


def test_vtk_writer_is_byte_identical_to_reference_text(tmp_path):
    """feed the oracle's state (bit-identical to the reference, see test_oracle.py) to our writer and
    compare with the .vtk the reference executable wrote"""
    from oracle.oracle import Oracle
    for name in ("tiny2d", "tiny3d"):
        c = getattr(cases, name)()
        o = Oracle.from_case(c)
        o.init()

        def emit(fn):
            f = dict(property=o.get("Property"), position=o.get("Position"), velocity=o.get("Velocity"),
                     force=o.get("Force"), acceleration=o.get("Acceleration"), stress=o.get("Stress"),
                     strain=o.get("Strain"), neighbor_count=o.get("NeighborCount"),
                     initial_structure_neighbor_count=o.get("InitialStructureNeighborCount"))
            solver.write_vtk_file(fn, o.get("InitialPosition"), f)
            return open(fn, "rb").read()

        assert emit(str(tmp_path / "o.vtk")) == _gold(f"{name}_output.vtk")
        o.step(1)   # t000.vtk is written after step 0 (quirk Q8)
        assert emit(str(tmp_path / "a.vtk")) == _gold(f"{name}_t000.vtk")
        o.step(3)   # t003.vtk after the 4th step
        assert emit(str(tmp_path / "b.vtk")) == _gold(f"{name}_t003.vtk")
        o.close()


def test_checkpoint_round_trip_is_lossless(tmp_path):
    """the binary checkpoint keeps every bit (the .prof text keeps 7 digits) and the wall centres"""
    from particlemethod_fsi_b200 import abi, cases, solver
    c = cases.tiny3d()
    rng = np.random.default_rng(12345)
    x = c.position + rng.uniform(-1e-4, 1e-4, c.position.shape) * np.pi      # not representable in 7 digits
    v = rng.standard_normal(c.velocity.shape)
    p = c.params.copy()
    p.wall_center[4][0], p.wall_center[5][2] = 0.123456789012345, -7.0 / 3.0
    fn = str(tmp_path / "s.ckp")
    solver.write_checkpoint(fn, 0.0123456789, p, c.property, x, c.initial_position, v)
    q = abi.Params()
    t, x2, x02, v2 = solver.read_checkpoint(fn, q)
    assert np.array_equal(t, c.property) and np.array_equal(x2, x) and np.array_equal(x02, c.initial_position) and np.array_equal(v2, v)
    assert q.time0 == 0.0123456789 and q.dim == 3 and q.particle_spacing == p.particle_spacing
    assert list(q.domain_min) == list(p.domain_min) and list(q.domain_max) == list(p.domain_max)
    assert q.wall_center[4][0] == 0.123456789012345 and q.wall_center[5][2] == -7.0 / 3.0
    # the text format of the reference loses the state (this is why the checkpoint exists)
    solver.write_prof_file(str(tmp_path / "s.prof"), 0.0, p, c.property, x, c.initial_position, v)
    _time, _hdr, _t, xp, _x0, _v = cases.read_grid_file(str(tmp_path / "s.prof"))
    assert not np.array_equal(xp, x)
    with open(str(tmp_path / "bad.ckp"), "wb") as f:
        f.write(b"not a checkpoint")
    with pytest.raises(solver.MphxError):
        solver.read_checkpoint(str(tmp_path / "bad.ckp"), q)


def test_boid_reader_round_trip_and_errors(tmp_path):
    """mphx_read_boid_file: the pre-processor's input (header + StartCuboid blocks) -> parameters and cuboids; the cuboids
    generate exactly the case's particle count; unsupported shapes and incomplete files are errors"""
    from particlemethod_fsi_b200 import cases, solver
    for case in (cases.dam2d(), cases.fsi3d_mini(), cases.bar2d()):
        fn = str(tmp_path / (case.name + ".boid"))
        cases.write_boid_file(fn, case)
        p = case.params.copy()
        p.particle_spacing = -1.0
        p.domain_min[0] = 123.0
        cubs = solver.read_boid_file(fn, p)
        assert len(cubs) == len(case.cuboids)
        for a, b in zip(cubs, case.cuboids):
            assert a.type == b.type and a.spacing == b.spacing
            assert tuple(a.lower) == tuple(b.lower) and tuple(a.upper) == tuple(b.upper) and tuple(a.velocity) == tuple(b.velocity)
        # header values come back through the `%e` text of the generator's .grid header (here: exactly representable in 7 digits)
        assert p.particle_spacing == float("%e" % case.params.particle_spacing)
        assert list(p.domain_min) == [float("%e" % v) for v in case.params.domain_min]
        assert list(p.domain_max) == [float("%e" % v) for v in case.params.domain_max]
        assert p.time0 == 0.0
        assert solver.generate_count(cubs) == case.n
    bad = tmp_path / "bad.boid"
    bad.write_text("ParticleDistance 0.001\nLowerDomain 0 0 0\nUpperDomain 1 1 1\nStartCuboid2\nEndCuboid2\n")
    with pytest.raises(Exception):
        solver.read_boid_file(str(bad), cases.dam2d().params.copy())
    bad.write_text("ParticleDistance 0.001\nLowerDomain 0 0 0\nUpperDomain 1 1 1\nStartCuboid\n Spacing 0.001\n Type 1\nEndCuboid\n")
    with pytest.raises(Exception):
        solver.read_boid_file(str(bad), cases.dam2d().params.copy())
    with pytest.raises(Exception):
        solver.read_boid_file(str(tmp_path / "missing.boid"), cases.dam2d().params.copy())
