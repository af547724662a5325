"""The drop-in executable (same CLI as the reference, src/main.cpp:501-508) end to end on the GPU:
its .prof/.vtk files against the text the reference executable wrote for the same inputs."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from particlemethod_fsi_b200 import cases

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "particlemethod_fsi_b200", "Mph_Elastic_Explicit")


def _gold(name):
    return gzip.open(os.path.join(GOLDEN, name + ".gz"), "rb").read().decode()


def _numbers(txt):
    out = []
    for tok in txt.split():
        try:
            out.append(float(tok))
        except ValueError:
            out.append(float(abs(hash(tok)) % 1000))
    return np.array(out)


@pytest.mark.parametrize("sync_io", ["0", "1"])   # background writer thread (default) / writes in the main thread
@pytest.mark.parametrize("name,dim,module", [("tiny2d", "2", "dam"), ("tiny3d", "3", "dam")])
def test_driver_outputs_match_reference_cli(name, dim, module, sync_io, tmp_path):
    c = getattr(cases, name)()
    c.rc.end_time = 3.5 * c.params.dt
    c.rc.output_interval = 2.0 * c.params.dt
    c.rc.vtk_output_interval = 3.0 * c.params.dt
    cases.write_grid_file(str(tmp_path / "c.grid"), c)
    cases.write_data_file(str(tmp_path / "t.data"), c.params, c.rc)
    r = subprocess.run([EXE, "t.data", "c.grid", "t%03d.prof", "t%03d.vtk", "t.log", "4", dim, module],
                       cwd=tmp_path, capture_output=True, text=True, env=dict(os.environ, MPHX_SYNC_IO=sync_io), timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    produced = sorted(f for f in os.listdir(tmp_path) if f.endswith((".prof", ".vtk")))
    assert produced == ["output.vtk", "t000.prof", "t000.vtk", "t002.prof", "t003.vtk"]   # quirk Q8 naming
    # state before the first step: byte-identical
    assert (tmp_path / "t000.prof").read_text() == _gold(f"{name}_t000.prof")
    assert (tmp_path / "output.vtk").read_text() == _gold(f"{name}_output.vtk")
    for fn in ("t002.prof", "t000.vtk", "t003.vtk"):
        mine, gold = (tmp_path / fn).read_text(), _gold(f"{name}_{fn}")
        ml, gl = mine.splitlines(), gold.splitlines()
        assert len(ml) == len(gl), fn
        # identical layout: every non-numeric line (headers, section names) equal
        for a, b in zip(ml, gl):
            if a != b:
                assert a[:1].isdigit() or a[:1] == "-", (fn, a, b)
        a, b = _numbers(mine), _numbers(gold)
        assert a.shape == b.shape
        # %e prints 7 digits (and the vtk casts to float): values agree to print precision
        assert np.abs(a - b).max() <= 2e-6 * max(1.0, np.abs(b).max()), fn
        if fn.endswith(".prof"):
            # type + Position + InitialPosition columns print identically (velocities of a fluid that
            # starts at rest are rounding noise at these early steps and are covered by the bound above)
            same = sum(x.split()[:7] == y.split()[:7] for x, y in zip(ml[2:], gl[2:])) / (len(gl) - 2)
            assert same > 0.995, (fn, same)
    log = (tmp_path / "t.log").read_text()
    for key in ("neighbor search:", "explicit calculation:", "virial calculation:", "other calculation:", "total:",
                "total (check):", "N0p ="):
        assert key in log


def test_driver_on_several_slabs_writes_the_same_files(tmp_path):
    """MPHX_NGPU: the same C++ main drives N slab contexts through mphx_multi_* (here two slabs sharing the one GPU of
    the test box); the ring reproduces the single context bit for bit, so every output file is byte-identical.
    Also: the virial sections behind MPHX_VTK_VIRIAL and the lossless checkpoint next to every .prof (MPHX_CHECKPOINT)."""
    c = cases.tiny3d()
    c.rc.end_time = 3.5 * c.params.dt
    c.rc.output_interval = 2.0 * c.params.dt
    c.rc.vtk_output_interval = 3.0 * c.params.dt
    outs = {}
    for tag, env in (("one", {}), ("two", dict(MPHX_NGPU="2", MPHX_DEVICES="0,0", CUDA_DEVICE_MAX_CONNECTIONS="32"))):
        d = tmp_path / tag
        d.mkdir()
        cases.write_grid_file(str(d / "c.grid"), c)
        cases.write_data_file(str(d / "t.data"), c.params, c.rc)
        r = subprocess.run([EXE, "t.data", "c.grid", "t%03d.prof", "t%03d.vtk", "t.log", "4", "3", "dam"], cwd=d, capture_output=True,
                           text=True, env=dict(os.environ, **env), timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[tag] = {f: (d / f).read_bytes() for f in sorted(os.listdir(d)) if f.endswith((".prof", ".vtk"))}
    assert list(outs["one"]) == list(outs["two"]) and len(outs["one"]) == 5
    for f in outs["one"]:
        assert outs["one"][f] == outs["two"][f], f
    d = tmp_path / "virial"
    d.mkdir()
    cases.write_grid_file(str(d / "c.grid"), c)
    cases.write_data_file(str(d / "t.data"), c.params, c.rc)
    r = subprocess.run([EXE, "t.data", "c.grid", "t%03d.prof", "t%03d.vtk", "t.log", "4", "3", "dam"], cwd=d, capture_output=True, text=True,
                       env=dict(os.environ, MPHX_VTK_VIRIAL="1", MPHX_CHECKPOINT="t%03d.ckp"), timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    vtk = (d / "t003.vtk").read_text()
    assert "SCALARS VirialPressureAtParticle float 1" in vtk and "SCALARS VirialStressAtParticle[1][0] float 1" in vtk
    assert vtk.replace("\r", "").count("\n") > outs["one"]["t003.vtk"].count(b"\n")
    assert sorted(f for f in os.listdir(d) if f.endswith(".ckp")) == ["t000.ckp", "t002.ckp"]
    # restart from the checkpoint of step 2: the run continues (Time and step index come from the file)
    r = subprocess.run([EXE, "t.data", "t002.ckp", "r%03d.prof", "r%03d.vtk", "r.log", "4", "3", "dam"], cwd=d, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert (d / "r003.vtk").exists()


def test_driver_rebalances_the_slabs_in_place(tmp_path):
    """MPHX_REBALANCE_EVERY: the driver re-cuts the slabs of a multi-GPU run while it steps (mphx_multi_rebalance); with the
    reference's rebuild-every-step schedule the files stay byte-identical to the single-context run"""
    c = cases.dam2d()
    c.velocity[c.property < 2, 0] = 1.5
    c.rc.end_time = 40.5 * c.params.dt
    c.rc.output_interval = 40.0 * c.params.dt
    c.rc.vtk_output_interval = 100.0 * c.params.dt
    outs, logs = {}, {}
    for tag, env in (("one", dict(MPHX_LIST_REUSE="0")),
                     ("four", dict(MPHX_LIST_REUSE="0", MPHX_NGPU="4", MPHX_DEVICES="0,0,0,0", MPHX_REBALANCE_EVERY="10",
                                   CUDA_DEVICE_MAX_CONNECTIONS="32"))):
        d = tmp_path / tag
        d.mkdir()
        cases.write_grid_file(str(d / "c.grid"), c)
        cases.write_data_file(str(d / "t.data"), c.params, c.rc)
        r = subprocess.run([EXE, "t.data", "c.grid", "t%03d.prof", "t%03d.vtk", "t.log", "4", "2", "dam"], cwd=d, capture_output=True,
                           text=True, env=dict(os.environ, **env), timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[tag] = {f: (d / f).read_bytes() for f in sorted(os.listdir(d)) if f.endswith((".prof", ".vtk"))}
        logs[tag] = (d / "t.log").read_text()
    assert list(outs["one"]) == list(outs["four"]) and len(outs["one"]) >= 3
    for f in outs["one"]:
        assert outs["one"][f] == outs["four"][f], f
    assert "re-balanced the slabs" in logs["four"]


@pytest.mark.parametrize("ngpu", ["1", "2"])
def test_driver_runs_from_the_preprocessor_input(ngpu, tmp_path):
    """a grid argument ending in .boid: the particles are generated on the device(s) (mphx_upload_generated /
    mphx_multi_upload_generated) from the pre-processor's own input -- every output file equals the run from the .grid text"""
    c = cases.tiny3d()
    c.rc.end_time = 3.5 * c.params.dt
    c.rc.output_interval = 2.0 * c.params.dt
    c.rc.vtk_output_interval = 3.0 * c.params.dt
    outs = {}
    env = dict(os.environ)
    if ngpu != "1":
        env.update(MPHX_NGPU=ngpu, MPHX_DEVICES=",".join(["0"] * int(ngpu)), CUDA_DEVICE_MAX_CONNECTIONS="32")
    for tag, grid in (("grid", "c.grid"), ("boid", "c.boid")):
        d = tmp_path / tag
        d.mkdir()
        cases.write_grid_file(str(d / "c.grid"), c)
        cases.write_boid_file(str(d / "c.boid"), c)
        cases.write_data_file(str(d / "t.data"), c.params, c.rc)
        r = subprocess.run([EXE, "t.data", grid, "t%03d.prof", "t%03d.vtk", "t.log", "4", "3", "dam"], cwd=d, capture_output=True,
                           text=True, env=env, timeout=300)
        assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-2000:])
        outs[tag] = {f: (d / f).read_bytes() for f in sorted(os.listdir(d)) if f.endswith((".prof", ".vtk"))}
    assert list(outs["grid"]) == list(outs["boid"]) and len(outs["grid"]) == 5
    for f in outs["grid"]:
        assert outs["grid"][f] == outs["boid"][f], f
