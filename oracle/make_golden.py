"""Generate the golden vectors under tests/golden/ FROM THE REFERENCE ITSELF -- TEST INFRASTRUCTURE.

Runs only where /root/reference has been compiled into oracle/_ref (oracle/build_ref.sh).  The
reference ships no tests or golden vectors, so these fixtures are dumps of
its own global arrays (through oracle/_ref/libref_*.so, the untouched src/main.cpp plus the
same-TU harness) and files written by its own executable.  Inputs are NOT stored: the cases are
rebuilt deterministically by particlemethod_fsi_b200/cases.py at test time (test_cases.py pins that
builder against the shipped results/Dam/dam.grid).

    python -m oracle.make_golden
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.refharness import RefHarness, REF_DIR  # noqa: E402
from particlemethod_fsi_b200 import cases  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
FIELDS = ["Position", "Velocity", "Force", "Acceleration", "PressureP", "VolStrainP", "DivergenceP",
          "NeighborCount", "InitialStructureNeighborCount", "Normalizer", "DeformGradient", "Strain",
          "Stress", "LambdaLames", "MuLames"]
SCALARS = ["N0a", "N0p", "Swa", "Swg", "Swp", "Swv", "R2g", "MaxRadius", "RadiusA", "RadiusP", "RadiusV",
           "ParticleVolume", "CofK"]
VARIANT = {(2, 1): "2d_bar", (2, 2): "2d_dam", (3, 2): "3d_dam"}


def cell_of_particle(h):
    ci = h.view("CellIndex")[: h.n]
    cp = h.view("CellParticle")[: h.n]
    out = np.empty(h.n, dtype=np.int32)
    out[cp] = ci
    return out


def neighbor_csr(h):
    cnt = h.get("NeighborCount")
    nb = h.view("Neighbor")
    rows = [np.sort(nb[i, :cnt[i]]) for i in range(h.n)]
    off = np.zeros(h.n + 1, dtype=np.int64)
    off[1:] = np.cumsum(cnt)
    return off, np.concatenate(rows).astype(np.int32) if rows else np.zeros(0, np.int32)


def dump(case, steps, with_text=False, fields=FIELDS):
    variant = VARIANT[(case.params.dim, case.params.clamp_module)]
    d = tempfile.mkdtemp()
    cases.write_grid_file(d + "/c.grid", case)
    cases.write_data_file(d + "/c.data", case.params, case.rc)
    h = RefHarness(variant, d + "/c.data", d + "/c.grid", nthreads=4)
    h.init()
    out = {"variant": np.array(variant), "n": np.array(case.n)}
    for s in SCALARS:
        out["const_" + s] = np.array(h.double(s))
    out["const_CofA"] = h.get("CofA")
    out["const_WallRotation"] = h.get("WallRotation")
    done = 0
    for target in steps:
        if target > done:
            h.step(target - done)
            done = target
        for f in fields:
            a = h.get(f)
            if a.dtype == np.int32:
                out[f"s{target}_{f}"] = a
            else:
                out[f"s{target}_{f}"] = a
        out[f"s{target}_CellIndex"] = cell_of_particle(h)
        off, ids = neighbor_csr(h)
        out[f"s{target}_NeighborSetsSha"] = np.array(hashlib.sha256(off.tobytes() + ids.tobytes()).hexdigest())
        out[f"s{target}_Time"] = np.array(h.double("Time"))
    np.savez_compressed(os.path.join(OUT, f"{case.name}.npz"), **out)
    print(f"{case.name}: N={case.n} variant={variant} steps={steps} -> {case.name}.npz")
    if with_text:
        # files written by the reference executable itself (its CLI, src/main.cpp:501-508)
        case.rc.end_time = 3.5 * case.params.dt
        case.rc.output_interval = 2.0 * case.params.dt
        case.rc.vtk_output_interval = 3.0 * case.params.dt
        cases.write_data_file(d + "/t.data", case.params, case.rc)
        exe = os.path.join(REF_DIR, f"Mph_Elastic_Explicit_{variant}")
        subprocess.run([exe, "t.data", "c.grid", "t%03d.prof", "t%03d.vtk", "t.log", "2"], cwd=d, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for fn in ("t000.prof", "t002.prof", "t000.vtk", "t003.vtk", "output.vtk"):
            dst = os.path.join(OUT, f"{case.name}_{fn}.gz")
            import gzip
            with open(os.path.join(d, fn), "rb") as fi, gzip.GzipFile(dst, "wb", mtime=0) as fo:
                fo.write(fi.read())
        print(f"{case.name}: reference CLI text outputs stored")
    h.close()


def main():
    os.makedirs(OUT, exist_ok=True)
    dump(cases.tiny2d(), [0, 1, 20], with_text=True)
    dump(cases.tiny3d(), [0, 1, 10], with_text=True)
    # C1 (the reference's own shipped case): aggregate-size fixture, three fields at step 100
    dump(cases.dam2d(), [100], fields=["Position", "Velocity", "PressureP", "NeighborCount"])
    with open(os.path.join(OUT, "README.md"), "w") as f:
        f.write("Golden vectors dumped from the reference itself by `python -m oracle.make_golden`\n"
                "(oracle/_ref/libref_*.so = untouched src/main.cpp + same-TU harness, and the reference CLI).\n"
                "Inputs are rebuilt by particlemethod_fsi_b200/cases.py; dam2d equals results/Dam/dam.grid.\n")


if __name__ == "__main__":
    main()
