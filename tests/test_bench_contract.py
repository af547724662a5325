"""bench.py contract on CPU: the reference arm runs here (the reference's CPU build or the oracle port
on a small bounded sample) and prints one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-particles", "2e4"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "particle_steps_per_sec" and line["unit"] == "particle-steps/s"
    assert line["value"] > 0 and line["vs_baseline"] is None and line["dtype"] == "f64"
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]


def test_non_zero_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--particles", "2e4"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_reference_arm_never_maps_the_product_library():
    """the reference arm must run the reference's CPU build only -- importing what it imports (bench,
    cases, the reference harness) must not load libmphx.so (the package binds it lazily, at the first call)"""
    code = ("import sys; sys.path.insert(0, %r); import bench; from particlemethod_fsi_b200 import cases; "
            "from oracle import refharness; c = cases.tiny2d(); print('libmphx' in open('/proc/self/maps').read())" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().splitlines()[-1] == "False"


def test_committed_bench_lines_keep_the_contract():
    """the lines the final GPU runs printed (profiles/r02b_bench_*.json: produced by bench.py on the B200 box) carry every key
    the driver and the judge read -- a guard against a bench.py edit that drops one"""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02b_bench_*gpu*.json")))
    assert len(files) >= 5
    for fn in files:
        line = json.loads([l for l in open(fn).read().splitlines() if l.startswith("{")][-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                  "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
            assert k in line, (fn, k)
        assert line["metric"] == "particle_steps_per_sec" and line["unit"] == "particle-steps/s" and line["dtype"] == "f64"
        assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["data"] == "synthetic"
        assert "workload" in line["config"] and "model" not in line["config"]
        assert abs(line["value"] - line["config"]["particles"] / (line["ms_per_step"] * 1e-3)) <= 1e-6 * line["value"]
        e = line["e2e"]
        assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < line["value"]
        r = line["roofline"]
        for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
            assert k in r, (fn, k)
        assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
        assert line["gpu_launches"] > 0
        if line["n_gpus"] == 1:
            assert r["fp64"]["peak_measured_tflops"] > 10 and 0 < r["fp64"]["pass2"]["frac"] < 1
        else:
            v = line["verify"]
            assert v["ring_vs_single_context"]["ok"] is True and v["timed_state"]["ok"] is True
            assert all(f["bit_equal"] for f in v["ring_vs_single_context"]["fields"].values())
            assert line["clocks"]["reasons"] == [] or all("thermal" not in x and "hw_slowdown" not in x for x in line["clocks"]["reasons"])
