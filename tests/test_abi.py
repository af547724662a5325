"""The C-ABI library loads without a GPU and exports exactly what include/mphx.h declares."""
import ctypes as C
import os
import re
import subprocess

from conftest import ROOT
from particlemethod_fsi_b200 import abi, solver


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "mphx.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mphx_[a-z_0-9]+)\s*\(", txt)))


def test_header_matches_declared_exports():
    assert header_symbols() == sorted(abi.EXPORTS)


def test_library_exports_every_symbol():
    so = solver.LIB_PATH
    out = subprocess.check_output(["nm", "-D", "--defined-only", so], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = [s for s in abi.EXPORTS if s not in exported]
    assert not missing, missing
    for s in abi.EXPORTS:  # and ctypes can resolve them
        getattr(solver.lib, s)


def test_struct_layout_matches_c():
    L = solver.lib
    L.mphx_abi_sizeof.argtypes = [C.c_int]
    assert L.mphx_abi_sizeof(0) == C.sizeof(abi.Params)
    assert L.mphx_abi_sizeof(1) == C.sizeof(abi.RunControl)
    assert L.mphx_abi_sizeof(2) == C.sizeof(abi.Constants)
    assert L.mphx_abi_sizeof(3) == C.sizeof(abi.HostViews)
    assert L.mphx_version() == 100


def test_library_carries_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "-lelf", solver.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_device():
    """On a box without a B200 the compute entry points must fail loudly, not fall back."""
    if solver.device_count() > 0:
        return
    from particlemethod_fsi_b200 import cases
    import pytest
    with pytest.raises(solver.MphxError) as e:
        solver.Solver(cases.tiny2d().params)
    assert e.value.code == abi.MPHX_ERR_NO_DEVICE


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "particlemethod_fsi_b200")
    for dp, _dn, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, fn), errors="replace").read()
                assert "oracle" not in txt.replace("oracle/", "").replace("the oracle", "") or fn == "cases.py", fn
