#!/usr/bin/env python
"""developer timing of the solid sub-step kernels alone (GPU box): one-thread-per-solid against a team per solid, on the
full set of solids and on the share one rank of k would advance (MPHX_DEBUG_SUBSTEP_SHARE: timing only, wrong physics)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from particlemethod_fsi_b200 import Solver, cases

case = cases.fsi3d_for_count(float(sys.argv[1]) if len(sys.argv) > 1 else 1.0e7)
print("particles", case.n, "solids", case.counts()[1], flush=True)
res = {}
for share in (1, 2, 4, 8):
    for name, env in (("one thread", {"MPHX_SOLID_TEAM": "0"}), ("one thread, 8 gathers in flight", {"MPHX_SOLID_TEAM": "0", "MPHX_DEBUG_DEEP": "1"}),
                      ("team of 16", {"MPHX_SOLID_TEAM": "2"})):
        os.environ.pop("MPHX_DEBUG_DEEP", None)
        os.environ["MPHX_DEBUG_SUBSTEP_SHARE"] = str(share)
        os.environ.update(env)
        s = Solver.from_case(case)
        s.set_overlap(False)
        s.step(3, sync=True)
        s.set_timing(True)
        s.step(6, sync=True)
        ms = s.kernel_timers_ms()[4] / 6
        s.set_timing(False)
        s.close()
        res[f"share 1/{share} {name}"] = round(ms, 4)
        print(f"share 1/{share}  {name:32s} {ms:.4f} ms per step (5 sub-steps, 10 kernels)", flush=True)
print(json.dumps(res))
