"""torchrun --nproc-per-node N tools/slab_dist_check.py [case] [steps]: the NCCL ring of x-slabs against
the single-context step (rank 0 runs both and compares).  Developer / GPU-box script."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from particlemethod_fsi_b200 import Solver, cases, slab

name = sys.argv[1] if len(sys.argv) > 1 else "fsi3d_mini"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
case = getattr(cases, name)() if hasattr(cases, name) else cases.fsi3d_for_count(float(name))
ring = slab.SlabSolver(case, slab.DistTransport(), device=dev)
ring.step(steps)
ring.sync()
got = ring.download("position", "velocity", "pressure_p", "cell_index")
info = ring.info()[0]
print(f"rank {rank}: columns {info['columns']} held {info['held']} ghosts {info['ghosts']}", flush=True)
if rank == 0:
    ref = Solver.from_case(case, device=local)
    ref.step(steps, sync=True)
    want = ref.download("position", "velocity", "pressure_p", "cell_index")
    for f in want:
        d = float(np.abs(want[f].astype(np.float64) - got[f]).max())
        print(f"{name} N={case.n} world={world} steps={steps} {f}: max|diff| = {d:.3e} equal={np.array_equal(want[f], got[f])}", flush=True)
        assert d <= 1e-12 * max(float(np.abs(want[f]).max()), 1e-300), f
    print("SLAB_DIST_OK", flush=True)
ring.close()
dist.destroy_process_group()
