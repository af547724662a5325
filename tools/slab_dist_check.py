"""torchrun --nproc-per-node N tools/slab_dist_check.py [case] [steps] [rebalance_every]: the ring of x-slabs, one process
per GPU (CUDA IPC mailboxes, NVLink), against the single-context step (rank 0 runs both and compares).  With
rebalance_every > 0 the slabs are re-cut in place while the ring steps (DistSlab.rebalance: column histograms all-reduced
over NCCL) and every step rebuilds (the reference's schedule), so the comparison stays bit for bit.  GPU-box script."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from particlemethod_fsi_b200 import Solver, cases, slab

name = sys.argv[1] if len(sys.argv) > 1 else "fsi3d_mini"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
every = int(sys.argv[3]) if len(sys.argv) > 3 else 0
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
case = getattr(cases, name)() if hasattr(cases, name) else cases.fsi3d_for_count(float(name))
if every > 0:
    case.velocity[case.property < 2, 0] = 1.5     # the fluid drifts in +x: the slabs unbalance
reuse = False if every > 0 else None
ring = slab.DistSlab(case, local, list_reuse=reuse)
first = list(ring.partition)
done = moved = 0
while done < steps:
    k = min(every, steps - done) if every > 0 else steps
    ring.step(k)
    ring.sync()
    done += k
    if every > 0 and done < steps:
        moved += ring.rebalance()
fields = ("position", "velocity", "pressure_p", "cell_index")
got = ring.download(*fields)
st = ring.status()
print(f"rank {rank}: columns {ring.partition[rank]} (first {first[rank]}) held {st['held']} ghosts {st['ghosts']} err {st['err']} cuts moved {moved}", flush=True)
if rank == 0:
    ref = Solver.from_case(case, device=local, list_reuse=reuse)
    ref.step(steps, sync=True)
    want = ref.download(*fields)
    for f in fields:
        d = float(np.abs(want[f].astype(np.float64) - got[f]).max())
        print(f"{name} N={case.n} world={world} steps={steps} {f}: max|diff| = {d:.3e} bit-equal={np.array_equal(want[f], got[f])}", flush=True)
        assert np.array_equal(want[f], got[f]), f
    assert every == 0 or moved > 0
    print("SLAB_DIST_OK", flush=True)
ring.close()
dist.destroy_process_group()
