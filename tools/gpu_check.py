"""Developer script: run the CUDA path against the oracle on the small cases and print error norms."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import particlemethod_fsi_b200 as pm
from particlemethod_fsi_b200 import cases
from oracle.oracle import Oracle

MAP = dict(position='Position', velocity='Velocity', force='Force', acceleration='Acceleration',
           pressure_p='PressureP', vol_strain_p='VolStrainP', divergence_p='DivergenceP',
           neighbor_count='NeighborCount', initial_structure_neighbor_count='InitialStructureNeighborCount',
           normalizer='Normalizer', deform_gradient='DeformGradient', strain='Strain', stress='Stress',
           lambda_lames='LambdaLames', mu_lames='MuLames')

def rel(a, b):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64)).max() if a.size else 0.0
    s = np.abs(b).max() if b.size else 0.0
    return d / s if s > 0 else d

def run(case, steps_list):
    o = Oracle.from_case(case); o.init()
    s = pm.Solver.from_case(case)
    print(f"== {case.name} N={case.n} {case.counts()} dim={case.params.dim}")
    done = 0
    for target in steps_list:
        k = target - done
        if k > 0:
            t = time.time(); s.step(k, sync=True); tg = time.time() - t
            t = time.time(); o.step(k); to = time.time() - t
            done = target
        else:
            tg = to = 0.0
        got = s.download(*MAP.keys(), 'cell_index')
        errs = {f: rel(got[f], o.get(r)) for f, r in MAP.items()}
        worst = {f: e for f, e in errs.items() if e > 1e-12}
        cell_ok = np.array_equal(got['cell_index'], o.cell_of_particle())
        print(f" step {target}: gpu {tg:.3f}s oracle {to:.3f}s cells_equal={cell_ok} max_err={max(errs.values()):.3e} >1e-12: "
              + ", ".join(f"{f}={e:.2e}" for f, e in worst.items()))
    # neighbour sets
    off, ids = s.neighbors()
    cnt, sets = o.neighbor_sets()
    same = all(np.array_equal(ids[off[i]:off[i+1]], sets[i]) for i in range(case.n))
    print(f" neighbour sets identical: {same} (total {off[-1]})")
    off, ids = s.initial_structure_neighbors()
    nb = o.view('InitialStructureNeighbor'); c0 = o.get('InitialStructureNeighborCount')
    same = all(np.array_equal(ids[off[i]:off[i+1]], np.sort(nb[i, :c0[i]])) for i in range(case.n))
    print(f" initial structure sets identical: {same} (total {off[-1]})")
    s.close(); o.close()

if __name__ == '__main__':
    which = sys.argv[1:] or ['dam2d', 'bar2d', 'fsi2d', 'fsi3d_mini']
    for w in which:
        c = getattr(cases, w)()
        run(c, [0, 1, 10, 100] if w != 'bar2d' else [0, 1, 10, 50])
