"""Slab decomposition parity (-m gpu): the ring of x-slabs must reproduce the single-context step.

All slabs of the ring run in ONE process on ONE device (slab.MultiSolver with a repeated device: the
mailboxes are ordinary device memory then), so this exercises every slab kernel -- votes, migration, halo
pack / push / unpack, halo refresh on list-reuse steps, PressureP exchange, replicated solids -- and the
device-side flag protocol on the single-GPU box.  The in-bucket order (by original id) and the stencil
order are the same in a slab and in the whole domain, and the rebuild / reuse decision is the OR of the
slabs' votes (= the single context's decision), so the sums are the same bits.
"""
import numpy as np
import pytest

from particlemethod_fsi_b200 import Solver, cases, slab

pytestmark = pytest.mark.gpu

FIELDS = ("position", "velocity", "pressure_p", "force", "cell_index", "stress", "property")


def _compare(case, world, steps, exact=True, fields=FIELDS, atol=None, list_reuse=None, rtol=1e-12):
    ref = Solver.from_case(case, list_reuse=list_reuse)
    ring = slab.MultiSolver(case, world, devices=[0] * world, list_reuse=list_reuse)
    done = 0
    for target in steps:
        ref.step(target - done, sync=True)
        ring.step(target - done)
        ring.sync()
        done = target
        a = ref.download(*fields)
        b = ring.download(*fields)
        for f in fields:
            if exact or f in ("cell_index", "property"):
                assert np.array_equal(a[f], b[f]), (case.name, world, target, f, float(np.abs(a[f] - b[f]).max()))
            else:
                scale = max(float(np.abs(a[f]).max()), 1e-300)
                assert float(np.abs(a[f] - b[f]).max()) <= rtol * scale + (atol or {}).get(f, 0.0), (case.name, world, target, f)
        assert ring.time == ref.time
    info = ring.info()
    st = ref.status()
    assert all(i["err"] == 0 for i in info), info
    # every slab took the same rebuild / reuse decisions as the single context
    assert all((i["builds"], i["reuses"]) == (st["builds"], st["reuses"]) for i in info), (info, st)
    ref.close()
    ring.close()
    return info


@pytest.mark.parametrize("world", [2, 3, 4])
def test_fsi3d_ring_equals_single_context(world):
    info = _compare(cases.fsi3d_mini(), world, [1, 5, 40])
    assert info[0]["reuses"] > 0   # (the candidate lists really were reused: the halo-refresh path ran)


@pytest.mark.parametrize("name,world", [("dam2d", 2), ("dam2d", 4), ("fsi2d", 3), ("bar2d", 2)])
def test_2d_ring_equals_single_context(name, world):
    _compare(getattr(cases, name)(), world, [1, 10, 60])


def test_ring_without_list_reuse_equals_single_context():
    """every step rebuilds (the reference's own schedule): migration + fresh halos every step"""
    _compare(cases.fsi3d_mini(), 3, [1, 12], list_reuse=False)


@pytest.mark.parametrize("name,world", [("tiny2d", 2), ("tiny3d", 3)])
def test_surface_tension_ring_equals_single_context(name, world):
    """a16/a17 on slabs: the second exchange also carries PressureA and GravityCenter of the halo copies"""
    case = getattr(cases, name)()
    case.params.surface_tension[0] = case.params.surface_tension[1] = 0.072
    case.params.interaction_ratio[1][4] = 0.6
    _compare(case, world, [1, 10, 40], fields=FIELDS + ("pressure_a", "gravity_center", "density_a"))


def test_particles_migrate_between_slabs():
    """give the fluid a uniform x velocity so that particles cross slab faces every few steps"""
    case = cases.dam2d()
    fl = case.property < 2
    case.velocity[fl, 0] = 1.5   # 1.5 m/s * 1e-4 s = 0.15 l0 per step
    info = _compare(case, 4, [1, 20, 120])
    assert sum(i["held"] for i in info) >= case.n
    assert info[0]["builds"] > 10


def test_flow_through_the_periodic_seam():
    """no walls: a fluid block drifting through the periodic x seam (ring closure rank 0 <-> rank W-1)"""
    case = cases.tiny2d()
    keep = case.property < 2
    sub = cases.Case("seam", case.params.copy(), case.rc, case.property[keep].copy(), case.position[keep].copy(),
                     case.initial_position[keep].copy(), case.velocity[keep].copy())
    sub.params.gravity[1] = 0.0
    w = sub.params.domain_max[0] - sub.params.domain_min[0]
    sub.position[:, 0] += (sub.params.domain_max[0] - sub.position[:, 0].max()) - 0.5 * sub.params.particle_spacing
    sub.velocity[:, 0] = 2.0
    assert w > 0
    # the ghost copies that cross the seam are shifted by the box width, which rounds differently from
    # the reference's Mod-based minimum image: 1e-12, not bit-equal
    # (PressureP of a force-free drifting block is round-off around zero: compare the kinematics)
    _compare(sub, 2, [1, 30, 200], exact=False, fields=("position", "velocity", "cell_index", "property"))


def test_solid_next_to_the_periodic_seam():
    """a replicated solid keeps its global x on every slab.  Here the plate touches the right edge
    of the periodic box and the water the left edge, so on rank 0 the plate is a neighbour THROUGH the seam
    (its gather record must sit one box width to the left) and on the last rank the water arrives as shifted
    ghosts.  Ring == single context (1e-12: the +-W shift rounds differently from the Mod-based minimum image)."""
    from particlemethod_fsi_b200 import abi
    l0 = 2.0e-3
    p, rc = cases.default_params(2, abi.MODULE_DAM)
    p.elastic_dt = 2.0e-5
    cubs = [cases.Cuboid(1, (0.0, 0.0, 0.0), (0.012, 0.03, l0), l0),
            cases.Cuboid(2, (0.08 - 3 * l0, 0.0, 0.0), (0.08, 0.024, l0), l0),
            cases.Cuboid(4, (0.0, -3 * l0, 0.0), (0.08, 0.0, l0), l0)]
    case = cases._assemble("seam_solid", p, rc, l0, (0.0, -7 * l0, 0.0), (0.08, 0.06, l0), cubs)
    assert case.counts()[1] > 0
    fields = ("position", "velocity", "pressure_p", "force", "cell_index", "property")
    # (PressureP of water at rest is rounding noise around zero, ~1e-11 for kappa = 1e4: absolute floors; the one-ulp
    # differences of the shifted separations feed kappa (sum w - N0p) and grow to ~1e-10 of the velocity in 60 steps)
    for world in (2, 3):
        _compare(case, world, [1, 10], exact=False, fields=fields, atol=dict(pressure_p=1e-9, force=1e-12))
        _compare(case, world, [60], exact=False, fields=fields, atol=dict(pressure_p=1e-9, force=1e-12), rtol=1e-9)


def test_compact_owned_io_round_trip():
    """mphx_download_owned / mphx_upload_owned: rows of every slab together are the whole case (+ the
    replicated solids once per slab); taking them back unchanged and stepping equals plain stepping."""
    case = cases.fsi3d_mini()
    ref = Solver.from_case(case)
    ring = slab.MultiSolver(case, 3, devices=[0, 0, 0])
    ref.step(3, sync=True)
    ring.step(3)
    ring.sync()
    ns = case.counts()[1]
    seen = np.zeros(case.n, dtype=np.int64)
    want = ref.download("position", "velocity")
    bufs = []
    for r in range(3):
        m, ids, x, v = ring.download_owned(r)
        np.add.at(seen, ids[:m], 1)
        assert np.array_equal(x[:m], want["position"][ids[:m]])
        assert np.array_equal(v[:m], want["velocity"][ids[:m]])
        bufs.append((r, m, ids, x, v))
    solid = (case.property >= 2) & (case.property < 4)
    assert np.all(seen[~solid] == 1) and np.all(seen[solid] == 3) and int(solid.sum()) == ns
    for r, m, ids, x, v in bufs:
        assert ring.upload_owned(r, m, ids, x, v) == 0
    # (an upload replaces the state: both sides rebuild their lists at the next step, so the sums stay the same bits)
    ids0 = np.empty(case.n, dtype=np.int32)
    x0, v0 = np.empty((case.n, 3)), np.empty((case.n, 3))
    m0 = ref.download_owned(ids0, x0, v0)
    ref.upload_owned(m0, ids0, x0, v0)
    ref.step(4, sync=True)
    ring.step(4)
    ring.sync()
    a, b = ref.download("position", "velocity"), ring.download("position", "velocity")
    assert np.array_equal(a["position"], b["position"]) and np.array_equal(a["velocity"], b["velocity"])
    # a stale upload (a step happened since the download) is refused
    r, m, ids, x, v = bufs[0]
    assert ring.upload_owned(r, m, ids, x, v) != 0
    # single context: owned = everything; modified rows are taken over
    ids = np.empty(case.n, dtype=np.int32)
    x = np.empty((case.n, 3))
    v = np.empty((case.n, 3))
    m = ref.download_owned(ids, x, v)
    assert m == case.n and np.array_equal(np.sort(ids), np.arange(case.n))
    v[:, 1] += 0.25
    ref.upload_owned(m, ids, x, v)
    got = ref.download("velocity")["velocity"]
    assert np.array_equal(got[ids], v)
    ref.close()
    ring.close()


def test_rebalanced_ring_equals_single_context():
    """In-place re-balancing (mphx_multi_rebalance): a fluid moving in +x unbalances the slabs; every 15 steps the cuts move
    by up to one halo width and the particles that change owner travel with the migration of the forced rebuild.  With the
    reference's own schedule (every step rebuilds) the ring stays bit-identical to the single context."""
    case = cases.dam2d()
    fl = case.property < 2
    case.velocity[fl, 0] = 1.5
    ref = Solver.from_case(case, list_reuse=False)
    ring = slab.MultiSolver(case, 4, devices=[0] * 4, list_reuse=False)
    first = ring.columns()
    moved_total = 0
    for block in range(8):
        ref.step(15, sync=True)
        ring.step(15)
        ring.sync()
        moved_total += ring.rebalance()
    ref.step(5, sync=True)
    ring.step(5)
    ring.sync()
    assert moved_total > 0 and ring.columns() != first
    a, b = ref.download(*FIELDS), ring.download(*FIELDS)
    for f in FIELDS:
        assert np.array_equal(a[f], b[f]), (f, float(np.abs(a[f] - b[f]).max()))
    info = ring.info()
    assert all(i["err"] == 0 for i in info), info
    ref.close()
    ring.close()


def test_rebalance_with_list_reuse_stays_within_tolerance():
    """the same with candidate-list reuse on: a re-cut forces a rebuild the single context does not make, so the sums run over
    lists of different age -- equal to round-off (1e-12), buckets exact"""
    case = cases.fsi3d_mini()
    ref = Solver.from_case(case)
    ring = slab.MultiSolver(case, 3, devices=[0] * 3)
    for block in range(3):
        ref.step(10, sync=True)
        ring.step(10)
        ring.sync()
        ring.rebalance()
    ref.step(3, sync=True)
    ring.step(3)
    ring.sync()
    a, b = ref.download(*FIELDS), ring.download(*FIELDS)
    for f in FIELDS:
        if f in ("cell_index", "property"):
            assert np.array_equal(a[f], b[f]), f
        else:
            scale = max(float(np.abs(a[f]).max()), 1e-300)
            tol = 1e-12 * scale + (1e-9 * scale if f in ("pressure_p", "force") else 0.0)
            assert float(np.abs(a[f] - b[f]).max()) <= tol, (f, float(np.abs(a[f] - b[f]).max()), scale)
    assert all(i["err"] == 0 for i in ring.info())
    ref.close()
    ring.close()


@pytest.mark.parametrize("name,world", [("fsi3d_mini", 3), ("fsi2d", 2)])
def test_generated_ring_equals_uploaded_single_context(name, world):
    """SURVEY 8(f) N4 on slabs: every slab fills the generator's lattice on its device and keeps its share (histogram from the
    axis tables, keep mask + scan + ids on the device: no particle array on the host); the run equals the single context fed
    with the .grid arrays, bit for bit"""
    case = getattr(cases, name)()
    ref = Solver.from_case(case)
    ring = slab.MultiSolver(case, world, devices=[0] * world, generated=True)
    ref.step(12, sync=True)
    ring.step(12)
    ring.sync()
    a, b = ref.download(*FIELDS), ring.download(*FIELDS)
    for f in FIELDS:
        assert np.array_equal(a[f], b[f]), (f, float(np.abs(a[f] - b[f]).max()))
    assert all(i["err"] == 0 for i in ring.info())
    ref.close()
    ring.close()
