"""Synthetic inputs for the BASELINE.json configs (host-side plumbing, no compute).

The reference's pre-processor (generator/generator.cpp) fills each `Cuboid` of a `.boid` file with
a lattice (generator/generator.cpp:654-680: start at lower+0.5*spacing, accumulate `p += spacing`
while `p < upper-0.49*spacing`, x outer / y / z inner) and writes the `.grid` text with `%e`
(generator/generator.cpp:839-862).  `/root/reference` is absent on the GPU box, so the same lattice
rule is restated here; tests/test_cases.py checks that `dam2d()` reproduces the shipped
results/Dam/dam.grid byte for byte (and, where oracle/_ref/GeneratorForMph exists, arbitrary boxes).

Coordinates are passed through the `%e` round trip (7 significant digits) so that arrays built in
memory equal what the solver would parse from the written file.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import abi


@dataclass
class Cuboid:
    type: int
    lower: tuple
    upper: tuple
    spacing: float
    velocity: tuple = (0.0, 0.0, 0.0)


@dataclass
class Case:
    name: str
    params: abi.Params
    rc: abi.RunControl
    property: np.ndarray          # int32 [N]
    position: np.ndarray          # float64 [N,3]
    initial_position: np.ndarray  # float64 [N,3]
    velocity: np.ndarray          # float64 [N,3]
    cuboids: list = field(default_factory=list)

    @property
    def n(self) -> int:
        return int(self.property.shape[0])

    def counts(self):
        t = self.property
        return int(((t >= 0) & (t < 2)).sum()), int(((t >= 2) & (t < 4)).sum()), int((t >= 4).sum())


def _e(v: float) -> float:
    """value after a printf("%e") / sscanf("%lf") round trip"""
    return float("%e" % v)


def _axis(lo: float, hi: float, space: float) -> np.ndarray:
    width = hi - lo
    count = int(round(width / space))  # C round(): half away from zero; widths here are never x.5
    spacing = width / count
    out = []
    p = lo + 0.5 * spacing
    while p < hi - 0.49 * spacing:
        out.append(_e(p))
        p += spacing
    return np.asarray(out, dtype=np.float64)


def cuboid_fill(cub: Cuboid) -> np.ndarray:
    ax = [_axis(cub.lower[d], cub.upper[d], cub.spacing) for d in range(3)]
    X, Y, Z = np.meshgrid(ax[0], ax[1], ax[2], indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)


# ---- default material tables = results/Dam/dam.data -------------------------------------------
def default_params(dim: int, module: int) -> tuple[abi.Params, abi.RunControl]:
    p = abi.Params()
    rc = abi.RunControl()
    p.dim = dim
    p.clamp_module = module
    p.ref_compat = abi.COMPAT_DOUBLE_UPDATE
    p.time0 = 0.0
    p.dt = 1.0e-4
    p.elastic_dt = 1.0e-4
    rc.output_interval = 1.0
    rc.vtk_output_interval = 1.0e-2
    rc.end_time = 1.0
    p.radius_ratio_a = p.radius_ratio_p = p.radius_ratio_v = 2.5
    tab = dict(
        density=[1.0e3, 1.0e3, 1.1e3, 1.0e3, 1.0e3, 6.0e3],
        bulk_modulus=[1.0e4, 1.0e4, 1.0e4, 1.0e6, 1.0e4, 1.0e5],
        bulk_viscosity=[1.0e1, 1.0e-1, 1.0e-1, 1.0e3, 1.0e-1, 1.0e2],
        shear_viscosity=[1.0e-2, 1.0e-3, 1.0e-2, 1.0e-1, 1.0e3, 1.0e-1],
        surface_tension=[0.0] * 6,
        # the file lists 4 values that land in types 2..5 (src/main.cpp:757-758)
        young_modulus=[0.0, 0.0, 1e5, 1e5, 1e8, 1e4],
        poisson_ratio=[0.0, 0.0, 0.2, 0.4, 0.3, 0.3],
    )
    for k, v in tab.items():
        arr = getattr(p, k)
        for i in range(6):
            arr[i] = v[i]
    for i in range(6):
        for j in range(6):
            p.interaction_ratio[i][j] = 1.0
    p.gravity[0], p.gravity[1], p.gravity[2] = 0.0, -1.0, 0.0
    return p, rc


def _assemble(name, p, rc, l0, dom_lo, dom_hi, cuboids) -> Case:
    p.particle_spacing = _e(l0)
    for d in range(3):
        p.domain_min[d] = _e(dom_lo[d])
        p.domain_max[d] = _e(dom_hi[d])
    xs, ts, vs = [], [], []
    for cub in cuboids:
        x = cuboid_fill(cub)
        xs.append(x)
        ts.append(np.full(x.shape[0], cub.type, dtype=np.int32))
        vs.append(np.tile(np.asarray([_e(c) for c in cub.velocity]), (x.shape[0], 1)))
    x = np.ascontiguousarray(np.concatenate(xs))
    t = np.ascontiguousarray(np.concatenate(ts))
    v = np.ascontiguousarray(np.concatenate(vs))
    # each class must be contiguous in file order (src/main.cpp:909-929)
    cls = np.where(t < 2, 0, np.where(t < 4, 1, 2))
    assert np.all(np.diff(cls[np.argsort(cls, kind="stable")]) >= 0)
    chg = np.flatnonzero(np.diff(cls)) + 1
    assert len(set(cls[np.r_[0, chg]])) == len(np.r_[0, chg]), "particle classes must be contiguous"
    return Case(name, p, rc, t, x, x.copy(), v, cuboids)


def dam2d() -> Case:
    """C1: results/Dam (2D dam break, 4850 fluid + 1800 wall = 6650 particles)."""
    p, rc = default_params(2, abi.MODULE_BAR)
    l0 = 0.001
    cubs = [
        Cuboid(1, (0.0, 0.003, 0.0), (0.05, 0.10, 0.001), l0),
        Cuboid(4, (0.0, 0.0, 0.0), (0.2, 0.003, 0.001), l0),
        Cuboid(4, (0.2, 0.0, 0.0), (0.203, 0.20, 0.001), l0),
        Cuboid(4, (-0.003, 0.0, 0.0), (0.0, 0.20, 0.001), l0),
    ]
    return _assemble("dam2d", p, rc, l0, (-0.01, 0.0, 0.0), (0.21, 0.40, 0.001), cubs)


def bar2d(l0: float = 1.0e-3, length: float = 0.2, height: float = 0.02, tip_velocity: float = 0.0) -> Case:
    """C2: 2D total-Lagrangian cantilever (solid only), clamp x0<0.001 (Bar_Module)."""
    p, rc = default_params(2, abi.MODULE_BAR)
    p.elastic_dt = 1.0e-5
    cubs = [Cuboid(2, (0.0, 0.0, 0.0), (length, height, l0), l0, (0.0, tip_velocity, 0.0))]
    return _assemble("bar2d", p, rc, l0, (-0.05, -0.2, 0.0), (length + 0.15, 0.2, l0), cubs)


def fsi2d(l0: float = 1.0e-3, tank=(0.4, 0.3), water=(0.1, 0.2), plate_x: float = 0.2,
          plate_t: float = 0.004, plate_h: float = 0.08, elastic_dt: float = 2.0e-5) -> Case:
    """C3: 2D dam break hitting an elastic plate clamped at the floor (DAM_Module, y0<0.002)."""
    p, rc = default_params(2, abi.MODULE_DAM)
    p.elastic_dt = elastic_dt
    w = 3 * l0
    Lx, Ly = tank
    cubs = [
        Cuboid(1, (0.0, 0.0, 0.0), (water[0], water[1], l0), l0),
        Cuboid(2, (plate_x, 0.0, 0.0), (plate_x + plate_t, plate_h, l0), l0),
        Cuboid(4, (-w, -w, 0.0), (Lx + w, 0.0, l0), l0),
        Cuboid(4, (-w, 0.0, 0.0), (0.0, Ly, l0), l0),
        Cuboid(4, (Lx, 0.0, 0.0), (Lx + w, Ly, l0), l0),
    ]
    m = 7 * l0
    return _assemble("fsi2d", p, rc, l0, (-m, -m, 0.0), (Lx + m, Ly + 10 * l0 + m, l0), cubs)


def fsi3d(l0: float, tank=(1.6, 0.8, 1.0), water=(0.6, 0.5, 1.0), plate_x: float = 0.9,
          plate_layers: int = 4, plate_h: float = 0.3, elastic_dt_ratio: int = 5, dt: float = 1.0e-4,
          z_walls: bool = True, _count_only: bool = False):
    """C4/C5: 3D dam break on an elastic plate.  x = flow/slab axis, y = up (gravity -y), z = span.

    The plate root is embedded three layers into the floor so that the DAM_Module clamp
    (y0 < 0.002, src/main.cpp:1968) holds it for any spacing.
    """
    p, rc = default_params(3, abi.MODULE_DAM)
    p.dt = dt
    p.elastic_dt = dt / elastic_dt_ratio
    w = 3 * l0
    Lx, Ly, Lz = tank
    pt = plate_layers * l0
    zlo, zhi = (0.0, Lz)
    cubs = [
        Cuboid(1, (0.0, 0.0, zlo), (water[0], water[1], min(water[2], Lz)), l0),
        Cuboid(2, (plate_x, -w, zlo), (plate_x + pt, plate_h, zhi), l0),
        # floor, split around the embedded plate root
        Cuboid(4, (-w, -w, zlo), (plate_x, 0.0, zhi), l0),
        Cuboid(4, (plate_x + pt, -w, zlo), (Lx + w, 0.0, zhi), l0),
        Cuboid(4, (-w, 0.0, zlo), (0.0, Ly, zhi), l0),
        Cuboid(4, (Lx, 0.0, zlo), (Lx + w, Ly, zhi), l0),
    ]
    m = 7 * l0
    if _count_only:
        if z_walls:
            cubs.append(Cuboid(4, (-w, -w, -w), (Lx + w, Ly, 0.0), l0))
            cubs.append(Cuboid(4, (-w, -w, Lz), (Lx + w, Ly, Lz + w), l0))
        return cubs
    if z_walls:
        cubs.append(Cuboid(4, (-w, -w, -w), (Lx + w, Ly, 0.0), l0))
        cubs.append(Cuboid(4, (-w, -w, Lz), (Lx + w, Ly, Lz + w), l0))
        dom_lo, dom_hi = (-m, -m, -m), (Lx + m, Ly + m, Lz + m)
    else:  # periodic span
        dom_lo, dom_hi = (-m, -m, 0.0), (Lx + m, Ly + m, Lz)
    return _assemble("fsi3d", p, rc, l0, dom_lo, dom_hi, cubs)


def fsi3d_mini() -> Case:
    """small 3D FSI case for parity tests (about 15k particles)."""
    l0 = 4.0e-3
    c = fsi3d(l0, tank=(0.20, 0.10, 0.048), water=(0.06, 0.08, 0.048), plate_x=0.10,
              plate_layers=3, plate_h=0.04, elastic_dt_ratio=5, z_walls=False)
    c.name = "fsi3d_mini"
    return c


def tiny2d() -> Case:
    """about 1.3k particles: 2D water column + clamped plate + walls (golden-fixture case)"""
    l0 = 2.0e-3
    c = fsi2d(l0, tank=(0.10, 0.06), water=(0.03, 0.04), plate_x=0.05, plate_t=3 * l0, plate_h=0.024,
              elastic_dt=2.0e-5)
    c.name = "tiny2d"
    return c


def tiny3d() -> Case:
    """about 2.5k particles: 3D water block + clamped plate + walls, periodic span (golden-fixture case)"""
    l0 = 4.0e-3
    c = fsi3d(l0, tank=(0.12, 0.06, 0.032), water=(0.036, 0.044, 0.032), plate_x=0.06, plate_layers=3,
              plate_h=0.028, elastic_dt_ratio=5, z_walls=False)
    c.name = "tiny3d"
    return c


def module_case(module: str) -> Case:
    """small 2D FSI case for one of the reference's other compile-time variants (src/main.cpp:56-59): a water block next
    to a plate standing on a floor, placed so that the variant's clamp condition (updateElasticPosition :1910-2082) holds
    for part of the plate.  module: turek | rolling1 | hydro | rolling2 | rollwall (Bar_Module + `#define Rolling`)."""
    l0 = 2.0e-3
    mod = dict(turek=abi.MODULE_TUREK_HRON, rolling1=abi.MODULE_ROLLING1, hydro=abi.MODULE_HYDROELASTIC,
               rolling2=abi.MODULE_ROLLING2, rollwall=abi.MODULE_BAR)[module]
    px, py = dict(turek=(0.201, 0.0), rolling1=(0.05, 0.0), hydro=(0.005, 0.0), rolling2=(0.05, 0.33), rollwall=(0.05, 0.0))[module]
    p, rc = default_params(2, mod)
    p.elastic_dt = 2.0e-5
    w = 3 * l0
    cubs = [Cuboid(1, (px - 0.022, py, 0.0), (px - l0, py + 0.012, l0), l0),
            Cuboid(2, (px, py, 0.0), (px + w, py + 0.02, l0), l0),
            Cuboid(4, (px - 0.03, py - w, 0.0), (px + 0.03, py, l0), l0)]
    if module == "rollwall":
        p.wall_module = abi.WALL_ROLLING
        p.wall_center[4][0], p.wall_center[4][1] = px, py
    m = 7 * l0
    c = _assemble("module_" + module, p, rc, l0, (px - 0.03 - m, py - w - m, 0.0), (px + 0.03 + m, py + 0.04, l0), cubs)
    return c


def fsi3d_for_count(n_target: float, **kw) -> Case:
    """scale l0 so that the default 3D geometry has about n_target particles (walls scale with the
    surface, so the spacing is found by a few fixed-point iterations on coarse counts)"""
    def count(l0):
        n = 0
        for cub in fsi3d(l0, **kw, _count_only=True):
            n += int(np.prod([len(_axis(cub.lower[d], cub.upper[d], cub.spacing)) for d in range(3)]))
        return n
    l0 = 0.02
    for _ in range(6):
        l0 = float("%.4e" % (l0 * (count(l0) / float(n_target)) ** (1.0 / 3.0)))
    c = fsi3d(l0, **kw)
    c.name = "fsi3d_%dk" % round(c.n / 1000)
    return c


# ---- file formats (Python side: only for building inputs / reading results in tests) ------------
def write_data_file(fn: str, p: abi.Params, rc: abi.RunControl):
    """keys of src/main.cpp:743-767"""
    def row(key, vals):
        return key + " " + " ".join("%.17g" % v for v in vals) + "\n"
    with open(fn, "w") as f:
        f.write("#######\n")
        f.write(row("Dt", [p.dt]))
        f.write(row("ElasticDt", [p.elastic_dt]))
        f.write(row("OutputInterval", [rc.output_interval]))
        f.write(row("VtkOutputInterval", [rc.vtk_output_interval]))
        f.write(row("EndTime", [rc.end_time]))
        f.write(row("RadiusRatioA", [p.radius_ratio_a]))
        f.write(row("RadiusRatioP", [p.radius_ratio_p]))
        f.write(row("RadiusRatioV", [p.radius_ratio_v]))
        f.write(row("Density", list(p.density)))
        f.write(row("BulkModulus", list(p.bulk_modulus)))
        f.write(row("BulkViscosity", list(p.bulk_viscosity)))
        f.write(row("ShearViscosity", list(p.shear_viscosity)))
        st = list(p.surface_tension)
        f.write(row("SurfaceTension", [st[0], st[1], st[4], st[5]]))
        f.write(row("YoungModulus", list(p.young_modulus)[2:6]))
        f.write(row("PoissonRatio", list(p.poisson_ratio)[2:6]))
        for i in range(6):
            f.write(row("InteractionRatio(Type%d)" % i, list(p.interaction_ratio[i])))
        f.write(row("Gravity", list(p.gravity)))
        for key, t in (("Wall6", 4), ("Wall7", 5)):
            f.write("%s  Center %s Velocity %s Omega %s\n" % (
                key, " ".join("%.17g" % v for v in p.wall_center[t]),
                " ".join("%.17g" % v for v in p.wall_velocity[t]),
                " ".join("%.17g" % v for v in p.wall_omega[t])))


def write_boid_file(fn: str, case: Case):
    """the pre-processor's input for `case` (generator/generator.cpp:127-262): header + one StartCuboid block per cuboid"""
    p = case.params
    with open(fn, "w") as f:
        f.write("# %s\n" % case.name)
        f.write("ParticleDistance %.17g\n" % p.particle_spacing)
        f.write("LowerDomain %.17g %.17g %.17g\n" % tuple(p.domain_min))
        f.write("UpperDomain %.17g %.17g %.17g\n" % tuple(p.domain_max))
        for cub in case.cuboids:
            f.write("StartCuboid\n    Spacing %.17g\n    Type %d\n    RigidType 10\n    Lower %.17g %.17g %.17g\n"
                    "    Upper %.17g %.17g %.17g\n    Velocity %.17g %.17g %.17g\n    Enthalpy 0.0\nEndCuboid\n"
                    % ((cub.spacing, cub.type) + tuple(cub.lower) + tuple(cub.upper) + tuple(cub.velocity)))


def write_grid_file(fn: str, case: Case):
    """generator/generator.cpp:839-862"""
    p = case.params
    with open(fn, "w") as f:
        f.write("%f\n" % 0.0)
        f.write("%d %e  %e %e %e  %e %e %e\n" % (
            case.n, p.particle_spacing, p.domain_min[0], p.domain_max[0], p.domain_min[1],
            p.domain_max[1], p.domain_min[2], p.domain_max[2]))
        x, x0, v, t = case.position, case.initial_position, case.velocity, case.property
        chunk = 200000
        for s in range(0, case.n, chunk):
            e = min(case.n, s + chunk)
            rows = np.concatenate([x[s:e], x0[s:e], v[s:e]], axis=1)
            lines = ["%d   %e %e %e %e %e %e  %e %e %e \n" % ((int(t[s + i]),) + tuple(rows[i]))
                     for i in range(e - s)]
            f.write("".join(lines))


def read_grid_file(fn: str):
    """.grid/.prof reader (src/main.cpp:788-904): returns time, header tuple, type, x, x0, v"""
    with open(fn) as f:
        time = float(f.readline().split()[0])
        h = f.readline().split()
        n = int(h[0])
        hdr = [float(s) for s in h[1:8]]
        data = np.loadtxt(f, dtype=np.float64, ndmin=2, max_rows=n)
    t = data[:, 0].astype(np.int32)
    return time, hdr, t, np.ascontiguousarray(data[:, 1:4]), np.ascontiguousarray(data[:, 4:7]), \
        np.ascontiguousarray(data[:, 7:10])
