"""GPU tests of the LiquidWorld / plugin surface added in round 2, all through the C ABI:
snapshot / restore (bit-exact), zero-copy position views, remove_fluid / remove_boundary, host plugins that walk the
materialised ParticlesContacts, CouplingManager hooks with a re-sampled boundary, particles_intersecting_shape."""
import numpy as np
import pytest

from oracle.oracle import OracleWorld
from salva_b200 import DFSPHSolver, IISPHSolver, LiquidWorld, scenes
from salva_b200.liquid_world import Ball, Capsule, CouplingManager, Cuboid

pytestmark = pytest.mark.gpu


def _scene(seed=3, forces=(), nx=10, ny=9, nz=8, compress=0.93, two_fluids=False):
    r = 0.05
    rng = np.random.default_rng(seed)
    pts = scenes.jitter(scenes.block_lattice(nx, ny, nz, r * compress), r, seed, amplitude=0.3)
    vel = rng.normal(0, 0.2, pts.shape).astype(np.float32)
    tank = scenes.open_tank((-r, -r, -r), (nx * 2 * r + r, 1.2, nz * 2 * r + r), r)
    fluids = [dict(positions=pts, velocities=vel, density0=1000.0, forces=list(forces))]
    if two_fluids:
        up = scenes.jitter(scenes.block_lattice(nx, 4, nz, r * compress, origin=(0.0, ny * 2 * r * compress, 0.0)), r, seed + 1, amplitude=0.3)
        fluids.append(dict(positions=up, velocities=rng.normal(0, 0.2, up.shape).astype(np.float32), density0=800.0, forces=list(forces)))
    return dict(particle_radius=r, smoothing_factor=2.0, dt=0.004, gravity=scenes.GRAVITY, fluids=fluids, boundaries=[dict(positions=tank)])


def _world(sc, solver=None):
    w = LiquidWorld(solver or DFSPHSolver(), particle_radius=sc["particle_radius"])
    fh, bh = scenes.populate(w, sc)
    return w, fh, bh


@pytest.mark.parametrize("case", ["dfsph_xsph", "iisph_2fluids", "becker"])
def test_snapshot_restore_is_bit_exact(case):
    """step 5, save, step 5 -> A; fresh world, load, step 5 -> B; A == B bit for bit (vc carry-over dfsph_solver.rs:704-706,
    dt lag timestep_manager.rs:29-30, IISPH warm start iisph_solver.rs:673-677, Becker rest pose becker2009_elasticity.rs:84-135)."""
    if case == "dfsph_xsph":
        sc, solver = _scene(5, forces=(scenes.xsph_viscosity(0.5, 0.2),)), DFSPHSolver
    elif case == "iisph_2fluids":
        sc, solver = _scene(7, forces=(scenes.artificial_viscosity(1.0, 0.0),), two_fluids=True), IISPHSolver
    else:
        sc, solver = _scene(9, forces=(scenes.becker2009_elasticity(1.0e5, 0.3, True),), nx=8, ny=8, nz=8, compress=1.0), DFSPHSolver
    a, fa, _ = _world(sc, solver())
    for _ in range(5):
        a.step(sc["dt"])
    blob = a.snapshot()
    for _ in range(5):
        a.step(sc["dt"])
    # an undisturbed run: taking the snapshot must not perturb the trajectory either (canonical in-cell order)
    u, fu, _ = _world(sc, solver())
    for _ in range(10):
        u.step(sc["dt"])
    b, fb, _ = _world(sc, solver())
    b.restore(blob)
    for _ in range(5):
        b.step(sc["dt"])
    for k in range(len(fa)):
        pa, va = a.read_fluid(fa[k])
        pb, vb = b.read_fluid(fb[k])
        pu, vu = u.read_fluid(fu[k])
        assert np.array_equal(pa, pb) and np.array_equal(va, vb)
        assert np.array_equal(a.debug(fa[k], "velocity_change"), b.debug(fb[k], "velocity_change"))
        assert np.array_equal(pa, pu) and np.array_equal(va, vu)
    with pytest.raises(Exception):
        b.restore(blob[:100])


def test_map_positions_is_a_device_view_in_original_order():
    import torch
    sc = _scene(11)
    w, fh, _ = _world(sc)
    for _ in range(3):
        w.step(sc["dt"])
    p, v = w.read_fluid(fh[0])
    tp = torch.as_tensor(w.map_positions(fh[0]), device="cuda")
    assert tp.is_cuda and tuple(tp.shape) == p.shape
    assert np.array_equal(tp.cpu().numpy(), p)
    tv = torch.as_tensor(w.map_positions(fh[0], velocities=True), device="cuda")
    assert np.array_equal(tv.cpu().numpy(), v)


def test_remove_fluid_and_boundary_match_a_world_built_without_them():
    """LiquidWorld::remove_fluid / remove_boundary (liquid_world.rs:171-178): handles of the survivors stay valid and the
    world continues exactly like the oracle's world that never had the removed objects."""
    sc = _scene(13, forces=(scenes.xsph_viscosity(0.5, 0.0),), two_fluids=True)
    r = sc["particle_radius"]
    lid = scenes._face(1, 1.1, (0.0, 0.0, 0.0), (0.5, 0.0, 0.4), 2 * r)
    gpu = LiquidWorld(particle_radius=r)
    fh, bh = scenes.populate(gpu, sc)
    extra = gpu.add_boundary(lid)
    gpu.force_iterations(1, 2)
    gpu.remove_fluid(fh[1])
    gpu.remove_boundary(extra)
    with pytest.raises(Exception):
        gpu.read_fluid(fh[1])
    cpu = OracleWorld(r, 2.0)
    only = dict(sc)
    only["fluids"] = sc["fluids"][:1]
    fc, _ = scenes.populate(cpu, only)
    cpu.force_iterations(1, 2)
    for _ in range(3):
        gpu.step(sc["dt"])
        cpu.step(sc["dt"])
    pg, _ = gpu.read_fluid(fh[0])
    pc, _ = cpu.read_fluid(fc[0])
    assert np.abs(pg - pc).max() <= 1e-3 * float(gpu.h)
    # the freed slot is reused under a new generation; the stale handle stays dead
    again = gpu.add_fluid(sc["fluids"][1]["positions"], density0=800.0)
    assert again != fh[1] and (again & 0xFFFF) == (fh[1] & 0xFFFF)
    gpu.step(sc["dt"])
    assert gpu.num_particles(again) == len(sc["fluids"][1]["positions"])
    with pytest.raises(Exception):
        gpu.num_particles(fh[1])
    # removal after steps (device state -> host -> erase) keeps the survivors' trajectories
    gpu.remove_fluid(again)
    gpu.step(sc["dt"])
    cpu.step(sc["dt"])
    cpu.step(sc["dt"])


def test_host_plugin_recomputes_xsph_from_materialised_contacts():
    """A NonPressureForce plugin that walks fluid_fluid_contacts / fluid_boundaries_contacts (nonpressure_force.rs:15-27)
    exactly like XSPHViscosity::solve (xsph_viscosity.rs:30-95) must reproduce the built-in force."""
    coeff_f, coeff_b = 0.5, 0.3

    def solve(ctx):
        ff, fb = ctx.fluid_fluid_contacts, ctx.fluid_boundaries_contacts
        n = len(ctx.positions)
        i_of = np.repeat(np.arange(n), np.diff(ff.offsets.astype(np.int64)))
        same = ff.j_model == ctx.fluid_index
        vol = ctx.volumes
        c = coeff_f * ff.weight * vol[ff.j] * ctx.density0 / ctx.densities[ff.j]
        dv = ctx.velocities[ff.j] - ctx.velocities[i_of]
        acc = np.zeros((n, 3), np.float64)
        np.add.at(acc, i_of[same], (c[:, None] * dv)[same] * ctx.inv_dt)
        ib = np.repeat(np.arange(n), np.diff(fb.offsets.astype(np.int64)))
        for b, bd in enumerate(ctx.boundaries):
            m = fb.j_model == b
            if not m.any():
                continue
            cb = coeff_b * fb.weight[m] * bd["volumes"][fb.j[m]] * ctx.density0 / ctx.densities[ib[m]]
            np.add.at(acc, ib[m], cb[:, None] * (bd["velocities"][fb.j[m]] - ctx.velocities[ib[m]]) * ctx.inv_dt)
        ctx.accelerations += acc.astype(np.float32)

    sc = _scene(17)
    tank = sc["boundaries"][0]
    tank["velocities"] = np.tile(np.array([0.3, 0.0, -0.2], np.float32), (len(tank["positions"]), 1))
    ref, fr, _ = _world(sc)
    ref.push_force(fr[0], *scenes.xsph_viscosity(coeff_f, coeff_b))
    plug, fp, _ = _world(sc)
    plug.push_host_force2(fp[0], solve)
    for w in (ref, plug):
        w.force_iterations(1, 2)
    for _ in range(3):
        ref.step(sc["dt"])
        plug.step(sc["dt"])
    ar, ap = ref.debug(fr[0], "acceleration"), plug.debug(fp[0], "acceleration")
    assert np.abs(ar - ap).max() <= 1e-5 * np.abs(ar).max()
    pr, _ = ref.read_fluid(fr[0])
    pp, _ = plug.read_fluid(fp[0])
    assert np.abs(pr - pp).max() <= 1e-5 * float(ref.h)


def test_materialised_contacts_match_the_oracle_lists():
    seen = {}

    def solve(ctx):
        seen["ff_counts"] = np.diff(ctx.fluid_fluid_contacts.offsets.astype(np.int64))
        seen["fb_counts"] = np.diff(ctx.fluid_boundaries_contacts.offsets.astype(np.int64))
        ff = ctx.fluid_fluid_contacts
        i_of = np.repeat(np.arange(len(ctx.positions)), seen["ff_counts"])
        d = ctx.positions[i_of] - ctx.positions[ff.j]
        seen["self"] = int(((ff.j == i_of) & (ff.j_model == ctx.fluid_index)).sum())
        seen["max_d"] = float(np.sqrt((d * d).sum(1)).max())
        seen["grad_on_self"] = float(np.abs(ff.gradient[ff.j == i_of]).max())

    sc = _scene(19)
    gpu, fg, _ = _world(sc)
    gpu.push_host_force2(fg[0], solve)
    cpu = OracleWorld(sc["particle_radius"], 2.0)
    fc, _ = scenes.populate(cpu, sc)
    gpu.step(sc["dt"])
    cpu.step(sc["dt"])
    assert np.array_equal(seen["ff_counts"], cpu.debug(fc[0], "num_fluid_contacts").astype(np.int64))   # exact, self included
    assert np.array_equal(seen["fb_counts"], cpu.debug(fc[0], "num_boundary_contacts").astype(np.int64))
    assert seen["self"] == len(sc["fluids"][0]["positions"]) and seen["grad_on_self"] == 0.0
    assert seen["max_d"] <= float(gpu.h) * (1 + 1e-6)


class _BallCoupling(CouplingManager):
    """DynamicContactSampling of a kinematic ball (fluids_pipeline.rs:192-255): fluid particles near the collider are
    projected onto its surface (one boundary particle each, count changes every step), penetrating ones are pushed out."""

    def __init__(self, boundary, fluid, radius):
        self.boundary, self.fluid, self.radius = boundary, fluid, radius
        self.center = np.array([0.5, 0.66, 0.4], np.float32)
        self.velocity = np.array([0.0, -2.0, 0.0], np.float32)
        self.counts, self.forces = [], []

    def update_boundaries(self, world, dt, inv_dt, h, particle_radius):
        self.center = (self.center + self.velocity * np.float32(0.004)).astype(np.float32)
        prediction, margin = 0.5 * h, 0.1 * particle_radius
        ext = self.radius + h + prediction
        kinds, handles, idx = world.particles_intersecting_aabb(self.center - ext, self.center + ext)
        assert (kinds == 0).all(), "only fluid particles are in the grid during update_boundaries (liquid_world.rs:86-103)"
        pos, vel = world.read_fluid(self.fluid)
        sel = idx[handles == self.fluid]
        pp = pos[sel] + vel[sel] * dt
        d = pp - self.center
        dist = np.sqrt((d * d).sum(1))
        n = d / np.maximum(dist, 1e-12)[:, None]
        proj = self.center + n * self.radius
        inside = dist < self.radius
        keep = inside | (dist - self.radius <= h + prediction)
        depth = self.radius - dist
        pos[sel[inside]] += n[inside] * (depth[inside] + margin)[:, None]
        vn = (n[inside] * vel[sel[inside]]).sum(1)
        vel[sel[inside]] -= n[inside] * np.minimum(vn, 0.0)[:, None]
        if inside.any():
            world.write_fluid(self.fluid, pos, vel)
        world.set_boundary_particles(self.boundary, proj[keep], np.tile(self.velocity, (int(keep.sum()), 1)))
        self.counts.append(int(keep.sum()))

    def transmit_forces(self, world, dt, inv_dt):
        _, f = world.read_boundary(self.boundary)
        self.forces.append(f.sum(0))


def test_coupling_manager_with_resampled_boundary():
    sc = _scene(23, nx=10, ny=6, nz=8)
    w, fh, _ = _world(sc)
    ball = w.add_boundary(np.zeros((0, 3), np.float32), want_forces=True)
    cm = _BallCoupling(ball, fh[0], 0.12)
    n0 = w.num_particles(fh[0])
    for _ in range(25):
        w.step_with_coupling(sc["dt"], scenes.GRAVITY, cm)
    assert max(cm.counts) > 0 and len(set(cm.counts)) > 1, "the sampled boundary must change size from step to step"
    assert w.num_particles(fh[0]) == n0
    pos, _ = w.read_fluid(fh[0])
    assert np.isfinite(pos).all()
    d = np.sqrt(((pos - cm.center) ** 2).sum(1))
    assert d.min() >= 0.12 - 0.06, "particles must be kept out of the ball"
    late = np.array(cm.forces[-8:])
    assert np.abs(late).max() > 0, "the fluid pushes back on the descending ball"


def test_particles_intersecting_shape_matches_brute_force():
    """liquid_world.rs:246-281: cells of the posed shape's AABB, distance_to_point(solid) <= particle_radius."""
    sc = _scene(29)
    w, fh, bh = _world(sc)
    w.step(sc["dt"])
    pos, _ = w.read_fluid(fh[0])
    tank = sc["boundaries"][0]["positions"]
    h, r = float(w.h), sc["particle_radius"]
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th), 0.0], [np.sin(th), np.cos(th), 0.0], [0.0, 0.0, 1.0]], np.float32)
    t = np.array([0.45, 0.3, 0.35], np.float32)

    def brute(kind, params, pts):
        loc = (pts - t) @ R  # R^T (p - t)
        if kind == 1:
            dist = np.maximum(np.sqrt((loc * loc).sum(1)) - params[0], 0.0)
            ext = np.full(3, params[0])
        elif kind == 2:
            e = np.maximum(np.abs(loc) - np.array(params), 0.0)
            dist = np.sqrt((e * e).sum(1))
            ext = np.abs(R) @ np.array(params)
        else:
            cy = np.clip(loc[:, 1], -params[0], params[0])
            q = loc.copy()
            q[:, 1] -= cy
            dist = np.maximum(np.sqrt((q * q).sum(1)) - params[1], 0.0)
            ext = np.abs(R[:, 1]) * params[0] + params[1]
        lo, hi = np.floor((t - ext) / np.float32(h)), np.floor((t + ext) / np.float32(h))
        cell = np.floor(pts / np.float32(h))
        in_cells = ((cell >= lo) & (cell <= hi)).all(1)
        return np.nonzero(in_cells & (dist <= r))[0]

    for shape in (Ball(0.17), Cuboid((0.2, 0.08, 0.15)), Capsule(0.15, 0.06)):
        kinds, handles, idx = w.particles_intersecting_shape(shape, t, R)
        want_f, want_b = brute(shape.kind, shape.params, pos), brute(shape.kind, shape.params, tank)
        got_f, got_b = np.sort(idx[kinds == 0]), np.sort(idx[kinds == 1])
        # the distance is evaluated in f32 on the device and in f64-ish numpy here: allow disagreement only within 1e-6 of the threshold
        assert len(want_f) > 10
        assert len(np.setxor1d(got_f, want_f)) <= 2 and len(np.setxor1d(got_b, want_b)) <= 2
        assert (handles[kinds == 0] == fh[0]).all() and (handles[kinds == 1] == bh[0]).all()


def test_aabb_query_accepts_infinite_and_rejects_nan_bounds():
    """hgrid.rs cells_intersecting_aabb returns every cell for an unbounded box; the cast of floor(+-inf / h) must not be
    undefined."""
    sc = _scene(31, nx=6, ny=5, nz=5)
    w, fh, bh = _world(sc)
    w.step(sc["dt"])
    inf = np.float32(np.inf)
    big = np.float32(3.0e38)
    for lo, hi in (((-inf,) * 3, (inf,) * 3), ((-big,) * 3, (big,) * 3)):
        kinds, handles, idx = w.particles_intersecting_aabb(lo, hi)
        assert (kinds == 0).sum() == w.num_particles(fh[0])
        assert (kinds == 1).sum() == len(sc["boundaries"][0]["positions"])
    with pytest.raises(Exception):
        w.particles_intersecting_aabb((np.nan, 0, 0), (1, 1, 1))


@pytest.mark.parametrize("kd,kg,solver", [(1, 2, 0), (2, 2, 0), (3, 1, 0), (1, 2, 1)], ids=["poly6+spiky", "spiky+spiky", "viscosity+poly6", "iisph-poly6+spiky"])
def test_non_default_solver_kernels_match_oracle(kd, kg, solver):
    """DFSPHSolver<KernelDensity, KernelGradient> / IISPHSolver<..> (dfsph_solver.rs:17-20, iisph_solver.rs:17-20)."""
    from salva_b200.liquid_world import CubicSplineKernel, Poly6Kernel, SpikyKernel, ViscosityKernel
    K = {0: CubicSplineKernel, 1: Poly6Kernel, 2: SpikyKernel, 3: ViscosityKernel}
    sc = _scene(37, forces=(scenes.xsph_viscosity(0.5, 0.2),))
    S = DFSPHSolver if solver == 0 else IISPHSolver
    gpu = LiquidWorld(S(K[kd], K[kg]), particle_radius=sc["particle_radius"])
    cpu = OracleWorld(sc["particle_radius"], 2.0, solver=solver, kernel_density=kd, kernel_gradient=kg)
    (fg,), _ = scenes.populate(gpu, sc)
    (fc,), _ = scenes.populate(cpu, sc)
    for w in (gpu, cpu):
        w.force_iterations(2, 3)
    for _ in range(4):
        gpu.step(sc["dt"])
        cpu.step(sc["dt"])
    rg, rc = gpu.debug(fg, "density"), cpu.debug(fc, "density")
    assert np.abs(rg - rc).max() <= 1e-5 * np.abs(rc).max()
    pg, vg = gpu.read_fluid(fg)
    pc, vc = cpu.read_fluid(fc)
    h = float(gpu.h)
    assert np.abs(pg - pc).max() <= 1e-3 * h
    assert np.abs(vg - vc).max() <= 1e-3 * h / sc["dt"]


def test_replace_particles_in_a_different_order_keeps_the_trajectory_bit_exact():
    """sph_fluid_replace_particles (what the slab re-balancing uses): hand the engine the same particles — positions,
    velocities, velocity_changes, ids — in a shuffled index order; the canonical in-cell order (ascending id) makes the
    continued run bit-identical to the undisturbed one, particle by particle."""
    sc = _scene(41, forces=(scenes.xsph_viscosity(0.5, 0.2),))
    a, fa, _ = _world(sc)
    b, fb, _ = _world(sc)
    for _ in range(4):
        a.step(sc["dt"])
        b.step(sc["dt"])
    p, v = b.read_fluid(fb[0])
    vc = b.debug(fb[0], "velocity_change")
    ids = b.read_ids(fb[0])
    perm = np.random.default_rng(0).permutation(len(p))
    b.replace_particles(fb[0], p[perm], v[perm], vc[perm], ids[perm])
    for _ in range(4):
        a.step(sc["dt"])
        b.step(sc["dt"])
    pa, va = a.read_fluid(fa[0])
    pb, vb = b.read_fluid(fb[0])
    ia, ib = a.read_ids(fa[0]), b.read_ids(fb[0])
    assert np.array_equal(ib, ids[perm])
    oa, ob_ = np.argsort(ia), np.argsort(ib)
    assert np.array_equal(pa[oa], pb[ob_]) and np.array_equal(va[oa], vb[ob_])


def test_handle_arena_follows_the_reference_unit_test():
    """The one unit test the reference holds for this boundary: ContiguousArena `smoke` (src/object/contiguous_arena.rs:172-184) —
    three inserts, every handle removes its own value exactly once, a second removal finds nothing.  Fluids and boundaries live in
    that arena (liquid_world.rs:161-178); the value is checked through the particle count stored under the handle."""
    r = 0.05
    w = LiquidWorld(particle_radius=r)
    sizes = (123, 456, 789)
    pts = [scenes.block_lattice(n, 1, 1, r, origin=(0.0, 3.0 * k, 0.0)) for k, n in enumerate(sizes)]
    fh = [w.add_fluid(p, density0=1000.0) for p in pts]
    bh = [w.add_boundary(p + np.float32(20.0)) for p in pts]
    assert len(set(fh)) == 3 and len(set(bh)) == 3
    for h, n in zip(fh, sizes):
        assert w.num_particles(h) == n
        w.remove_fluid(h)                       # Some(value)
        with pytest.raises(Exception):
            w.remove_fluid(h)                   # None
    for h in bh:
        w.remove_boundary(h)
        with pytest.raises(Exception):
            w.remove_boundary(h)
    w.close()
