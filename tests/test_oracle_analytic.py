"""Pins the CPU oracle against analytic known answers derived from the reference's formulas
(SURVEY.md Appendix B).  The reference ships no golden vectors for this path (PARITY UNPINNED
upstream), so these closed-form values are the first anchor of the oracle."""
import numpy as np
import pytest

from oracle.oracle import OracleWorld, lib
from salva_b200 import scenes


def test_kernel_w0_and_normalisation():
    L = lib()
    for h in (0.2, 0.1):
        w0 = L.orc_kernel_w(0.0, h)
        assert w0 == pytest.approx(8.0 / (np.pi * h ** 3), rel=1e-6)  # W(0) = sigma (Appendix B: 318.3099 / 2546.479)
        # integral of W over the ball of radius h equals 1 (cubic_spline_kernel.rs normaliser 8/(pi h^3))
        r = (np.arange(20000) + 0.5) * (h / 20000)
        w = np.array([L.orc_kernel_w(float(x), h) for x in r[::20]])
        integral = np.sum(4 * np.pi * r[::20] ** 2 * w) * (h / 1000)
        assert integral == pytest.approx(1.0, rel=2e-3)
    assert L.orc_kernel_w(0.0, 0.2) == pytest.approx(318.3099, rel=1e-6)
    assert L.orc_kernel_w(0.0, 0.1) == pytest.approx(2546.479, rel=1e-6)


def test_kernel_derivative_matches_finite_difference():
    L = lib()
    h = 0.2
    for r in (0.01, 0.05, 0.099, 0.101, 0.15, 0.19):
        fd = (L.orc_kernel_w(r + 1e-4, h) - L.orc_kernel_w(r - 1e-4, h)) / 2e-4
        assert L.orc_kernel_dw(r, h) == pytest.approx(fd, rel=2e-2, abs=1e-1)
    assert L.orc_kernel_dw(0.0, h) == 0.0          # q <= 1e-5  -> 0
    assert L.orc_kernel_dw(0.2000001 * 1.01, h) == 0.0  # q > 1 -> 0


@pytest.mark.parametrize("r,alpha_expected", [(0.05, 3.677856e-8), (0.025, 9.194639e-9)])
def test_rest_lattice_interior_particle(r, alpha_expected):
    """Interior particle of an unjittered cubic lattice: rho = 0.79998 rho0, 33 contacts incl. self
    (27 strictly inside + 6 at exactly d == h, up to rounding), alpha as tabulated in Appendix B."""
    n = 9
    pts = scenes.block_lattice(n, n, n, r)
    w = OracleWorld(r, 2.0)
    f = w.add_fluid(pts, density0=1000.0)
    w.force_iterations(0, 0)
    w.step(1.0 / 200.0, gravity=(0.0, 0.0, 0.0))
    centre = (n // 2) * n * n + (n // 2) * n + (n // 2)
    dens = w.debug(f, "density")
    alpha = w.debug(f, "alpha")
    cnt = w.debug(f, "num_fluid_contacts")
    assert dens[centre] == pytest.approx(799.978, rel=2e-5)
    assert 27 <= cnt[centre] <= 33  # the six d == h pairs are accepted or not by f32 rounding
    assert alpha[centre] == pytest.approx(alpha_expected, rel=2e-4)


def test_momentum_conservation_of_pressure_solve():
    """The fluid-fluid pair terms of compute_velocity_changes are antisymmetric (dfsph_solver.rs:237-255):
    sum_i m_i * vc_i is unchanged by the pressure loop when there are no boundaries and no gravity."""
    r = 0.05
    rng = np.random.default_rng(3)
    pts = scenes.jitter(scenes.block_lattice(8, 8, 8, r * 0.9), r, 7)  # compressed lattice => positive pressure
    vel = rng.normal(0, 0.05, pts.shape).astype(np.float32)
    w = OracleWorld(r, 2.0)
    f = w.add_fluid(pts, density0=1000.0, velocities=vel)
    w.force_iterations(2, 3)
    w.step(1.0 / 200.0, gravity=(0.0, 0.0, 0.0))
    p, v = w.read_fluid(f)
    vc = w.debug(f, "velocity_change")
    mom0 = vel.astype(np.float64).sum(axis=0)
    mom1 = (v.astype(np.float64) + vc.astype(np.float64)).sum(axis=0)
    assert np.abs(vc).max() > 1e-4  # the solve did something
    assert np.allclose(mom0, mom1, atol=2e-3 * np.abs(v).sum())


def test_first_step_has_zero_inv_dt():
    """timestep_manager.rs:29-30 + dfsph_solver.rs:702: dt/inv_dt are 0 until advance() mid-step, so XSPH
    contributes nothing on the very first step (xsph_viscosity.rs:92-93)."""
    r = 0.05
    pts = scenes.jitter(scenes.block_lattice(6, 6, 6, r), r, 11)
    vel = np.random.default_rng(0).normal(0, 0.1, pts.shape).astype(np.float32)
    a = OracleWorld(r, 2.0)
    fa = a.add_fluid(pts, velocities=vel)
    a.push_force(fa, *scenes.xsph_viscosity(0.5, 0.0))
    b = OracleWorld(r, 2.0)
    fb = b.add_fluid(pts, velocities=vel)
    for w in (a, b):
        w.force_iterations(1, 1)
        w.step(0.005)
    assert np.array_equal(a.debug(fa, "acceleration"), b.debug(fb, "acceleration"))
    a.step(0.005)
    b.step(0.005)
    assert not np.array_equal(a.debug(fa, "acceleration"), b.debug(fb, "acceleration"))


def test_host_force_callback_adds_to_accelerations():
    """User-defined NonPressureForce (nonpressure_force.rs:10-30) on the oracle: a callback that cancels gravity leaves an
    isolated particle at rest."""
    w = OracleWorld(0.05, 2.0)
    f = w.add_fluid(np.array([[0.0, 1.0, 0.0], [5.0, 1.0, 0.0]], np.float32))
    seen = []

    def solve(dt, inv_dt, h, pos, vel, dens, acc):
        seen.append((dt, len(pos)))
        acc[:, 1] += 9.81

    w.push_host_force(f, solve)
    for _ in range(3):
        w.step(0.01)
    p, v = w.read_fluid(f)
    assert len(seen) == 3 and seen[0] == (0.0, 2) and seen[1][0] == pytest.approx(0.01)   # dt lags one step (dfsph_solver.rs:702)
    assert np.abs(p - np.array([[0.0, 1.0, 0.0], [5.0, 1.0, 0.0]])).max() < 1e-6


def _aabb_brute_force(h, r, mins, maxs, groups):
    """liquid_world.rs:211-243 restated densely: groups = [(kind, handle, positions_at_last_grid_build, positions_now)]."""
    f32 = np.float32
    lo = np.floor(np.asarray(mins, f32) / f32(h))
    hi = np.floor(np.asarray(maxs, f32) / f32(h))
    out = []
    for kind, handle, p_grid, p_now in groups:
        cell = np.floor(p_grid.astype(f32) / f32(h))
        in_cells = np.all((cell >= lo) & (cell <= hi), axis=1)
        ex = np.maximum(np.maximum(np.asarray(mins, f32) - p_now, p_now - np.asarray(maxs, f32)), f32(0))
        near = np.sqrt((ex.astype(f32) ** 2).sum(axis=1, dtype=f32)) < f32(r)
        out += [(kind, handle, int(i)) for i in np.nonzero(in_cells & near)[0]]
    return sorted(out)


def test_particles_intersecting_aabb_matches_brute_force():
    r = 0.05
    w = OracleWorld(r, 2.0)
    pts = scenes.jitter(scenes.block_lattice(8, 7, 6, r), r, 4, amplitude=0.3)
    floor = scenes.open_tank((-r, -r, -r), (8 * 2 * r + r, 0.4, 6 * 2 * r + r), r)
    f = w.add_fluid(pts, velocities=np.full_like(pts, 0.8))
    b = w.add_boundary(floor)
    mins, maxs = (0.13, -0.2, 0.11), (0.47, 0.33, 0.38)
    k, hd, ix = w.particles_intersecting_aabb(mins, maxs)
    assert len(k) == 0                                  # no step yet: the grid is empty (liquid_world.rs:90-117)
    w.step(0.004)
    before, _ = w.read_fluid(f)
    w.step(0.004)                                       # grid holds `before`, positions have moved on
    now, _ = w.read_fluid(f)
    k, hd, ix = w.particles_intersecting_aabb(mins, maxs)
    got = sorted(zip(k.tolist(), hd.tolist(), ix.tolist()))
    want = _aabb_brute_force(w.h, r, mins, maxs, [(0, f, before, now), (1, b, floor, floor)])
    assert got == want
    assert 20 < sum(1 for e in got if e[0] == 0) < len(pts) and any(e[0] == 1 for e in got)


def test_non_default_kernels_are_normalised_and_consistent():
    """kernel/poly6_kernel.rs, spiky_kernel.rs, viscosity_kernel.rs (dim3): each integrates to 1 over its support and
    scalar_apply_diff is the derivative of scalar_apply (the analytic pin of the oracle's restatement of them)."""
    from oracle.oracle import lib
    L = lib()
    h = 0.2
    r = np.linspace(1e-4, h, 4001)
    for kind in (1, 2, 3):
        W = np.array([L.orc_kernel_w_kind(kind, float(x), h) for x in r], np.float64)
        dW = np.array([L.orc_kernel_dw_kind(kind, float(x), h) for x in r], np.float64)
        assert abs(np.trapezoid(4 * np.pi * r * r * W, r) - 1.0) < 2e-3
        num = np.gradient(W, r)
        assert np.abs(dW - num)[5:-5].max() <= 2e-2 * np.abs(dW).max()
        assert L.orc_kernel_w_kind(kind, 1.01 * h, h) == 0.0 and L.orc_kernel_dw_kind(kind, 1.01 * h, h) == 0.0
    assert L.orc_kernel_w_kind(3, 0.0, h) == 0.0      # viscosity kernel: `r > 0` guard (viscosity_kernel.rs:24)
