// A kinematic ball pressed into a block of fluid through the CouplingManager hooks (coupling/coupling_manager.rs:9-28,
// LiquidWorld::step_with_coupling liquid_world.rs:67-158) on the C++ host mirror.  The manager below is the ball case of
// the rapier integration's ColliderSampling::DynamicContactSampling (integrations/rapier/fluids_pipeline.rs:192-255): every
// step the fluid particles near the collider are found with particles_intersecting_aabb, projected onto its surface (one
// boundary particle per fluid particle, so the boundary's particle COUNT changes every step), penetrating particles are
// pushed out, and after the solve the boundary forces are summed for the body.  Also uses non-default solver kernels
// (DFSPHSolver<Poly6Kernel, SpikyKernel>, dfsph_solver.rs:17-20) when built against libsalva_b200_kernels.so with -DUSE_KERNELS.
//   g++ -std=c++17 -Iinclude examples/coupling3.cpp -Lsalva_b200 -lsalva_b200 -Wl,-rpath,$PWD/salva_b200 -o coupling3
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "salva3d_b200.hpp"

using namespace salva3d;

struct BallCoupling : CouplingManager {
    FluidHandle fluid;
    BoundaryHandle boundary;
    Point3 center{0.5f, 0.66f, 0.4f};
    Vector3 velocity{0.f, -2.f, 0.f};
    Real radius = 0.12f, step_dt = 0.004f;
    Vector3 total_force;
    size_t last_count = 0;

    void update_boundaries(LiquidWorld& world, Real dt, Real, Real h, Real particle_radius) override {
        center.y += velocity.y * step_dt;
        const Real prediction = 0.5f * h, margin = 0.1f * particle_radius, ext = radius + h + prediction;
        Fluid& f = world.fluids_mut()[fluid];
        Boundary& b = world.boundaries_mut()[boundary];
        b.positions.clear();
        b.velocities.clear();
        for (const auto& id : world.particles_intersecting_aabb({center.x - ext, center.y - ext, center.z - ext}, {center.x + ext, center.y + ext, center.z + ext})) {
            if (id.is_boundary) continue;  // never happens here: only the fluids are in the grid at this point (liquid_world.rs:86-103)
            Point3& p = f.positions[id.index];
            Vector3& v = f.velocities[id.index];
            const Real px = p.x + v.x * dt - center.x, py = p.y + v.y * dt - center.y, pz = p.z + v.z * dt - center.z;
            const Real dist = std::sqrt(px * px + py * py + pz * pz);
            if (!(dist > 1.0e-7f)) continue;
            const Real nx = px / dist, ny = py / dist, nz = pz / dist;
            if (dist < radius) {  // proj.is_inside: push out along the normal and remove the approaching velocity
                const Real push = radius - dist + margin;
                p.x += nx * push; p.y += ny * push; p.z += nz * push;
                const Real vn = nx * v.x + ny * v.y + nz * v.z;
                if (vn < 0.f) { v.x -= nx * vn; v.y -= ny * vn; v.z -= nz * vn; }
            } else if (dist - radius > h + prediction) {
                continue;
            }
            b.positions.push_back({center.x + nx * radius, center.y + ny * radius, center.z + nz * radius});
            b.velocities.push_back(velocity);
        }
        last_count = b.positions.size();
    }
    void transmit_forces(LiquidWorld& world, Real, Real) override {
        total_force = Vector3();
        for (const Vector3& f : world.boundaries()[boundary].forces) { total_force.x += f.x; total_force.y += f.y; total_force.z += f.z; }
    }
};

int main(int argc, char** argv) {
    const int steps = argc > 1 ? std::atoi(argv[1]) : 25;
    const Real r = 0.05f;
    try {
#ifdef USE_KERNELS
        LiquidWorld world(DFSPHSolver<Poly6Kernel, SpikyKernel>(), r, 2.0f);
#else
        LiquidWorld world(DFSPHSolver<>(), r, 2.0f);
#endif
        std::vector<Point3> pts;
        for (int i = 0; i < 10; ++i)
            for (int j = 0; j < 6; ++j)
                for (int k = 0; k < 8; ++k) pts.push_back({(2 * i + 1) * r * 0.93f, (2 * j + 1) * r * 0.93f, (2 * k + 1) * r * 0.93f});
        std::vector<Point3> tank;
        for (int i = -1; i <= 10; ++i)
            for (int k = -1; k <= 8; ++k) tank.push_back({(2 * i + 1) * r, -r, (2 * k + 1) * r});
        BallCoupling cm;
        cm.fluid = world.add_fluid(Fluid(pts, r, 1000.0f));
        world.add_boundary(Boundary(tank));
        cm.boundary = world.add_boundary(Boundary({}, InteractionGroups(), true));
        size_t min_count = ~size_t(0), max_count = 0;
        for (int s = 0; s < steps; ++s) {
            world.step_with_coupling(cm.step_dt, {0.f, -9.81f, 0.f}, &cm);
            min_count = std::min(min_count, cm.last_count);
            max_count = std::max(max_count, cm.last_count);
        }
        std::printf("coupling3: %zu particles, %d steps, sampled boundary %zu..%zu particles, force on the ball (%.3f, %.3f, %.3f)\n",
                    world.fluids()[cm.fluid].num_particles(), steps, min_count, max_count, cm.total_force.x, cm.total_force.y, cm.total_force.z);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "coupling3: %s\n", e.what());
        return 2;
    }
    return 0;
}
