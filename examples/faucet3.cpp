// examples3d/faucet3.rs (faucet3.rs:19-137) on the C++ host mirror, without the rapier testbed: a fluid that starts EMPTY
// (XSPHViscosity(0.5, 0) + Akinci2013SurfaceTension(1, 10)), a 10 x 10 sheet of particles appended every 0.06 s of simulated
// time (Fluid::add_particles fluid.rs:126-150), falling onto a static ball sampled on its surface, and particles below
// y = -2 deleted at the next timestep (Fluid::delete_particle_at_next_timestep fluid.rs:71-76).  Prints the bookkeeping.
//   g++ -std=c++17 -Iinclude examples/faucet3.cpp -Lsalva_b200 -lsalva_b200 -Wl,-rpath,$PWD/salva_b200 -o faucet3
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "salva3d_b200.hpp"

using namespace salva3d;

// Surface samples of a ball at spacing ~2r (the reference ray-samples the collider: sampling/ray_sampling.rs; any
// deterministic cover of the surface at the particle spacing plays the same role here).
static std::vector<Point3> ball_surface(Real radius, Real particle_rad) {
    const double pi = 3.14159265358979323846;
    const int n = (int)std::ceil(4.0 * pi * radius * radius / (4.0 * particle_rad * particle_rad));
    std::vector<Point3> pts;
    const double golden = pi * (3.0 - std::sqrt(5.0));
    for (int i = 0; i < n; ++i) {
        const double y = 1.0 - 2.0 * (i + 0.5) / n, rho = std::sqrt(1.0 - y * y), th = golden * i;
        pts.push_back({(Real)(radius * rho * std::cos(th)), (Real)(radius * y), (Real)(radius * rho * std::sin(th))});
    }
    return pts;
}

int main(int argc, char** argv) {
    const Real PARTICLE_RADIUS = 0.025f / 2.0f, SMOOTHING_FACTOR = 2.0f, dt = 1.0f / 200.0f;
    const int steps = argc > 1 ? std::atoi(argv[1]) : 60;
    try {
        LiquidWorld world(DFSPHSolver<>(), PARTICLE_RADIUS, SMOOTHING_FACTOR);
        Fluid fluid({}, PARTICLE_RADIUS, 1000.0f, InteractionGroups());  // faucet3.rs:40-45: no particle yet
        fluid.nonpressure_forces.push_back(std::make_shared<XSPHViscosity>(0.5f, 0.0f));
        fluid.nonpressure_forces.push_back(std::make_shared<Akinci2013SurfaceTension>(1.0f, 10.0f));
        const FluidHandle fh = world.add_fluid(std::move(fluid));
        world.add_boundary(Boundary(ball_surface(0.15f, PARTICLE_RADIUS)));  // the "ground" ball, faucet3.rs:51-65
        Real last_t = 0.0f;
        size_t emitted = 0, deleted = 0, peak = 0;
        for (int s = 0; s < steps; ++s) {
            const Real t = (s + 1) * dt;
            Fluid& f = world.fluids_mut()[fh];
            for (size_t i = 0; i < f.num_particles(); ++i)  // faucet3.rs:77-81
                if (f.positions[i].y < -2.0f) {
                    f.delete_particle_at_next_timestep(i);
                    ++deleted;
                }
            if (t - last_t >= 0.06f) {  // faucet3.rs:83-104
                last_t = t;
                const Real height = 0.6f, diam = PARTICLE_RADIUS * 2.0f;
                const int nparticles = 10;
                const Real shift = -nparticles * PARTICLE_RADIUS;
                std::vector<Point3> particles;
                std::vector<Vector3> velocities;
                for (int i = 0; i < nparticles; ++i)
                    for (int j = 0; j < nparticles; ++j) {
                        particles.push_back({i * diam + shift, height, j * diam + shift});
                        velocities.push_back(Vector3{0.0f, 0.0f, 0.0f});
                    }
                f.add_particles(particles, &velocities);
                emitted += particles.size();
            }
            world.step(dt, Vector3{0.0f, -9.81f, 0.0f});
            peak = std::max(peak, world.fluids()[fh].num_particles());
        }
        const Fluid& f = world.fluids()[fh];
        Real ymin = 1.0e30f, ymax = -1.0e30f;
        for (const Point3& p : f.positions) { ymin = std::min(ymin, p.y); ymax = std::max(ymax, p.y); }
        std::printf("faucet3: %d steps, emitted %zu, deleted %zu, alive %zu (peak %zu), y in [%.3f, %.3f]\n", steps, emitted, deleted, f.num_particles(), peak,
                    f.num_particles() ? ymin : 0.0f, f.num_particles() ? ymax : 0.0f);
        if (f.num_particles() + deleted != emitted) {
            std::fprintf(stderr, "faucet3: particle bookkeeping does not add up\n");
            return 1;
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "faucet3: %s\n", e.what());
        return 2;
    }
    return 0;
}
