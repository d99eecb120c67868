// examples3d/basic3.rs (basic3.rs:16-118) on the C++ host mirror: 15^3 fluid block, r = 0.05, DFSPH +
// ArtificialViscosity(1, 0), dt = 1/200, a ground plane of boundary particles.  Prints the centre of mass.
//   g++ -std=c++17 -Iinclude examples/basic3.cpp -Lsalva_b200 -lsalva_b200 -Wl,-rpath,$PWD/salva_b200 -o basic3
#include <cstdio>
#include <cstdlib>

#include "salva3d_b200.hpp"

using namespace salva3d;

// examples3d/helper.rs:4-20
static Fluid cube_fluid(int ni, int nj, int nk, float particle_rad, float density) {
    std::vector<Point3> points;
    float hx = ni * particle_rad, hy = nj * particle_rad, hz = nk * particle_rad;
    for (int i = 0; i < ni; ++i)
        for (int j = 0; j < nj; ++j)
            for (int k = 0; k < nk; ++k)
                points.push_back({i * particle_rad * 2.0f + particle_rad - hx, j * particle_rad * 2.0f + particle_rad - hy,
                                  k * particle_rad * 2.0f + particle_rad - hz});
    return Fluid(points, particle_rad, density, InteractionGroups());
}

int main(int argc, char** argv) {
    const float PARTICLE_RADIUS = 0.05f, SMOOTHING_FACTOR = 2.0f;
    int steps = argc > 1 ? atoi(argv[1]) : 20;
    try {
        LiquidWorld world(DFSPHSolver<>(), PARTICLE_RADIUS, SMOOTHING_FACTOR);
        const int n = 15;
        Fluid fluid = cube_fluid(n, n, n, PARTICLE_RADIUS, 1000.0f);
        for (auto& p : fluid.positions) p.y += 0.2f + n * PARTICLE_RADIUS;  // transform_by(translation) basic3.rs:37-41
        fluid.nonpressure_forces.push_back(std::make_shared<ArtificialViscosity>(1.0f, 0.0f));
        FluidHandle fh = world.add_fluid(std::move(fluid));
        std::vector<Point3> ground;  // top face of the ground cuboid, sampled at 2r
        for (int i = -25; i <= 25; ++i)
            for (int k = -25; k <= 25; ++k) ground.push_back({i * 0.1f, 0.2f, k * 0.1f});
        world.add_boundary(Boundary(ground));
        for (int s = 0; s < steps; ++s) world.step(1.0f / 200.0f, Vector3{0.0f, -9.81f, 0.0f});
        const Fluid& f = world.fluids()[fh];
        double cx = 0, cy = 0, cz = 0;
        for (auto& p : f.positions) { cx += p.x; cy += p.y; cz += p.z; }
        size_t np = f.num_particles();
        sph_step_stats st = world.counters();
        printf("basic3: %zu particles, %d steps, h = %.3f, centre of mass = (%.6f, %.6f, %.6f), last step %.3f ms, %u+%u iterations\n", np, steps, world.h(),
               cx / np, cy / np, cz / np, st.step_ms, st.n_divergence_iter, st.n_pressure_iter);
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
    return 0;
}
