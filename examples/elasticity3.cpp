// examples3d/elasticity3.rs (elasticity3.rs:20-110) on the C++ host mirror, without the rapier testbed: two 12 x 6 x 12 blocks
// with Becker2009Elasticity (Young modulus 500 000 and 100 000, Poisson ratio 0.3, non-linear strain) + XSPHViscosity(0.5, 1),
// dropped one above the other onto a ground plate.  The reference samples the ground through rapier's contact sampling; here the
// plate's top face is sampled once at the particle spacing.  Prints both centres of mass and the blocks' extents.
//   g++ -std=c++17 -Iinclude examples/elasticity3.cpp -Lsalva_b200 -lsalva_b200 -Wl,-rpath,$PWD/salva_b200 -o elasticity3
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
    const float PARTICLE_RADIUS = 0.025f, SMOOTHING_FACTOR = 2.0f;
    const int steps = argc > 1 ? std::atoi(argv[1]) : 40;
    try {
        LiquidWorld world(DFSPHSolver<>(), PARTICLE_RADIUS, SMOOTHING_FACTOR);
        const float ground_thickness = 0.2f, ground_half_width = 1.5f, height = 0.4f;
        const int nparticles = 6;
        FluidHandle handles[2];
        const float young[2] = {500000.0f, 100000.0f}, lift[2] = {1.0f, 4.0f};  // elasticity3.rs:41-44,66-69
        for (int b = 0; b < 2; ++b) {
            Fluid fluid = cube_fluid(nparticles * 2, nparticles, nparticles * 2, PARTICLE_RADIUS, 1000.0f);
            for (auto& p : fluid.positions) p.y += ground_thickness + PARTICLE_RADIUS * nparticles * lift[b] + height;
            fluid.nonpressure_forces.push_back(std::make_shared<Becker2009Elasticity>(young[b], 0.3f, true));
            fluid.nonpressure_forces.push_back(std::make_shared<XSPHViscosity>(0.5f, 1.0f));
            handles[b] = world.add_fluid(std::move(fluid));
        }
        std::vector<Point3> ground;  // top face of the ground cuboid (half extents 1.5, 0.2, 1.5) at spacing 2r
        const int half = (int)(ground_half_width / (2.0f * PARTICLE_RADIUS));
        for (int i = -half; i <= half; ++i)
            for (int k = -half; k <= half; ++k) ground.push_back({i * 2.0f * PARTICLE_RADIUS, ground_thickness, k * 2.0f * PARTICLE_RADIUS});
        world.add_boundary(Boundary(ground));
        for (int s = 0; s < steps; ++s) world.step(1.0f / 200.0f, Vector3{0.0f, -9.81f, 0.0f});
        for (int b = 0; b < 2; ++b) {
            const Fluid& f = world.fluids()[handles[b]];
            double cy = 0;
            float ymin = 1e30f, ymax = -1e30f;
            for (auto& p : f.positions) { cy += p.y; ymin = std::min(ymin, p.y); ymax = std::max(ymax, p.y); }
            std::printf("elasticity3: block %d (E = %.0f): %zu particles, centre y = %.5f, y in [%.4f, %.4f]\n", b, young[b], f.num_particles(),
                        cy / f.num_particles(), ymin, ymax);
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "elasticity3: %s\n", e.what());
        return 2;
    }
    return 0;
}
