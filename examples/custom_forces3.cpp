// examples3d/custom_forces3.rs (custom_forces3.rs:18-90) on the C++ host mirror: a 10^3 block, zero gravity, two user
// NonPressureForce plugins (attractors at (+-1, 0, 0)) pushed as trait objects.  Each plugin runs as arbitrary host code
// once per step through sph_fluid_push_host_force.
//   g++ -std=c++17 -Iinclude examples/custom_forces3.cpp -Lsalva_b200 -lsalva_b200 -Wl,-rpath,$PWD/salva_b200 -o custom_forces3
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "salva3d_b200.hpp"

using namespace salva3d;

// custom_forces3.rs:66-90: acc += dir / dist towards `origin` (Unit::try_new_and_get with min norm 0.1)
struct CustomForceField : CustomNonPressureForce {
    Point3 origin;
    explicit CustomForceField(Point3 o) : origin(o) {}
    void solve(Real, Real, Real, size_t n, const Point3* positions, const Vector3*, const Real*, Vector3* accelerations) override {
        for (size_t i = 0; i < n; ++i) {
            float dx = origin.x - positions[i].x, dy = origin.y - positions[i].y, dz = origin.z - positions[i].z;
            float dist = std::sqrt(dx * dx + dy * dy + dz * dz);
            if (dist > 0.1f) {
                accelerations[i].x += dx / dist / dist;
                accelerations[i].y += dy / dist / dist;
                accelerations[i].z += dz / dist / dist;
            }
        }
    }
};

int main(int argc, char** argv) {
    const float PARTICLE_RADIUS = 0.025f, SMOOTHING_FACTOR = 2.0f;
    int steps = argc > 1 ? atoi(argv[1]) : 20;
    try {
        LiquidWorld world(DFSPHSolver<>(), PARTICLE_RADIUS, SMOOTHING_FACTOR);
        const int n = 10;
        std::vector<Point3> points;  // examples3d/helper.rs:4-20
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j)
                for (int k = 0; k < n; ++k)
                    points.push_back({i * PARTICLE_RADIUS * 2.0f + PARTICLE_RADIUS - n * PARTICLE_RADIUS,
                                      j * PARTICLE_RADIUS * 2.0f + PARTICLE_RADIUS - n * PARTICLE_RADIUS,
                                      k * PARTICLE_RADIUS * 2.0f + PARTICLE_RADIUS - n * PARTICLE_RADIUS});
        Fluid fluid(points, PARTICLE_RADIUS, 1000.0f, InteractionGroups());
        fluid.nonpressure_forces.push_back(std::make_shared<CustomForceField>(Point3{1.0f, 0.0f, 0.0f}));
        fluid.nonpressure_forces.push_back(std::make_shared<CustomForceField>(Point3{-1.0f, 0.0f, 0.0f}));
        FluidHandle fh = world.add_fluid(std::move(fluid));
        for (int s = 0; s < steps; ++s) world.step(1.0f / 200.0f, Vector3{0.0f, 0.0f, 0.0f});
        const Fluid& f = world.fluids()[fh];
        double spread = 0;
        for (auto& p : f.positions) spread += std::fabs(p.x);
        printf("custom_forces3: %zu particles, %d steps, mean |x| = %.6f\n", f.num_particles(), steps, spread / f.num_particles());
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
    return 0;
}
