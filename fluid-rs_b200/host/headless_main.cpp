// headless_main.cpp — the reference's `main` (3d:514-568) without the terminal: seed the default
// scene, step it, print particle count and the five phase timers.  Build:
//   g++ -std=c++17 -O2 -I include fluid-rs_b200/host/headless_main.cpp -L fluid-rs_b200/csrc -lfluid_b200
#include <cstdio>
#include <random>

#include "simulation.hpp"

int main() {
    using namespace fluid_b200;
    Simulation<3> sim(default_config(3));
    std::mt19937 rng(20260101);
    std::uniform_real_distribution<float> u(16.0f, 32.0f);   // 3d:528-530
    std::vector<Particle<3>> ps(4096);
    for (auto& p : ps) {
        for (float& x : p.pos) x = u(rng);
        p.mass = 1.0f;
    }
    sim.add_particles(ps);                       // particles first, then set_rect, as main does (3d:525-537)
    sim.set_rect({0, 0, 0}, {64, 64, 64});
    for (int frame = 0; frame < 10; ++frame) sim.step();
    std::printf("particles: %zu\n", sim.iter_particle().size());
    for (auto& [label, sec] : sim.debug_elapseds) std::printf("%s: %.3f us\n", label.c_str(), sec * 1e6);
    return 0;
}
