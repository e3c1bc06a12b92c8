// simulation.hpp — header-only C++ mirror of fluid-rs's `Simulation` (src/3d_multi.rs:50-134,383-387)
// over the C ABI of include/fluid_b200.h.  Same method names and argument meaning as the reference;
// errors become exceptions (the reference panics via unwrap() in the same situations).
#pragma once

#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/fluid_b200.h"

namespace fluid_b200 {

struct Error : std::runtime_error {
    fluid_status status;
    Error(fluid_status st, const std::string& msg) : std::runtime_error(msg), status(st) {}
};

inline void check(fluid_status st) {
    if (st != FLUID_OK) throw Error(st, fluid_last_error());
}

// `Config::default()` (3d:17-33 / 2d:17-33)
inline fluid_config default_config(int dim) {
    fluid_config c{};
    check(fluid_config_default(dim, &c));
    return c;
}

// `struct Particle` (3d:35-41): pos, vel, affine_momentum (column-major), mass — packed f32.
template <int DIM>
struct Particle {
    float pos[DIM]{};
    float vel[DIM]{};
    float affine_momentum[DIM * DIM]{};
    float mass = 1.0f;
};
static_assert(sizeof(Particle<3>) == 16 * sizeof(float), "3D record is 16 packed floats");
static_assert(sizeof(Particle<2>) == 9 * sizeof(float), "2D record is 9 packed floats");

template <int DIM>
class Simulation {
  public:
    explicit Simulation(const fluid_config& config, int device = 0) : config(config) {   // Simulation::new, 3d:64
        if (config.dim != DIM) throw Error(FLUID_ERR_INVALID_ARG, "config.dim does not match Simulation<DIM>");
        check(fluid_create(&config, device, &h_));
    }
    ~Simulation() { fluid_destroy(h_); }
    Simulation(const Simulation&) = delete;
    Simulation& operator=(const Simulation&) = delete;

    void set_rect(const std::array<float, DIM>& min, const std::array<float, DIM>& max) {   // 3d:79
        check(fluid_set_rect(h_, min.data(), max.data()));
    }
    void add_particle(const Particle<DIM>& p) {                                             // 3d:104
        check(fluid_add_particles(h_, reinterpret_cast<const float*>(&p), nullptr, 1));
    }
    void add_particles(const std::vector<Particle<DIM>>& ps) {
        check(fluid_add_particles(h_, reinterpret_cast<const float*>(ps.data()), nullptr,
                                  static_cast<int64_t>(ps.size())));
    }
    // step(&mut self, mouse_pos: &Option<Vec2>) (3d:110): nullptr = None
    void step(const float* mouse_xy = nullptr) {
        check(fluid_step(h_, mouse_xy));
        double sec[FLUID_NUM_PHASES];
        check(fluid_get_phase_times(h_, sec, nullptr));
        debug_elapseds.clear();
        for (int i = 0; i < FLUID_NUM_PHASES; ++i) debug_elapseds.emplace_back(fluid_phase_label(i), sec[i]);
    }
    // iter_particle (3d:383): a host copy of every particle stored in an a_rect block
    const std::vector<Particle<DIM>>& iter_particle(std::vector<int32_t>* ids = nullptr) {
        int64_t n = 0;
        check(fluid_slot_count(h_, &n));   // upper bound, no device work
        readback_.resize(static_cast<size_t>(n));
        if (ids) ids->resize(static_cast<size_t>(n));
        check(fluid_read_particles(h_, reinterpret_cast<float*>(readback_.data()), ids ? ids->data() : nullptr, n, &n));
        readback_.resize(static_cast<size_t>(n));
        return readback_;
    }
    float dt() const { return config.dt; }   // sim.config.dt (3d:541)

    // beyond the reference's six: bit-reproducible node sums (the reference fixes every sum's order, 3d:403-408),
    // block-sparse node storage (its hash map of blocks, 3d:52-55), and what that storage takes
    void set_deterministic(bool on) { check(fluid_set_deterministic(h_, on ? 1 : 0)); }
    void set_sparse(int64_t max_blocks) { check(fluid_set_sparse(h_, max_blocks)); }   // before set_rect
    std::array<int64_t, 6> memory_stats() {
        std::array<int64_t, 6> o{};
        check(fluid_memory_stats(h_, o.data()));
        return o;
    }

    fluid_config config;
    std::vector<std::pair<std::string, double>> debug_elapseds;   // (label, seconds), 3d:60,112-132

  private:
    fluid_sim* h_ = nullptr;
    std::vector<Particle<DIM>> readback_;
};

}  // namespace fluid_b200
