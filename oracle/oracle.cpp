// oracle.cpp — CPU restatement of fluid-rs's Simulation (TEST INFRASTRUCTURE, not product).
//
// PARITY UNPINNED: the reference (GossiperLoturot/fluid-rs) ships no tests, golden vectors
// or fixtures, and cannot be compiled here (no rustc/cargo, glam 0.30.7 / ahash 0.8.12 not
// vendored).  This file restates src/3d_multi.rs and src/2d_multi.rs line by line in the
// Rust evaluation order (no FMA contraction: build with -ffp-contract=off) and is pinned only
// by the closed-form / conservation checks in tests/test_oracle.py.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.  The product (libfluid_b200.so) never links or calls it.
//
// Citations: "3d:" = /root/reference/src/3d_multi.rs, "2d:" = src/2d_multi.rs.
// glam semantics relied on (restated from the glam 0.30 public behaviour):
//   Mat*Vec = (col0*v.x + col1*v.y) + col2*v.z ; f32*Mat and Mat*f32 scale each element ;
//   Vec/f32 divides each component ; clamp = max(min) then min(max) ;
//   f32::div_euclid(a,b) = trunc(a/b), minus one when a % b < 0 (b > 0) ;
//   `as i32` truncates toward zero, saturates, NaN -> 0.

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <unordered_map>
#include <utility>
#include <vector>

extern "C" {
// Same field meaning as `struct Config` (3d:3-15); vectors padded to 3 lanes.
struct orc_config {
    int32_t dim;
    float dt;
    int32_t iterations;
    int32_t grid_res;
    float gravity[3];
    float rest_density;
    float dynamic_viscosity;
    float eos_stiffness;
    float eos_power;
    float mouse_radius;
    float clip_min[3];
    float clip_max[3];
    float boundary_damp_dist;
    float pressure_clamp;  // 3d:218 (-0.1) | 2d:212 (-0.0)
};
}

namespace {

// Rust `f as i32`: truncating, saturating, NaN -> 0.
inline int32_t rust_f32_as_i32(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return std::numeric_limits<int32_t>::max();
    if (f <= -2147483648.0f) return std::numeric_limits<int32_t>::min();
    return static_cast<int32_t>(f);
}

// f32::div_euclid (Rust std): q = trunc(a / b); if a % b < 0 { b > 0 ? q - 1 : q + 1 }.
inline float rust_div_euclid(float a, float b) {
    float q = std::trunc(a / b);
    if (std::fmod(a, b) < 0.0f) return b > 0.0f ? q - 1.0f : q + 1.0f;
    return q;
}

template <int D>
struct Rec {               // `struct Particle` (3d:35-41) plus an id and two debug taps
    float pos[D];
    float vel[D];
    float C[D * D];        // column-major: C[D*col + row]
    float mass;
    int32_t id;
    float dbg_density;     // local `density`  (3d:198)
    float dbg_pressure;    // local `pressure` (3d:217)
};

template <int D>
struct Node {              // `struct Cell` (3d:43-48)
    float vel[D];
    float mass;
    bool is_computed;
};

template <int D>
using Key = std::array<int32_t, D>;

template <int D>
struct KeyHash {
    size_t operator()(const Key<D>& k) const {
        uint64_t h = 0x9E3779B97F4A7C15ull;
        for (int a = 0; a < D; ++a) {
            h ^= static_cast<uint32_t>(k[a]) + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        }
        return static_cast<size_t>(h);
    }
};

// grid_search (3d:403-408): row-major, x fastest, last axis outermost.
template <int D, class F>
inline void for_box(const Key<D>& lo, const Key<D>& hi, F&& f) {
    if constexpr (D == 2) {
        for (int32_t y = lo[1]; y < hi[1]; ++y)
            for (int32_t x = lo[0]; x < hi[0]; ++x) f(Key<2>{x, y});
    } else {
        for (int32_t z = lo[2]; z < hi[2]; ++z)
            for (int32_t y = lo[1]; y < hi[1]; ++y)
                for (int32_t x = lo[0]; x < hi[0]; ++x) f(Key<3>{x, y, z});
    }
}

template <int D>
struct Sim {
    orc_config cfg;
    std::unordered_map<Key<D>, std::vector<Rec<D>>, KeyHash<D>> blocks;  // particles_mul
    std::vector<Node<D>> grid;                                            // grid_mul
    Key<D> grid_size{};
    std::vector<int32_t> touched;                                         // sparse_grid
    std::vector<std::vector<Rec<D>>> mailbox;                             // swap_mul
    Key<D> mailbox_size{};
    Key<D> p_lo{}, p_hi{}, a_lo{}, a_hi{};
    double phase_seconds[5] = {0, 0, 0, 0, 0};
    int32_t next_id = 0;
    int64_t dropped = 0;

    // key_from_pos (3d:398-401)
    Key<D> key_of(const float* pos) const {
        Key<D> k;
        const float res = static_cast<float>(cfg.grid_res);
        for (int a = 0; a < D; ++a) k[a] = rust_f32_as_i32(rust_div_euclid(pos[a], res));
        return k;
    }

    // set_rect (3d:79-102)
    void set_rect(const float* mn, const float* mx) {
        Key<D> kmin = key_of(mn), kmax = key_of(mx);
        for (int a = 0; a < D; ++a) {
            a_lo[a] = kmin[a];
            a_hi[a] = kmax[a] + 1;
            p_lo[a] = a_lo[a] - 1;
            p_hi[a] = a_hi[a] + 1;
        }
        for_box<D>(p_lo, p_hi, [&](Key<D> k) { (void)blocks[k]; });
        int64_t n_nodes = 1, n_boxes = 1;
        for (int a = 0; a < D; ++a) {
            grid_size[a] = (p_hi[a] - p_lo[a]) * cfg.grid_res;
            mailbox_size[a] = p_hi[a] - p_lo[a];
            n_nodes *= grid_size[a];
            n_boxes *= mailbox_size[a];
        }
        grid.assign(static_cast<size_t>(n_nodes), Node<D>{});
        mailbox.assign(static_cast<size_t>(n_boxes), {});
        // The reference leaves sparse_grid as is (stale indices); resetting is the only safe
        // reading when the grid is reallocated.
        touched.clear();
    }

    // add_particle (3d:104-108)
    void add(const Rec<D>& r) { blocks[key_of(r.pos)].push_back(r); }

    // Per-particle stencil data shared by the three phases (3d:153-161).
    struct Stencil {
        int32_t cell[D];
        float W[3][D];
    };
    static inline void make_stencil(const Rec<D>& p, Stencil& s) {
        for (int a = 0; a < D; ++a) {
            float fl = std::floor(p.pos[a]);
            s.cell[a] = rust_f32_as_i32(fl);
            float c = p.pos[a] - (static_cast<float>(s.cell[a]) + 0.5f);
            // quadratic_weights (3d:390-396)
            s.W[0][a] = (0.5f * (0.5f - c)) * (0.5f - c);
            s.W[1][a] = 0.75f - c * c;
            s.W[2][a] = (0.5f * (0.5f + c)) * (0.5f + c);
        }
    }
    // Node visit: n in {0,1,2}^D; returns linear index or -1 when outside the p_rect grid
    // (3d:166-172).  dn = pos - (cell_n + 0.5), w = product of per-axis weights.
    inline int64_t visit(const Rec<D>& p, const Stencil& s, const int* n, float* dn,
                         float& w) const {
        int32_t cn[D];
        bool outside = false;
        for (int a = 0; a < D; ++a) {
            cn[a] = s.cell[a] + n[a] - 1;
            dn[a] = p.pos[a] - (static_cast<float>(cn[a]) + 0.5f);
            if (cn[a] < p_lo[a] * cfg.grid_res) outside = true;
            if (cn[a] >= p_hi[a] * cfg.grid_res) outside = true;
        }
        if constexpr (D == 2) w = s.W[n[0]][0] * s.W[n[1]][1];
        else w = s.W[n[0]][0] * s.W[n[1]][1] * s.W[n[2]][2];
        if (outside) return -1;
        int32_t idx = 0, stride = 1;
        for (int a = 0; a < D; ++a) {
            idx += (cn[a] - p_lo[a] * cfg.grid_res) * stride;
            stride *= grid_size[a];
        }
        return idx;
    }
    template <class F>
    static inline void for_stencil(F&& f) {
        int n[3] = {0, 0, 0};
        if constexpr (D == 2) {
            for (n[1] = 0; n[1] < 3; ++n[1])
                for (n[0] = 0; n[0] < 3; ++n[0]) f(n);
        } else {
            for (n[2] = 0; n[2] < 3; ++n[2])
                for (n[1] = 0; n[1] < 3; ++n[1])
                    for (n[0] = 0; n[0] < 3; ++n[0]) f(n);
        }
    }
    // glam Mat*Vec: (col0*x + col1*y) + col2*z
    static inline void mat_vec(const float* M, const float* v, float* out) {
        for (int r = 0; r < D; ++r) {
            float acc = M[r] * v[0];
            for (int c = 1; c < D; ++c) acc = acc + M[D * c + r] * v[c];
            out[r] = acc;
        }
    }

    // clear_grid (3d:136-146)
    void clear_grid() {
        for (int32_t idx : touched) {
            Node<D>& nd = grid[static_cast<size_t>(idx)];
            for (int a = 0; a < D; ++a) nd.vel[a] = 0.0f;
            nd.mass = 0.0f;
            nd.is_computed = false;
        }
        touched.clear();
    }

    // p2g_1 (3d:148-183)
    void p2g_1() {
        for_box<D>(p_lo, p_hi, [&](Key<D> k) {
            const std::vector<Rec<D>>& list = blocks.find(k)->second;
            for (const Rec<D>& p : list) {
                Stencil s;
                make_stencil(p, s);
                for_stencil([&](const int* n) {
                    float dn[D], w;
                    int64_t idx = visit(p, s, n, dn, w);
                    float nd[D], q[D];
                    for (int a = 0; a < D; ++a) nd[a] = -dn[a];
                    mat_vec(p.C, nd, q);
                    float mc = w * p.mass;
                    if (idx >= 0) {
                        Node<D>& g = grid[static_cast<size_t>(idx)];
                        g.mass += mc;
                        for (int a = 0; a < D; ++a) g.vel[a] += mc * (p.vel[a] + q[a]);
                        touched.push_back(static_cast<int32_t>(idx));
                    }
                });
            }
        });
    }

    // p2g_2 (3d:185-247)
    void p2g_2() {
        for_box<D>(p_lo, p_hi, [&](Key<D> k) {
            std::vector<Rec<D>>& list = blocks.find(k)->second;
            for (Rec<D>& p : list) {
                Stencil s;
                make_stencil(p, s);
                float density = 0.0f;
                for_stencil([&](const int* n) {
                    float dn[D], w;
                    int64_t idx = visit(p, s, n, dn, w);
                    if (idx >= 0) density += grid[static_cast<size_t>(idx)].mass * w;
                });
                float volume = p.mass / density;
                float eos = cfg.eos_stiffness *
                            (std::pow(density / cfg.rest_density, cfg.eos_power) - 1.0f);
                float pressure = std::fmax(cfg.pressure_clamp, eos);  // f32::max(clamp, eos)
                p.dbg_density = density;
                p.dbg_pressure = pressure;

                float T[D * D];
                const float s1 = -4.0f * volume;
                for (int c = 0; c < D; ++c) {
                    for (int r = 0; r < D; ++r) {
                        float strain = p.C[D * c + r] + p.C[D * r + c];
                        float visc = cfg.dynamic_viscosity * strain;
                        float ident = (c == r) ? 1.0f : 0.0f;
                        float stress = (-pressure) * ident + visc;
                        T[D * c + r] = (s1 * stress) * cfg.dt;
                    }
                }
                for_stencil([&](const int* n) {
                    float dn[D], w;
                    int64_t idx = visit(p, s, n, dn, w);
                    if (idx >= 0) {
                        float M[D * D], nd[D], f[D];
                        for (int e = 0; e < D * D; ++e) M[e] = T[e] * w;
                        for (int a = 0; a < D; ++a) nd[a] = -dn[a];
                        mat_vec(M, nd, f);
                        Node<D>& g = grid[static_cast<size_t>(idx)];
                        for (int a = 0; a < D; ++a) g.vel[a] += f[a];
                    }
                });
            }
        });
    }

    // update_grid (3d:249-259)
    void update_grid() {
        for (int32_t idx : touched) {
            Node<D>& g = grid[static_cast<size_t>(idx)];
            if (!g.is_computed && g.mass > 0.0f) {
                for (int a = 0; a < D; ++a) g.vel[a] = g.vel[a] / g.mass;
                for (int a = 0; a < D; ++a) g.vel[a] += cfg.dt * cfg.gravity[a];
                g.is_computed = true;
            }
        }
    }

    // g2p (3d:261-381)
    void g2p(const float* mouse) {
        std::vector<std::pair<int32_t, Key<D>>> movers;
        for_box<D>(a_lo, a_hi, [&](Key<D> k) {
            std::vector<Rec<D>>& list = blocks.find(k)->second;
            movers.clear();
            for (size_t i = 0; i < list.size(); ++i) {
                Rec<D>& p = list[i];
                for (int a = 0; a < D; ++a) p.vel[a] = 0.0f;
                Stencil s;
                make_stencil(p, s);
                float B[D * D];
                for (int e = 0; e < D * D; ++e) B[e] = 0.0f;
                for_stencil([&](const int* n) {
                    float dn[D], w;
                    int64_t idx = visit(p, s, n, dn, w);
                    if (idx >= 0) {
                        const Node<D>& g = grid[static_cast<size_t>(idx)];
                        float wv[D];
                        for (int a = 0; a < D; ++a) wv[a] = g.vel[a] * w;
                        for (int c = 0; c < D; ++c)
                            for (int r = 0; r < D; ++r) B[D * c + r] += wv[r] * (-dn[c]);
                        for (int a = 0; a < D; ++a) p.vel[a] += wv[a];
                    }
                });
                for (int e = 0; e < D * D; ++e) p.C[e] = 4.0f * B[e];
                for (int a = 0; a < D; ++a) p.pos[a] += p.vel[a] * cfg.dt;

                if (mouse) {  // 3d:305-310 (xy only)
                    float dx = p.pos[0] - mouse[0], dy = p.pos[1] - mouse[1];
                    float len2 = dx * dx + dy * dy;
                    if (len2 < cfg.mouse_radius * cfg.mouse_radius) {
                        // normalize_or_zero: v * (1/len) when that is finite and > 0
                        float rcp = 1.0f / std::sqrt(len2);
                        if (std::isfinite(rcp) && rcp > 0.0f) {
                            p.vel[0] += dx * rcp;
                            p.vel[1] += dy * rcp;
                        }
                    }
                }
                for (int a = 0; a < D; ++a) {  // Vec3::clamp = max(min).min(max)
                    float v = p.pos[a];
                    v = (v > cfg.clip_min[a]) ? v : cfg.clip_min[a];
                    v = (v < cfg.clip_max[a]) ? v : cfg.clip_max[a];
                    p.pos[a] = v;
                }
                for (int a = 0; a < D; ++a) {  // predictive soft wall (3d:320-343)
                    float nxt = p.pos[a] + p.vel[a];
                    float wmin = cfg.clip_min[a] + cfg.boundary_damp_dist;
                    float wmax = cfg.clip_max[a] - cfg.boundary_damp_dist;
                    if (nxt < wmin) p.vel[a] += wmin - nxt;
                    if (nxt > wmax) p.vel[a] += wmax - nxt;
                }
                Key<D> nk = key_of(p.pos);
                if (nk != k) movers.emplace_back(static_cast<int32_t>(i), nk);
            }
            // reverse order + swap_remove (3d:353-367)
            for (auto it = movers.rbegin(); it != movers.rend(); ++it) {
                size_t i = static_cast<size_t>(it->first);
                Rec<D> p = list[i];
                list[i] = list.back();
                list.pop_back();
                const Key<D>& nk = it->second;
                bool outside = false;
                for (int a = 0; a < D; ++a)
                    if (nk[a] < p_lo[a] || nk[a] >= p_hi[a]) outside = true;
                if (outside) {
                    ++dropped;
                    continue;
                }
                int64_t bi = 0, stride = 1;
                for (int a = 0; a < D; ++a) {
                    bi += (nk[a] - p_lo[a]) * stride;
                    stride *= mailbox_size[a];
                }
                mailbox[static_cast<size_t>(bi)].push_back(p);
            }
        });
        // deliver (3d:370-380)
        for_box<D>(p_lo, p_hi, [&](Key<D> k) {
            int64_t bi = 0, stride = 1;
            for (int a = 0; a < D; ++a) {
                bi += (k[a] - p_lo[a]) * stride;
                stride *= mailbox_size[a];
            }
            std::vector<Rec<D>>& box = mailbox[static_cast<size_t>(bi)];
            std::vector<Rec<D>>& list = blocks.find(k)->second;
            list.insert(list.end(), box.begin(), box.end());
            box.clear();
        });
    }

    void run_phase(int ph, const float* mouse) {
        auto t0 = std::chrono::steady_clock::now();
        switch (ph) {
            case 0: clear_grid(); break;
            case 1: p2g_1(); break;
            case 2: p2g_2(); break;
            case 3: update_grid(); break;
            case 4: g2p(mouse); break;
            default: return;
        }
        phase_seconds[ph] =
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    // one iteration of step()'s loop (3d:111-133)
    void substep(const float* mouse) {
        for (int ph = 0; ph < 5; ++ph) run_phase(ph, mouse);
    }

    template <class F>
    void for_active(F&& f) const {  // iter_particle (3d:383-387)
        for_box<D>(a_lo, a_hi, [&](Key<D> k) {
            auto it = blocks.find(k);
            if (it == blocks.end()) return;
            for (const Rec<D>& p : it->second) f(p);
        });
    }
    template <class F>
    void for_deposit(F&& f) const {  // every particle the p2g phases visit (p_rect order)
        for_box<D>(p_lo, p_hi, [&](Key<D> k) {
            auto it = blocks.find(k);
            if (it == blocks.end()) return;
            for (const Rec<D>& p : it->second) f(p);
        });
    }
};

struct Handle {
    int dim;
    Sim<2>* s2 = nullptr;
    Sim<3>* s3 = nullptr;
};

template <int D>
void unpack(const float* rec, int32_t id, Rec<D>& r) {
    const float* q = rec;
    for (int a = 0; a < D; ++a) r.pos[a] = *q++;
    for (int a = 0; a < D; ++a) r.vel[a] = *q++;
    for (int e = 0; e < D * D; ++e) r.C[e] = *q++;
    r.mass = *q++;
    r.id = id;
    r.dbg_density = 0.0f;
    r.dbg_pressure = 0.0f;
}
template <int D>
void pack(const Rec<D>& r, float* rec) {
    float* q = rec;
    for (int a = 0; a < D; ++a) *q++ = r.pos[a];
    for (int a = 0; a < D; ++a) *q++ = r.vel[a];
    for (int e = 0; e < D * D; ++e) *q++ = r.C[e];
    *q++ = r.mass;
}

#define DISPATCH(h, ...)                   \
    do {                                   \
        if ((h)->dim == 2) {               \
            auto& S = *(h)->s2;            \
            constexpr int D = 2;           \
            (void)D;                       \
            __VA_ARGS__;                   \
        } else {                           \
            auto& S = *(h)->s3;            \
            constexpr int D = 3;           \
            (void)D;                       \
            __VA_ARGS__;                   \
        }                                  \
    } while (0)

}  // namespace

extern "C" {

void orc_config_default(int32_t dim, orc_config* c) {  // Config::default (3d:17-33, 2d:17-33)
    std::memset(c, 0, sizeof(*c));
    c->dim = dim;
    c->dt = dim == 2 ? 0.032f : 0.066f;
    c->iterations = static_cast<int32_t>(1.0 / 0.032);
    c->grid_res = dim == 2 ? 32 : 16;
    c->gravity[1] = 0.3f;
    c->rest_density = dim == 2 ? 4.0f : 1.0f;
    c->dynamic_viscosity = 0.1f;
    c->eos_stiffness = 10.0f;
    c->eos_power = 4.0f;
    c->mouse_radius = 10.0f;
    for (int a = 0; a < 3; ++a) {
        c->clip_min[a] = 0.0f;
        c->clip_max[a] = 64.0f;
    }
    c->boundary_damp_dist = 3.0f;
    c->pressure_clamp = dim == 2 ? -0.0f : -0.1f;
}

void* orc_create(const orc_config* cfg) {
    if (!cfg || (cfg->dim != 2 && cfg->dim != 3)) return nullptr;
    Handle* h = new Handle;
    h->dim = cfg->dim;
    if (cfg->dim == 2) {
        h->s2 = new Sim<2>;
        h->s2->cfg = *cfg;
    } else {
        h->s3 = new Sim<3>;
        h->s3->cfg = *cfg;
    }
    return h;
}
void orc_destroy(void* vh) {
    Handle* h = static_cast<Handle*>(vh);
    if (!h) return;
    delete h->s2;
    delete h->s3;
    delete h;
}
void orc_set_rect(void* vh, const float* mn, const float* mx) {
    Handle* h = static_cast<Handle*>(vh);
    DISPATCH(h, S.set_rect(mn, mx));
}
void orc_add_particles(void* vh, const float* recs, const int32_t* ids, int64_t n) {
    Handle* h = static_cast<Handle*>(vh);
    DISPATCH(h, {
        const int stride = 2 * D + D * D + 1;
        for (int64_t i = 0; i < n; ++i) {
            Rec<D> r;
            unpack<D>(recs + i * stride, ids ? ids[i] : S.next_id, r);
            S.next_id = (ids ? ids[i] : S.next_id) + 1;
            S.add(r);
        }
    });
}
void orc_substeps(void* vh, int32_t n, const float* mouse) {
    Handle* h = static_cast<Handle*>(vh);
    DISPATCH(h, { for (int32_t i = 0; i < n; ++i) S.substep(mouse); });
}
void orc_step(void* vh, const float* mouse) {  // step (3d:110-134)
    Handle* h = static_cast<Handle*>(vh);
    DISPATCH(h, { for (int32_t i = 0; i < S.cfg.iterations; ++i) S.substep(mouse); });
}
void orc_phase(void* vh, int32_t ph, const float* mouse) {
    Handle* h = static_cast<Handle*>(vh);
    DISPATCH(h, S.run_phase(ph, mouse));
}
int64_t orc_count(void* vh, int32_t which) {  // 0: a_rect blocks, 1: p_rect blocks, 2: dropped
    Handle* h = static_cast<Handle*>(vh);
    int64_t n = 0;
    DISPATCH(h, {
        if (which == 0) S.for_active([&](const Rec<D>&) { ++n; });
        else if (which == 1) S.for_deposit([&](const Rec<D>&) { ++n; });
        else n = S.dropped;
    });
    return n;
}
// Read particles in iter_particle order (which=0) or p_rect deposit order (which=1).
int64_t orc_read(void* vh, int32_t which, float* recs, int32_t* ids, float* density,
                 float* pressure, int32_t* cell, int32_t* key) {
    Handle* h = static_cast<Handle*>(vh);
    int64_t i = 0;
    DISPATCH(h, {
        const int stride = 2 * D + D * D + 1;
        auto emit = [&](const Rec<D>& p) {
            if (recs) pack<D>(p, recs + i * stride);
            if (ids) ids[i] = p.id;
            if (density) density[i] = p.dbg_density;
            if (pressure) pressure[i] = p.dbg_pressure;
            if (cell)
                for (int a = 0; a < D; ++a)
                    cell[i * D + a] = rust_f32_as_i32(std::floor(p.pos[a]));
            if (key) {
                Key<D> k = S.key_of(p.pos);
                for (int a = 0; a < D; ++a) key[i * D + a] = k[a];
            }
            ++i;
        };
        if (which == 0) S.for_active(emit);
        else S.for_deposit(emit);
    });
    return i;
}
void orc_rects(void* vh, int32_t* a_lo, int32_t* a_hi, int32_t* p_lo, int32_t* p_hi,
               int32_t* origin, int32_t* size) {
    Handle* h = static_cast<Handle*>(vh);
    DISPATCH(h, {
        for (int a = 0; a < D; ++a) {
            a_lo[a] = S.a_lo[a];
            a_hi[a] = S.a_hi[a];
            p_lo[a] = S.p_lo[a];
            p_hi[a] = S.p_hi[a];
            origin[a] = S.p_lo[a] * S.cfg.grid_res;
            size[a] = S.grid_size[a];
        }
    });
}
// Node grid, reference layout: per node vel[D] then mass.
int64_t orc_read_grid(void* vh, float* out) {
    Handle* h = static_cast<Handle*>(vh);
    int64_t n = 0;
    DISPATCH(h, {
        n = static_cast<int64_t>(S.grid.size());
        if (out)
            for (int64_t i = 0; i < n; ++i) {
                for (int a = 0; a < D; ++a) out[i * (D + 1) + a] = S.grid[i].vel[a];
                out[i * (D + 1) + D] = S.grid[i].mass;
            }
    });
    return n;
}
int64_t orc_touched_count(void* vh) {
    Handle* h = static_cast<Handle*>(vh);
    int64_t n = 0;
    DISPATCH(h, n = static_cast<int64_t>(S.touched.size()));
    return n;
}
void orc_phase_seconds(void* vh, double* out) {
    Handle* h = static_cast<Handle*>(vh);
    DISPATCH(h, { for (int i = 0; i < 5; ++i) out[i] = S.phase_seconds[i]; });
}
// Stand-alone helpers so the integer rules can be tested directly.
void orc_key_from_pos(const float* pos, int64_t n, int32_t dim, int32_t grid_res,
                      int32_t* key, int32_t* cell) {
    const float res = static_cast<float>(grid_res);
    for (int64_t i = 0; i < n * dim; ++i) {
        if (key) key[i] = rust_f32_as_i32(rust_div_euclid(pos[i], res));
        if (cell) cell[i] = rust_f32_as_i32(std::floor(pos[i]));
    }
}
void orc_quadratic_weights(float c, float* w3) {  // 3d:390-396 for one lane
    w3[0] = (0.5f * (0.5f - c)) * (0.5f - c);
    w3[1] = 0.75f - c * c;
    w3[2] = (0.5f * (0.5f + c)) * (0.5f + c);
}

}  // extern "C"
