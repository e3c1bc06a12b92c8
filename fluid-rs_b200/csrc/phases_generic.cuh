// phases_generic.cuh — the five reference phases as straightforward particle-per-thread code.
//
// This is the GENERIC path: any dim (2 or 3), any grid_res, any particle density.  It scatters
// with vector float atomics (red.global.add.v4.f32 -> SASS REDG.E.ADD.F32x4) straight into the
// dense node grid and gathers from it through L2.  The tiled sm_100a path in phases_tiled.cuh
// replaces it for large 3D scenes; this one is the 2D path, the fallback for shapes the tiled
// kernels do not cover, and — as k_substeps_resident, one cooperative launch per step() with the
// particle state held in registers — the path of the reference's two default scenes (4,096
// particles: launch latency, not bandwidth, is what bounds them).
//
// Node record in HBM: float4 {momentum.xyz (velocity after update), mass} = `struct Cell`
// (3d:43-48) without `is_computed`: update_grid is folded into g2p's node read, which applies
// `v = mom / mass + dt*g` exactly where the reference's guard `mass > 0` holds (3d:253-256).
//
// DETERMINISTIC mode (template flag DET): the reference fixes the order of every node sum
// (grid_search, 3d:403-408); float atomics do not, so two runs differ in the last bits.  With DET the
// deposits are rounded to 2^-34 and accumulated as 64-bit INTEGERS (red.global.add.u64): integer
// addition is associative, so the node sums — and with them every particle — are bit-for-bit the
// same from run to run, for any particle order and any number of GPUs.  One deposit carries an
// absolute rounding error of 2^-35 (3e-11), far below f32 resolution of the sums it joins.
#pragma once

#include <cooperative_groups.h>

#include "common.cuh"
#include "sort.cuh"

namespace fluid {

constexpr float FIX_SCALE = 17179869184.0f;           // 2^34
constexpr float FIX_INV = 1.0f / 17179869184.0f;

template <int DIM>
__device__ __forceinline__ int node_index(const Geo& g, const Stencil<DIM>& s, int ox, int oy,
                                          int oz) {
    int idx = g.guard + (s.base[0] + ox) + (s.base[1] + oy) * g.size[0];
    if (DIM == 3) idx += (s.base[2] + oz) * g.size[0] * g.size[1];
    return idx;
}

// The node grid as the generic kernels see it: float4 records, or (DET) four 64-bit fixed-point sums.
struct NodeGrid {
    float4* f;
    long long* fx;   // 4 per node: momentum.xyz, mass (DET only)
};

__device__ __forceinline__ long long to_fixed(float v) { return __float2ll_rn(v * FIX_SCALE); }
__device__ __forceinline__ float from_fixed(long long v) { return __ll2float_rn(v) * FIX_INV; }

template <bool DET>
__device__ __forceinline__ void node_add(const NodeGrid& ng, int idx, float4 v) {
    if (!DET) {
        atomicAdd(&ng.f[idx], v);
    } else {
        unsigned long long* p = reinterpret_cast<unsigned long long*>(ng.fx + 4 * static_cast<size_t>(idx));
        const float c[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long q = to_fixed(c[k]);
            if (q != 0) atomicAdd(p + k, static_cast<unsigned long long>(q));
        }
    }
}
template <bool DET>
__device__ __forceinline__ float node_mass(const NodeGrid& ng, int idx) {
    if (!DET) return __ldcg(reinterpret_cast<const float*>(&ng.f[idx]) + 3);   // .w lane only: others are being written
    return from_fixed(__ldcg(ng.fx + 4 * static_cast<size_t>(idx) + 3));
}
// `more`: a second fixed-point buffer whose sums belong to the same nodes (slab runs keep the force deposits of
// "p2g 2" apart so that each halo exchange carries one phase's deposits): added as INTEGERS, then converted, so
// the result equals the single-buffer sum bit for bit.
template <bool DET>
__device__ __forceinline__ float4 node_load(const NodeGrid& ng, int idx, const long long* more = nullptr) {
    if (!DET) return __ldcg(&ng.f[idx]);
    const longlong2* p = reinterpret_cast<const longlong2*>(ng.fx + 4 * static_cast<size_t>(idx));
    longlong2 a = __ldcg(p), b = __ldcg(p + 1);
    if (more) {
        const longlong2* p2 = reinterpret_cast<const longlong2*>(more + 4 * static_cast<size_t>(idx));
        const longlong2 a2 = __ldcg(p2), b2 = __ldcg(p2 + 1);
        a.x += a2.x; a.y += a2.y; b.x += b2.x; b.y += b2.y;
    }
    return make_float4(from_fixed(a.x), from_fixed(a.y), from_fixed(b.x), from_fixed(b.y));
}
template <bool DET>
__device__ __forceinline__ void node_zero(const NodeGrid& ng, int idx) {
    if (!DET) {
        ng.f[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        longlong2* p = reinterpret_cast<longlong2*>(ng.fx + 4 * static_cast<size_t>(idx));
        p[0] = make_longlong2(0, 0);
        p[1] = make_longlong2(0, 0);
    }
}

template <int DIM>
__device__ __forceinline__ void load_C(const Particles& q, int i, float* C) {
    float4 a = q.CA[i];
    C[0] = a.x; C[1] = a.y; C[2] = a.z; C[3] = a.w;
    if (DIM == 3) {
        float4 b = q.CB[i];
        C[4] = b.x; C[5] = b.y; C[6] = b.z; C[7] = b.w;
        C[8] = q.CC[i];
    }
}

// ---- one particle through each phase ------------------------------------------------------------

// p2g_1 (3d:148-183): node.mass += w*m ; node.mom += w*m*(v + C*(x_n - x_p))
template <int DIM, bool DET>
__device__ __forceinline__ void p2g1_particle(const Geo& g, const NodeGrid& ng, const Stencil<DIM>& s,
                                              const float* vel, const float* C, float m) {
    constexpr int NZ = DIM == 3 ? 3 : 1;
#pragma unroll
    for (int oz = 0; oz < NZ; ++oz)
#pragma unroll
        for (int oy = 0; oy < 3; ++oy)
#pragma unroll
            for (int ox = 0; ox < 3; ++ox) {
                float w = s.w[0][ox] * s.w[1][oy];
                if (DIM == 3) w *= s.w[2][oz];
                float d[3] = {s.d[0][ox], s.d[1][oy], DIM == 3 ? s.d[2][oz] : 0.0f};
                float mc = w * m;
                float mom[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int r = 0; r < DIM; ++r) {
                    float qv = 0.0f;
#pragma unroll
                    for (int c = 0; c < DIM; ++c) qv += C[DIM * c + r] * d[c];
                    mom[r] = mc * (vel[r] + qv);
                }
                node_add<DET>(ng, node_index<DIM>(g, s, ox, oy, oz), make_float4(mom[0], mom[1], mom[2], mc));
            }
}

// p2g_2 (3d:185-247): density from node masses, Tait pressure, stress, force scatter.
template <int DIM, bool DET>
__device__ __forceinline__ void p2g2_particle(const Geo& g, const NodeGrid& ng, const NodeGrid& ng_out, const Stencil<DIM>& s,
                                              const float* C, float m, float& density, float& pressure) {
    constexpr int NZ = DIM == 3 ? 3 : 1;
    density = 0.0f;
#pragma unroll
    for (int oz = 0; oz < NZ; ++oz)
#pragma unroll
        for (int oy = 0; oy < 3; ++oy)
#pragma unroll
            for (int ox = 0; ox < 3; ++ox) {
                float w = s.w[0][ox] * s.w[1][oy];
                if (DIM == 3) w *= s.w[2][oz];
                density += node_mass<DET>(ng, node_index<DIM>(g, s, ox, oy, oz)) * w;
            }
    float volume = __fdiv_rn(m, density);
    pressure = tait_pressure(g, density);

    // T = -4 * V * (-p I + mu (C + C^T)) * dt   (3d:222-225)
    float T[9];
    const float s1 = -4.0f * volume;
#pragma unroll
    for (int c = 0; c < DIM; ++c)
#pragma unroll
        for (int r = 0; r < DIM; ++r) {
            float stress = g.mu * (C[DIM * c + r] + C[DIM * r + c]);
            if (c == r) stress -= pressure;
            T[DIM * c + r] = (s1 * stress) * g.dt;
        }
#pragma unroll
    for (int oz = 0; oz < NZ; ++oz)
#pragma unroll
        for (int oy = 0; oy < 3; ++oy)
#pragma unroll
            for (int ox = 0; ox < 3; ++ox) {
                float w = s.w[0][ox] * s.w[1][oy];
                if (DIM == 3) w *= s.w[2][oz];
                float d[3] = {s.d[0][ox], s.d[1][oy], DIM == 3 ? s.d[2][oz] : 0.0f};
                float f[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int r = 0; r < DIM; ++r) {
                    float acc = 0.0f;
#pragma unroll
                    for (int c = 0; c < DIM; ++c) acc += T[DIM * c + r] * d[c];
                    f[r] = w * acc;
                }
                node_add<DET>(ng_out, node_index<DIM>(g, s, ox, oy, oz), make_float4(f[0], f[1], f[2], 0.0f));
            }
}

// Everything g2p does to one particle after the node gather (3d:300-343): C = 4B, advect,
// mouse push, clamp, predictive soft wall.  Shared by the generic and tiled kernels.
template <int DIM>
__device__ __forceinline__ void integrate_particle(const Geo& g, float* pos, float* vel,
                                                   const float* mouse) {
#pragma unroll
    for (int a = 0; a < DIM; ++a) pos[a] += vel[a] * g.dt;
    if (mouse) {  // 3d:305-310: xy only, unit push inside mouse_radius
        float dx = pos[0] - mouse[0], dy = pos[1] - mouse[1];
        float len2 = dx * dx + dy * dy;
        if (len2 < g.mouse_r2) {
            float rcp = __frcp_rn(__fsqrt_rn(len2));   // normalize_or_zero
            if (isfinite(rcp) && rcp > 0.0f) {
                vel[0] += dx * rcp;
                vel[1] += dy * rcp;
            }
        }
    }
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
        float x = pos[a];
        x = (x > g.clip_lo[a]) ? x : g.clip_lo[a];   // Vec3::clamp = max(min).min(max)
        x = (x < g.clip_hi[a]) ? x : g.clip_hi[a];
        pos[a] = x;
        float nxt = x + vel[a];                        // pos + vel, NOT vel*dt (3d:320)
        if (nxt < g.wall_lo[a]) vel[a] += g.wall_lo[a] - nxt;
        if (nxt > g.wall_hi[a]) vel[a] += g.wall_hi[a] - nxt;
    }
}

// If the advanced particle's key left p_rect the reference drops it (3d:356-366): tombstone.
template <int DIM>
__device__ __forceinline__ bool left_p_rect(const Geo& g, const float* pos) {
    return classify_pos<DIM>(g, pos) == CLS_LIMBO;
}

// update_grid + g2p (3d:249-343) for one particle: new velocity, B (C = 4B), new position.
template <int DIM, bool DET>
__device__ __forceinline__ void g2p_particle(const Geo& g, const NodeGrid& ng, const Stencil<DIM>& s,
                                             const float* mouse, float* pos, float* vel, float* B,
                                             const long long* more = nullptr) {
    constexpr int NZ = DIM == 3 ? 3 : 1;
#pragma unroll
    for (int a = 0; a < 3; ++a) vel[a] = 0.0f;
#pragma unroll
    for (int k = 0; k < 9; ++k) B[k] = 0.0f;
#pragma unroll
    for (int oz = 0; oz < NZ; ++oz)
#pragma unroll
        for (int oy = 0; oy < 3; ++oy)
#pragma unroll
            for (int ox = 0; ox < 3; ++ox) {
                float w = s.w[0][ox] * s.w[1][oy];
                if (DIM == 3) w *= s.w[2][oz];
                float d[3] = {s.d[0][ox], s.d[1][oy], DIM == 3 ? s.d[2][oz] : 0.0f};
                float4 nd = node_load<DET>(ng, node_index<DIM>(g, s, ox, oy, oz), more);
                float nv[3] = {nd.x, nd.y, nd.z};
                if (nd.w > 0.0f) {   // update_grid (3d:253-256)
#pragma unroll
                    for (int a = 0; a < DIM; ++a) nv[a] = __fdiv_rn(nv[a], nd.w) + g.dtg[a];
                }
#pragma unroll
                for (int r = 0; r < DIM; ++r) {
                    float wv = nv[r] * w;
                    vel[r] += wv;
#pragma unroll
                    for (int c = 0; c < DIM; ++c) B[DIM * c + r] += wv * d[c];
                }
            }
    integrate_particle<DIM>(g, pos, vel, mouse);
    if (left_p_rect<DIM>(g, pos)) pos[0] = __int_as_float(0x7f800000);   // dropped: tombstone
}

// ---- one kernel per phase (any particle count) ----------------------------------------------------

template <int DIM, bool DET>
__global__ void __launch_bounds__(128)
k_p2g1_generic(const __grid_constant__ Geo g, Particles q, const int* __restrict__ src,
               const int* __restrict__ n_deposit, NodeGrid ng) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= *n_deposit) return;
    const int i = src[d];
    float4 p = q.P[i];
    float4 v = q.V[i];
    float pos[3] = {p.x, p.y, p.z};
    float vel[3] = {v.x, v.y, v.z};
    float C[9];
    load_C<DIM>(q, i, C);
    Stencil<DIM> s;
    make_stencil<DIM>(g, pos, s);
    p2g1_particle<DIM, DET>(g, ng, s, vel, C, p.w);
}

template <int DIM, bool DET>
__global__ void __launch_bounds__(128)
k_p2g2_generic(const __grid_constant__ Geo g, Particles q, const int* __restrict__ src,
               const int* __restrict__ n_deposit, NodeGrid ng, NodeGrid ng_out,
               float* __restrict__ dbg_density, float* __restrict__ dbg_pressure) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= *n_deposit) return;
    const int i = src[d];
    float4 p = q.P[i];
    float pos[3] = {p.x, p.y, p.z};
    float C[9];
    load_C<DIM>(q, i, C);
    Stencil<DIM> s;
    make_stencil<DIM>(g, pos, s);
    float density, pressure;
    p2g2_particle<DIM, DET>(g, ng, ng_out, s, C, p.w, density, pressure);
    if (dbg_density) dbg_density[d] = density;
    if (dbg_pressure) dbg_pressure[d] = pressure;
}

template <int DIM, bool DET>
__global__ void __launch_bounds__(128)
k_g2p_generic(const __grid_constant__ Geo g, Particles q, const int* __restrict__ src,
              const int* __restrict__ n_deposit, NodeGrid ng, const long long* __restrict__ more,
              const float* __restrict__ mouse) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= *n_deposit) return;
    if (mouse && mouse[2] == 0.0f) mouse = nullptr;   // {x, y, present}
    const int i = src[d];
    float4 p = q.P[i];
    float pos[3] = {p.x, p.y, p.z};
    if (classify_pos<DIM>(g, pos) != CLS_ACTIVE) return;   // g2p walks a_rect blocks only
    Stencil<DIM> s;
    make_stencil<DIM>(g, pos, s);
    float vel[3], B[9];
    g2p_particle<DIM, DET>(g, ng, s, mouse, pos, vel, B, more);
    float4 v_old = q.V[i];
    q.P[i] = make_float4(pos[0], pos[1], DIM == 3 ? pos[2] : 0.0f, p.w);
    q.V[i] = make_float4(vel[0], vel[1], DIM == 3 ? vel[2] : 0.0f, v_old.w);
    q.CA[i] = make_float4(4.0f * B[0], 4.0f * B[1], 4.0f * B[2], 4.0f * B[3]);
    if (DIM == 3) {
        q.CB[i] = make_float4(4.0f * B[4], 4.0f * B[5], 4.0f * B[6], 4.0f * B[7]);
        q.CC[i] = 4.0f * B[8];
    }
}

// ---- the whole step() in ONE cooperative launch (small scenes) -------------------------------------
//
// The reference's default scenes hold 4,096 particles (3d:524-537): a substep is a dozen kernels of a few
// microseconds each, so step() = 31 substeps is bound by launch latency and kernel tails, not by bandwidth.
// Here one thread owns one particle for the whole step(): position, velocity, C and mass stay in REGISTERS
// across all substeps, the phases are separated by grid-wide barriers (cooperative launch) instead of kernel
// boundaries, and no neighbour search runs at all (particle-per-thread deposits need no order).  The node
// grid is double buffered: substep k deposits into buffer k & 1 while every thread zeroes, in the other buffer,
// the 3^d nodes its particle touched in substep k - 1 (clear_grid, 3d:136-146: only what was written) — so a
// substep needs three barriers: p2g 1 | p2g 2 | g2p.  Both buffers are clean when the launch starts (host
// memset); the last substep always lands in buffer 0, the handle's `grid`, where read-backs expect it.
// Thread 0 stamps %globaltimer at the phase boundaries of the last substep (the reference's phase timers).
template <int DIM, bool DET>
__global__ void __launch_bounds__(128)
k_substeps_resident(const __grid_constant__ Geo g, Particles q, int n, NodeGrid g0, NodeGrid g1,
                    const float* __restrict__ mouse, int n_substeps, unsigned long long* __restrict__ stamps) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    if (mouse && mouse[2] == 0.0f) mouse = nullptr;   // {x, y, present}
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool mine = i < n;
    float pos[3] = {0.f, 0.f, 0.f}, vel[3] = {0.f, 0.f, 0.f}, C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    float m = 0.0f, idw = 0.0f;
    if (mine) {
        const float4 p = q.P[i], v = q.V[i];
        pos[0] = p.x; pos[1] = p.y; pos[2] = p.z; m = p.w;
        vel[0] = v.x; vel[1] = v.y; vel[2] = v.z; idw = v.w;
        load_C<DIM>(q, i, C);
    }
    constexpr int NZ = DIM == 3 ? 3 : 1;
    int prev_idx = -1;                                  // node (0,0,0) of the stencil deposited in the substep before
    const int first = (n_substeps - 1) & 1;             // so that the last substep uses buffer 0
    const bool clock = i == 0 && stamps != nullptr;
    for (int k = 0; k < n_substeps; ++k) {
        const bool use0 = ((first + k) & 1) == 0;
        const NodeGrid& G = use0 ? g0 : g1;
        const NodeGrid& Z = use0 ? g1 : g0;
        const bool last = k == n_substeps - 1;
        if (clock && last) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(stamps[0]));
        // clear_grid: what this particle wrote a substep ago, in the buffer nobody reads any more
        if (prev_idx >= 0) {
#pragma unroll
            for (int oz = 0; oz < NZ; ++oz)
#pragma unroll
                for (int oy = 0; oy < 3; ++oy)
#pragma unroll
                    for (int ox = 0; ox < 3; ++ox)
                        node_zero<DET>(Z, prev_idx + ox + oy * g.size[0] + (DIM == 3 ? oz * g.size[0] * g.size[1] : 0));
        }
        if (clock && last) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(stamps[1]));
        // class from the position alone, as the reference's block map has it (tombstone = dropped)
        int cls = CLS_LIMBO;
        if (mine && !is_tombstone(pos[0])) cls = classify_pos<DIM>(g, pos);
        const bool deposits = cls == CLS_ACTIVE || cls == CLS_FROZEN;   // p_rect blocks (3d:149)
        Stencil<DIM> s;
        make_stencil<DIM>(g, pos, s);
        prev_idx = deposits ? node_index<DIM>(g, s, 0, 0, 0) : -1;
        if (deposits) p2g1_particle<DIM, DET>(g, G, s, vel, C, m);
        grid.sync();
        if (clock && last) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(stamps[2]));
        if (deposits) {
            float density, pressure;
            p2g2_particle<DIM, DET>(g, G, G, s, C, m, density, pressure);
        }
        grid.sync();
        if (clock && last) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(stamps[3]));
        if (cls == CLS_ACTIVE) {                                        // a_rect blocks (3d:263)
            float B[9];
            g2p_particle<DIM, DET>(g, G, s, mouse, pos, vel, B);
#pragma unroll
            for (int c = 0; c < DIM * DIM; ++c) C[c] = 4.0f * B[c];
        }
        if (!last) grid.sync();   // every gather done before the next substep zeroes this buffer
        if (clock && last) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(stamps[4]));
    }
    if (mine) {
        q.P[i] = make_float4(pos[0], pos[1], DIM == 3 ? pos[2] : 0.0f, m);
        q.V[i] = make_float4(vel[0], vel[1], DIM == 3 ? vel[2] : 0.0f, idw);
        q.CA[i] = make_float4(C[0], C[1], C[2], C[3]);
        if (DIM == 3) {
            q.CB[i] = make_float4(C[4], C[5], C[6], C[7]);
            q.CC[i] = C[8];
        }
    }
}

// Slab runs on the particle-per-thread path (deterministic mode): after g2p, the particles whose cell left this
// rank's z slab are packed for the neighbour (17 words: 16 f32 + id, as k_g2p_tiled packs them).  They stay in
// storage as dead entries (the next sort files them under "migrated") until the host compacts them away.
__global__ void __launch_bounds__(256)
k_pack_migrants_generic(const __grid_constant__ Geo g, Particles q, const int* __restrict__ src,
                        const int* __restrict__ n_deposit, float* rec_lo, float* rec_hi, int cap, int* __restrict__ scal) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= *n_deposit) return;
    const int i = src[d];
    const float4 a = q.P[i];
    if (is_tombstone(a.x)) return;
    const int rz = rust_as_i32(floorf(a.z)) - g.org[2];
    if (rz >= g.slab_lo && rz < g.slab_hi) return;
    float pos[3] = {a.x, a.y, a.z};
    if (classify_pos<3>(g, pos) == CLS_LIMBO) return;   // left p_rect altogether: not the neighbour's either
    const int side = rz < g.slab_lo ? 0 : 1;
    float* rec = side == 0 ? rec_lo : rec_hi;
    const int slot = atomicAdd(&scal[SCAL_MIG_LO + side], 1);
    if (!rec || slot >= cap) {
        scal[SCAL_MIG_OVERFLOW] = 1;
        return;
    }
    float* r = rec + static_cast<size_t>(slot) * MIG_WORDS;
    const float4 v = q.V[i], ca = q.CA[i], cb = q.CB[i];
    r[0] = a.x; r[1] = a.y; r[2] = a.z;
    r[3] = v.x; r[4] = v.y; r[5] = v.z;
    r[6] = ca.x; r[7] = ca.y; r[8] = ca.z; r[9] = ca.w;
    r[10] = cb.x; r[11] = cb.y; r[12] = cb.z; r[13] = cb.w;
    r[14] = q.CC[i];
    r[15] = a.w;
    r[16] = v.w;   // id bits
}

// own fixed-point planes += the neighbour's (integer addition: both ranks end with the same bits, in any order)
__global__ void k_accumulate_fixed(long long* __restrict__ own, const long long* __restrict__ recv, int64_t n) {
    const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e < n) own[e] += recv[e];
}

// DET: the float view of the fixed-point grid (read-backs and parity taps read float4 records)
__global__ void k_fixed_to_float(const long long* __restrict__ fx, float4* __restrict__ f, int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const longlong2* p = reinterpret_cast<const longlong2*>(fx + 4 * i);
    const longlong2 a = p[0], b = p[1];
    f[i] = make_float4(from_fixed(a.x), from_fixed(a.y), from_fixed(b.x), from_fixed(b.y));
}

// Node grid in the reference's layout for read-back: vel[dim] then mass, velocities updated
// where mass > 0 (what update_grid leaves behind, 3d:249-259).
template <int DIM>
__global__ void k_export_grid(const __grid_constant__ Geo g, const float4* __restrict__ grid,
                              int n_nodes, float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    float4 nd = grid[g.guard + i];
    float nv[3] = {nd.x, nd.y, nd.z};
    if (nd.w > 0.0f)
        for (int a = 0; a < DIM; ++a) nv[a] = __fdiv_rn(nv[a], nd.w) + g.dtg[a];
    for (int a = 0; a < DIM; ++a) out[i * (DIM + 1) + a] = nv[a];
    out[i * (DIM + 1) + DIM] = nd.w;
}

}  // namespace fluid
