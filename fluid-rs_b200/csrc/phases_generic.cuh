// phases_generic.cuh — the five reference phases as straightforward particle-per-thread kernels.
//
// This is the GENERIC path: any dim (2 or 3), any grid_res, any particle density.  It scatters
// with vector float atomics (red.global.add.v4.f32 -> SASS REDG.E.ADD.F32x4) straight into the
// dense node grid and gathers from it through L2.  The tiled sm_100a path in phases_tiled.cuh
// replaces it for 3D scenes; this one stays as the 2D path and as the fallback for shapes the
// tiled kernels do not cover.
//
// Node record in HBM: float4 {momentum.xyz (velocity after update), mass} = `struct Cell`
// (3d:43-48) without `is_computed`: update_grid is folded into g2p's node read, which applies
// `v = mom / mass + dt*g` exactly where the reference's guard `mass > 0` holds (3d:253-256).
#pragma once

#include "common.cuh"
#include "sort.cuh"

namespace fluid {

template <int DIM>
__device__ __forceinline__ int node_index(const Geo& g, const Stencil<DIM>& s, int ox, int oy,
                                          int oz) {
    int idx = g.guard + (s.base[0] + ox) + (s.base[1] + oy) * g.size[0];
    if (DIM == 3) idx += (s.base[2] + oz) * g.size[0] * g.size[1];
    return idx;
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

// p2g_1 (3d:148-183): node.mass += w*m ; node.mom += w*m*(v + C*(x_n - x_p))
template <int DIM>
__global__ void __launch_bounds__(128)
k_p2g1_generic(const __grid_constant__ Geo g, Particles q, const int* __restrict__ src,
               const int* __restrict__ n_deposit, float4* __restrict__ grid) {
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
    const float m = p.w;
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
                atomicAdd(&grid[node_index<DIM>(g, s, ox, oy, oz)],
                          make_float4(mom[0], mom[1], mom[2], mc));
            }
}

// p2g_2 (3d:185-247): density from node masses, Tait pressure, stress, force scatter.
template <int DIM>
__global__ void __launch_bounds__(128)
k_p2g2_generic(const __grid_constant__ Geo g, Particles q, const int* __restrict__ src,
               const int* __restrict__ n_deposit, float4* __restrict__ grid,
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
    constexpr int NZ = DIM == 3 ? 3 : 1;

    float density = 0.0f;
#pragma unroll
    for (int oz = 0; oz < NZ; ++oz)
#pragma unroll
        for (int oy = 0; oy < 3; ++oy)
#pragma unroll
            for (int ox = 0; ox < 3; ++ox) {
                float w = s.w[0][ox] * s.w[1][oy];
                if (DIM == 3) w *= s.w[2][oz];
                // .w lane only: the momentum lanes are being written by other threads
                const float* node =
                    reinterpret_cast<const float*>(&grid[node_index<DIM>(g, s, ox, oy, oz)]);
                density += __ldcg(node + 3) * w;
            }
    float volume = __fdiv_rn(p.w, density);
    float pressure = tait_pressure(g, density);
    if (dbg_density) dbg_density[d] = density;
    if (dbg_pressure) dbg_pressure[d] = pressure;

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
                atomicAdd(&grid[node_index<DIM>(g, s, ox, oy, oz)],
                          make_float4(f[0], f[1], f[2], 0.0f));
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

// update_grid + g2p (3d:249-381)
template <int DIM>
__global__ void __launch_bounds__(128)
k_g2p_generic(const __grid_constant__ Geo g, Particles q, const int* __restrict__ src,
              const int* __restrict__ n_deposit, const float4* __restrict__ grid,
              const float* __restrict__ mouse) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= *n_deposit) return;
    const int i = src[d];
    float4 p = q.P[i];
    float pos[3] = {p.x, p.y, p.z};
    if (classify_pos<DIM>(g, pos) != CLS_ACTIVE) return;   // g2p walks a_rect blocks only
    Stencil<DIM> s;
    make_stencil<DIM>(g, pos, s);
    constexpr int NZ = DIM == 3 ? 3 : 1;
    float vel[3] = {0.0f, 0.0f, 0.0f};
    float B[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int oz = 0; oz < NZ; ++oz)
#pragma unroll
        for (int oy = 0; oy < 3; ++oy)
#pragma unroll
            for (int ox = 0; ox < 3; ++ox) {
                float w = s.w[0][ox] * s.w[1][oy];
                if (DIM == 3) w *= s.w[2][oz];
                float d[3] = {s.d[0][ox], s.d[1][oy], DIM == 3 ? s.d[2][oz] : 0.0f};
                float4 nd = __ldg(&grid[node_index<DIM>(g, s, ox, oy, oz)]);
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
    float4 v_old = q.V[i];
    q.P[i] = make_float4(pos[0], pos[1], DIM == 3 ? pos[2] : 0.0f, p.w);
    q.V[i] = make_float4(vel[0], vel[1], DIM == 3 ? vel[2] : 0.0f, v_old.w);
    q.CA[i] = make_float4(4.0f * B[0], 4.0f * B[1], 4.0f * B[2], 4.0f * B[3]);
    if (DIM == 3) {
        q.CB[i] = make_float4(4.0f * B[4], 4.0f * B[5], 4.0f * B[6], 4.0f * B[7]);
        q.CC[i] = 4.0f * B[8];
    }
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
