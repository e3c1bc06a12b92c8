// phases_tiled2d.cuh — the 2D hot path (2d_multi.rs:148-359) on warp-private shared-memory node tiles, sm_100a.
//
// The 2D twin of phases_tiled.cuh.  One warp owns one tile of 8 x 8 cells and the 10 x 10 nodes its particles can
// touch (3^2 stencil reach, 2d:157).  The sort is the same code as in 3D (sort.cuh, ORDER_CLASS_RR; a 2D cell is a
// "column" whose depth slot 0 alone is used): the tile's particles come in windows of <= 32 in which no two share
// a cell, balanced over the eight 16-byte bank groups (x + 2y) mod 8 of the float4 node tile (index x + 10y).  For a
// fixed stencil offset (ox, oy) the lanes of a window touch 32 different nodes, so the read-modify-writes are plain
// LDS / FFMA / STS with one __syncwarp per offset: no atomics, no conflict passes.  Particle streams are gathered
// through src[]; g2p writes the advanced particles at their sorted slots of the other buffer and counts them for the
// next substep's neighbour search; the kernels clear the node blocks they own on the way.
//
// Node record: float4 {momentum.x, momentum.y, 0, mass} as in the particle-per-thread 2D kernels.
// Phase mapping (2d:110-134):
//   k_mass_tiled2   "p2g 1"  node mass m_i = sum_p w_ip m_p                                   (2d:164,174)
//   k_p2g_tiled2    "p2g 2"  density, Tait pressure, stress (2d:195-218) and ONE scatter of
//                            w_ip (m v + (m C + T)(x_i - x_p)) = p2g_1's momentum (2d:162-175) + p2g_2's force (2d:233)
//   k_g2p_tiled2    "update" + "g2p": v_i = mom / m + dt g while loading the tile (2d:244-246), gather, C = 4B,
//                            advect, mouse, clamp, soft wall (2d:258-325), counting for the next sort
#pragma once

#include "common.cuh"
#include "phases_generic.cuh"
#include "phases_tiled.cuh"
#include "sort.cuh"

namespace fluid {

struct T2 {
    static constexpr int X = Tile<2>::X, Y = Tile<2>::Y;
    static constexpr int NX = X + 2, NY = Y + 2;
    static constexpr int NODES = NX * NY;            // 100 footprint nodes
    static constexpr int SLOTS = 104;                // float4 slots per tile (13 x 128 bytes)
    static constexpr int WARPS = 4;
    static constexpr int THREADS = WARPS * 32;
};
constexpr int FOOT2_ITERS = (T2::NODES + 31) / 32;   // 4 footprint nodes per lane

struct Tile2Ctx {
    int c0[2];     // first cell of the tile, relative to the grid origin
    int tile, base, count, windows, per, extra;
    bool edge;     // the footprint sticks out of the p_rect grid
};

__device__ __forceinline__ void tile2_from_list(const Geo& g, const int4 e, Tile2Ctx& tc) {
    tc.tile = e.x;
    tc.base = e.y;
    tc.count = e.z;
    tc.windows = e.w;
    tc.per = e.z / e.w;
    tc.extra = e.z - tc.per * e.w;
    const int tx = e.x % g.tdim[0], ty = e.x / g.tdim[0];
    tc.c0[0] = tx * T2::X;
    tc.c0[1] = ty * T2::Y;
    tc.edge = tc.c0[0] == 0 || tc.c0[1] == 0 || tc.c0[0] + T2::X + 1 > g.size[0] || tc.c0[1] + T2::Y + 1 > g.size[1];
}
__device__ __forceinline__ void window2_range(const Tile2Ctx& tc, int w, int& off, int& len) {
    off = w * tc.per + min(w, tc.extra);
    len = w < tc.windows ? tc.per + (w < tc.extra ? 1 : 0) : 0;
}
// footprint node k (0..99): global node index, -1 outside the grid
__device__ __forceinline__ int foot2_global(const Geo& g, const Tile2Ctx& tc, int k) {
    const int ly = k / T2::NX, lx = k - ly * T2::NX;
    const int x = tc.c0[0] - 1 + lx, y = tc.c0[1] - 1 + ly;
    if (k >= T2::NODES || x < 0 || y < 0 || x >= g.size[0] || y >= g.size[1]) return -1;
    return g.guard + x + y * g.size[0];
}
// the tile's own 8 x 8 node block of `arr` := 0 (clear_grid for the nodes this tile owns)
template <typename T>
__device__ __forceinline__ void zero_own_block2(const Geo& g, const Tile2Ctx& tc, int lane, T* __restrict__ arr) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int x = tc.c0[0] + (lane & 7), y = tc.c0[1] + (lane >> 3) + 4 * j;
        if (x < g.size[0] && y < g.size[1]) arr[g.guard + x + y * g.size[0]] = T{};
    }
}

struct Stencil2 {
    float wx[3], wy[3];   // zeroed outside the p_rect grid (2d:166-167)
    float cx, cy;         // pos - (cell + 0.5)
    int node0;            // tile slot of stencil offset (0,0)
};
__device__ __forceinline__ void tile2_stencil(const Geo& g, const Tile2Ctx& tc, float px, float py, Stencil2& s) {
    const float fx = floorf(px), fy = floorf(py);
    s.cx = px - (fx + 0.5f);
    s.cy = py - (fy + 0.5f);
    axis_weights(s.cx, s.wx);
    axis_weights(s.cy, s.wy);
    const int rx = rust_as_i32(fx) - g.org[0], ry = rust_as_i32(fy) - g.org[1];
    if (tc.edge) {
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            const int nx = rx - 1 + o, ny = ry - 1 + o;
            if (nx < 0 || nx >= g.size[0]) s.wx[o] = 0.0f;
            if (ny < 0 || ny >= g.size[1]) s.wy[o] = 0.0f;
        }
    }
    const int lx = min(max(rx - tc.c0[0], 0), T2::X - 1), ly = min(max(ry - tc.c0[1], 0), T2::Y - 1);
    s.node0 = lx + T2::NX * ly;
}

// ---- p2g 1: node masses -----------------------------------------------------------------------------
__global__ void __launch_bounds__(T2::THREADS)
k_mass_tiled2(const __grid_constant__ Geo g, const float4* __restrict__ P, const int* __restrict__ src,
              const int4* __restrict__ tiles, const int* __restrict__ n_active, float* __restrict__ gmass,
              float4* __restrict__ grid) {
    __shared__ float sm[T2::WARPS * T2::SLOTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* tile = sm + warp * T2::SLOTS;
    const int n_act = *n_active;
    const int n_warps = gridDim.x * T2::WARPS;
    for (int a = blockIdx.x * T2::WARPS + warp; a < n_act; a += n_warps) {
        Tile2Ctx tc;
        tile2_from_list(g, __ldg(&tiles[a]), tc);
        if (tc.count == 0) continue;   // a pseudo tile (ignored / dropped particles)
        for (int k = lane; k < T2::SLOTS; k += 32) tile[k] = 0.0f;
        __syncwarp();
        zero_own_block2(g, tc, lane, grid);   // nothing touches `grid` in this kernel; "p2g 2" deposits into it next
        for (int w = 0; w < tc.windows; ++w) {
            int off, len;
            window2_range(tc, w, off, len);
            const bool active = lane < len;
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active) p = __ldg(&P[__ldg(&src[tc.base + off + lane])]);
            Stencil2 s;
            tile2_stencil(g, tc, p.x, p.y, s);
            const int n0 = active ? s.node0 : 0;
#pragma unroll
            for (int oy = 0; oy < 3; ++oy)
#pragma unroll
                for (int ox = 0; ox < 3; ++ox) {
                    float* nd = tile + n0 + ox + T2::NX * oy;   // this lane's node for this offset: no other lane's
                    const float v = *nd + s.wx[ox] * s.wy[oy] * p.w;
                    if (active) *nd = v;
                    __syncwarp();
                }
        }
#pragma unroll
        for (int it = 0; it < FOOT2_ITERS; ++it) {
            const int k = lane + 32 * it;
            const int gi = foot2_global(g, tc, k);
            if (gi >= 0) {
                const float v = tile[k];
                if (v != 0.0f) atomicAdd(&gmass[gi], v);
            }
        }
        __syncwarp();
    }
}

// ---- p2g 2: density, pressure, stress; fused momentum + force scatter -------------------------------------
__global__ void __launch_bounds__(T2::THREADS)
k_p2g_tiled2(const __grid_constant__ Geo g, Particles q, const int* __restrict__ src, const int4* __restrict__ tiles,
             const int* __restrict__ n_active, const float* __restrict__ gmass, float4* __restrict__ grid,
             float* __restrict__ dbg_density, float* __restrict__ dbg_pressure) {
    __shared__ __align__(16) float4 sm_acc[T2::WARPS * T2::SLOTS];
    __shared__ float sm_mass[T2::WARPS * T2::SLOTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* acc = sm_acc + warp * T2::SLOTS;
    float* ms = sm_mass + warp * T2::SLOTS;
    const int n_act = *n_active;
    const int n_warps = gridDim.x * T2::WARPS;
    for (int a = blockIdx.x * T2::WARPS + warp; a < n_act; a += n_warps) {
        Tile2Ctx tc;
        tile2_from_list(g, __ldg(&tiles[a]), tc);
        if (tc.count == 0) continue;
#pragma unroll
        for (int it = 0; it < FOOT2_ITERS; ++it) {   // complete node masses of the footprint, accumulators := 0
            const int k = lane + 32 * it;
            const int gi = foot2_global(g, tc, k);
            if (k < T2::SLOTS) {
                ms[k] = gi >= 0 ? __ldg(&gmass[gi]) : 0.0f;
                acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        __syncwarp();
        for (int w = 0; w < tc.windows; ++w) {
            int off, len;
            window2_range(tc, w, off, len);
            const bool active = lane < len;
            const int d = tc.base + off + lane;   // sorted slot
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f), v = p, ca = p;
            if (active) {
                const int i = __ldg(&src[d]);
                p = __ldg(&q.P[i]);
                v = __ldg(&q.V[i]);
                ca = __ldg(&q.CA[i]);
            }
            Stencil2 s;
            tile2_stencil(g, tc, p.x, p.y, s);
            const int n0 = active ? s.node0 : 0;
            // density = sum_i m_i w_ip (2d:195-209)
            float density = 0.0f;
#pragma unroll
            for (int oy = 0; oy < 3; ++oy) {
                const float* row = ms + n0 + T2::NX * oy;
                density += (row[0] * s.wx[0] + row[1] * s.wx[1] + row[2] * s.wx[2]) * s.wy[oy];
            }
            const float m = p.w;
            float volume = 0.0f, pressure = 0.0f;
            if (active) {
                volume = m * __frcp_rn(density);
                pressure = tait_pressure_fast(g, density);
                if (dbg_density) dbg_density[d] = density;
                if (dbg_pressure) dbg_pressure[d] = pressure;
            }
            // M = m C + T,  T = -4 V (-p I + mu (C + C^T)) dt  (2d:216-219), column-major 2x2
            const float C[4] = {ca.x, ca.y, ca.z, ca.w};
            const float s1 = -4.0f * volume * g.dt;
            float M[4];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    float stress = g.mu * (C[2 * c + r] + C[2 * r + c]);
                    if (c == r) stress -= pressure;
                    M[2 * c + r] = m * C[2 * c + r] + s1 * stress;
                }
            // value at offset o: w (b + ox M col0 + oy M col1),  b = m v + M (-1 - c)
            const float dx0 = -1.0f - s.cx, dy0 = -1.0f - s.cy;
            const float b0 = m * v.x + M[0] * dx0 + M[2] * dy0, b1 = m * v.y + M[1] * dx0 + M[3] * dy0;
#pragma unroll
            for (int oy = 0; oy < 3; ++oy) {
                const float r0 = b0 + oy * M[2], r1 = b1 + oy * M[3];
#pragma unroll
                for (int ox = 0; ox < 3; ++ox) {
                    const float w = s.wx[ox] * s.wy[oy];
                    float4* nd = acc + n0 + ox + T2::NX * oy;
                    float4 a4 = *nd;
                    a4.x += w * (r0 + ox * M[0]);
                    a4.y += w * (r1 + ox * M[1]);
                    a4.w += w * m;
                    if (active) *nd = a4;
                    __syncwarp();
                }
            }
        }
#pragma unroll
        for (int it = 0; it < FOOT2_ITERS; ++it) {
            const int k = lane + 32 * it;
            const int gi = foot2_global(g, tc, k);
            if (gi >= 0) {
                const float4 v4 = acc[k];
                if (v4.w != 0.0f || v4.x != 0.0f || v4.y != 0.0f) atomicAdd(&grid[gi], v4);
            }
        }
        __syncwarp();
    }
}

// ---- update + g2p ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(T2::THREADS)
k_g2p_tiled2(const __grid_constant__ Geo g, Particles q, Particles qn, const int* __restrict__ src,
             const int4* __restrict__ tiles, const int* __restrict__ n_active, const float4* __restrict__ grid,
             const float* __restrict__ mouse, SortTables st, float* __restrict__ gmass, int* __restrict__ gz,
             const int* __restrict__ epoch_dev) {
    __shared__ __align__(16) float4 sm[T2::WARPS * T2::SLOTS];
    __shared__ int scnt_all[T2::WARPS * TILE_CELLS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* vt = sm + warp * T2::SLOTS;
    int* scnt = scnt_all + warp * TILE_CELLS;
    const int epoch = *epoch_dev + 1;
    if (mouse && mouse[2] == 0.0f) mouse = nullptr;   // {x, y, present}
    const int n_act = *n_active;
    const int n_warps = gridDim.x * T2::WARPS;
    for (int a = blockIdx.x * T2::WARPS + warp; a < n_act; a += n_warps) {
        Tile2Ctx tc;
        tile2_from_list(g, __ldg(&tiles[a]), tc);
        if (tc.count == 0) {
            if (lane == 0) st.imm_cnt[a] = 0;
            continue;
        }
#pragma unroll
        for (int j = 0; j < TILE_CELLS / 32; ++j) scnt[lane + 32 * j] = 0;
        // clear_grid for the node masses of this tile's own block, one substep ahead (nothing reads `gmass` between
        // "p2g 2" and the next "p2g 1"); the stamp tells k_clear_tiles to skip the block
        zero_own_block2(g, tc, lane, gmass);
        if (lane == 0) gz[tc.tile] = epoch;
        // footprint node records; update_grid in place: v = mom / m + dt g where m > 0 (2d:244-246)
#pragma unroll
        for (int it = 0; it < FOOT2_ITERS; ++it) {
            const int k = lane + 32 * it;
            const int gi = foot2_global(g, tc, k);
            if (k < T2::SLOTS) {
                float4 nd = gi >= 0 ? __ldg(&grid[gi]) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (nd.w > 0.0f) {
                    const float inv = __frcp_rn(nd.w);
                    nd.x = nd.x * inv + g.dtg[0];
                    nd.y = nd.y * inv + g.dtg[1];
                }
                vt[k] = nd;
            }
        }
        __syncwarp();
        int n_leave = 0;
        for (int it = 0; it < tc.count; it += 32) {
            const bool active = it + lane < tc.count;
            const int d = tc.base + it + lane;   // sorted slot = index in the new buffer
            int i = 0;
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f), v_old = p;
            if (active) {
                i = __ldg(&src[d]);
                p = __ldg(&q.P[i]);
                v_old = __ldg(&q.V[i]);
            }
            float pos[3] = {p.x, p.y, 0.0f};
            const bool advance = active && classify_pos<2>(g, pos) == CLS_ACTIVE;   // g2p walks a_rect blocks only
            if (advance) {
                Stencil2 s;
                tile2_stencil(g, tc, p.x, p.y, s);
                float vel[3] = {0.f, 0.f, 0.f}, B[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int oy = 0; oy < 3; ++oy)
#pragma unroll
                    for (int ox = 0; ox < 3; ++ox) {
                        const float4 nd = vt[s.node0 + ox + T2::NX * oy];
                        const float w = s.wx[ox] * s.wy[oy];
                        const float dx = (ox - 1) - s.cx, dy = (oy - 1) - s.cy;
                        const float wvx = nd.x * w, wvy = nd.y * w;
                        vel[0] += wvx;
                        vel[1] += wvy;
                        B[0] += wvx * dx;   // B[2c + r] += (w v_r) d_c  (2d:276-281)
                        B[1] += wvy * dx;
                        B[2] += wvx * dy;
                        B[3] += wvy * dy;
                    }
                integrate_particle<2>(g, pos, vel, mouse);
                if (left_p_rect<2>(g, pos)) pos[0] = __int_as_float(0x7f800000);   // dropped: tombstone
                qn.P[d] = make_float4(pos[0], pos[1], 0.0f, p.w);
                qn.V[d] = make_float4(vel[0], vel[1], 0.0f, v_old.w);
                qn.CA[d] = make_float4(4.0f * B[0], 4.0f * B[1], 4.0f * B[2], 4.0f * B[3]);
            } else if (active) {   // frozen halo particle: carried over unchanged (2d:149 vs 2d:254)
                qn.P[d] = p;
                qn.V[d] = v_old;
                qn.CA[d] = __ldg(&q.CA[i]);
            }
            // start the next substep's neighbour search: bucket of the (possibly moved) particle; the ones that
            // stay in this tile are ranked with shared-memory integer atomics, the few that leave are listed
            int cls = -1, bucket = 0;
            if (active) bucket = bucket_of<2>(g, make_float4(pos[0], pos[1], 0.f, 0.f), cls);
            const bool stays = active && (bucket >> 8) == tc.tile;
            const bool leaves = active && !stays;
            if (active) st.gcell[d] = bucket;
            if (stays) st.rank[d] = atomicAdd(&scnt[bucket & (TILE_CELLS - 1)], 1);
            const unsigned lm = __ballot_sync(0xffffffffu, leaves);
            if (leaves) st.imm_list[tc.base + n_leave + __popc(lm & ((1u << lane) - 1u))] = d;
            n_leave += __popc(lm);
        }
        __syncwarp();
        int total = 0;
#pragma unroll
        for (int j = 0; j < TILE_CELLS / 32; ++j) {
            const int c = scnt[lane + 32 * j];
            st.count[tc.tile * TILE_CELLS + lane + 32 * j] = c;
            total += c;
        }
        total = __reduce_add_sync(0xffffffffu, total);
        if (lane == 0) {
            st.tile_total[tc.tile] = total;
            st.imm_cnt[a] = n_leave;
        }
        __syncwarp();
    }
}

}  // namespace fluid
