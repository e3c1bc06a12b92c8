// phases_tiled.cuh — the 3D hot path: warp-private shared-memory node tiles, sm_100a.
//
// One warp owns one tile of 8x8x4 cells (Tile<3>) and the 10x10x6 nodes its particles can touch
// (3^3 stencil reach, 3d:157-158).  The sort (sort.cuh, ORDER_CLASS_RR or ORDER_CLASS_Q) deals the tile's
// particles into windows of <= 32 in which no two particles share an (x,y) column.  A warp
// processes one window per iteration, one particle per lane: for a fixed stencil offset (ox,oy)
// the 32 lanes touch 32 different node columns and the three nodes along z belong to the lane
// alone, so the read-modify-writes into the shared-memory tile are plain LDS / FFMA / STS — three
// independent chains in flight per lane, one warp barrier per (ox,oy) — with no atomics (shared
// float atomics are a CAS loop on sm_100a, ATOMS.CAST.SPIN) and no conflict passes.  Lanes without a
// particle skip the window's shared-memory work (a quarter warp none of whose lanes takes part in a
// 128-bit access costs no wavefront).  The tiles of the active-tile list are handed to the resident warps
// through a ticket counter (k_p2g_tiled, k_g2p_tiled), so a warp that drew light tiles takes more of them.
// Tiles move between HBM and shared memory through the tensor memory accelerator where their
// footprint lies inside the grid: k_p2g_tiled flushes its accumulators with six tensor reductions
// (cp.reduce.async.bulk.tensor.4d .add, SASS UTMAREDG.4D.ADD), k_g2p_tiled loads its footprint with
// six tensor copies (cp.async.bulk.tensor.4d, SASS UTMALDG.4D, mbarrier completion); tiles on the rim
// of the grid and the node masses use vector / scalar reductions (red.global.add, SASS REDG) and
// cp.async.  In slab runs (PEER) the deposits into the node planes shared with a neighbour GPU also go
// into the neighbour's mapped arrays (red.relaxed.sys).  Particle streams are gathered through src[]
// (sorted slot -> storage index); g2p writes the advanced particles at their sorted slots of the other
// buffer, so no separate reorder pass exists.  The tile kernels also clear the grid on the way
// (k_mass_tiled the node records, k_g2p_tiled the node masses of the blocks they own).
//
// Shared-memory node index: x + 10*y + 104*z.  The plane stride is padded from 100 to 104
// (= 0 mod 8, and 104 nodes = 13 * 128 bytes for the TMA boxes) so that the 16-byte bank group of a
// float4 node, (x + 2y) mod 8, depends on the column only; the sort enumerates columns so that
// neighbouring lanes fall into different groups.  Node masses are kept as z quads (T3::QSLOTS).
//
// Phase mapping to the reference (3d:110-134):
//   k_mass_tiled  "p2g 1"  node mass   m_i  = sum_p w_ip m_p                      (3d:164,175)
//   k_p2g_tiled   "p2g 2"  density, Tait pressure, stress (3d:198-225) and ONE scatter of
//                          w_ip * (m v + (m C + T) (x_i - x_p)) = the momentum of p2g_1
//                          (3d:163,176) plus the force term of p2g_2 (3d:242); the w lane of
//                          the float4 node carries w_ip m_p again so g2p reads one record
//   k_g2p_tiled   "update" + "g2p": v_i = mom/m + dt g while loading the tile (3d:253-256),
//                          gather, C = 4B, advect, mouse, clamp, soft wall (3d:267-343); it also
//                          counts the particles for the next substep's neighbour search
#pragma once

#include "common.cuh"
#include "phases_generic.cuh"
#include "sort.cuh"

#include <cuda.h>   // CUtensorMap (the descriptor is built on the host in fluid_abi.cu; no driver API is linked)

namespace fluid {

constexpr int T3_NX_NY_BYTES = 10 * 10 * 16;   // one footprint plane of float4 nodes
struct T3 {
    static constexpr int X = Tile<3>::X, Y = Tile<3>::Y, Z = Tile<3>::Z;
    static constexpr int NX = X + 2, NY = Y + 2, NZ = Z + 2;
    static constexpr int NODES = NX * NY * NZ;       // 600 footprint nodes
    static constexpr int PLANE = 104;                // padded z stride in shared memory
    static constexpr int SLOTS = PLANE * NZ;         // 624 shared-memory slots per tile
    // Node-MASS tiles pack the z axis: column (x,y) holds two overlapping quads of nodes along z,
    // Q0 = z 0..3 and Q1 = z 2..5, one float4 each at x + 10*y + 104*Q (the float4 tile's indexing with Q
    // as the plane, so the same bank-group rule holds).  A particle in cell layer lz touches nodes
    // lz..lz+2: all inside quad lz >> 1, so one 128-bit access per stencil column replaces three 32-bit
    // ones (which ran at 2.0 wavefronts each under the class-major lane order).  Nodes z = 2,3 live in
    // both quads: a deposit goes to one of them and the flush adds the two.
    static constexpr int QSLOTS = PLANE * 2;         // 208 float4 per mass tile
    static constexpr int WARPS = 4;                  // tiles per CTA (no CTA-level sync is used)
    static constexpr int THREADS = WARPS * 32;
};
constexpr int FOOT_ITERS = (T3::NODES + 31) / 32;    // 19 footprint nodes per lane
constexpr int FOOT_ROWS = T3::NY * T3::NZ;           // 60 rows of 10 nodes
constexpr int FOOT_STEPS = FOOT_ROWS / 3;            // a warp takes 3 rows (30 lanes) per step

// ---- TMA (cp.async.bulk.tensor) helpers: the node grid as a 4-D tensor {4 floats, x, y, z}; one box =
// one 10x10 plane of a tile's footprint, dense in shared memory (1600 B at a 128-byte aligned address).
// The copies do not go through the LSU pipe, which is what binds the tile kernels.  What a box may do at the rim of
// the tensor was measured on B200 (tools/tma_probe.cu, profiles/r02_tma_probe.log): LOADS take any start
// coordinate, negative ones included, and zero-fill what lies outside; a box may overhang the far end for loads
// and reductions alike; but a REDUCTION whose box starts at a negative coordinate, and any box whose negative
// start is not 16-byte aligned (the 4-byte node-mass tensor at x = -1), raise "illegal instruction".  So g2p
// loads every tile's footprint by TMA, and p2g flushes by TMA unless the tile touches the low rim of the grid.
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, int parity) {
    const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_plane(void* smem_dst, const CUtensorMap* tm, int x, int y, int z, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(reinterpret_cast<unsigned long long>(tm)), "r"(0),
                 "r"(x), "r"(y), "r"(z), "r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_plane(const CUtensorMap* tm, int x, int y, int z, const void* smem_src) {
    asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                 ::"l"(reinterpret_cast<unsigned long long>(tm)), "r"(0), "r"(x), "r"(y), "r"(z),
                 "r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_src))) : "memory");
}
__device__ __forceinline__ void tma_load_box3(void* smem_dst, const CUtensorMap* tm, int x, int y, int z, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(reinterpret_cast<unsigned long long>(tm)),
                 "r"(x), "r"(y), "r"(z), "r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
constexpr int FOOT_PLANE_BYTES = T3_NX_NY_BYTES;
// node-mass box of the TMA footprint load: rows of 12 floats (the inner extent of a box has to be a multiple
// of 16 bytes), 10 rows, 6 planes
constexpr int MBOX_X = 12, MBOX_FLOATS = MBOX_X * 10 * 6;

struct TileCtx {
    int c0[3];     // first cell of the tile, relative to the grid origin
    int tile;      // tile id
    int base;      // first particle slot
    int count;     // particles in the tile
    int windows;   // W
    int per, extra;
    bool quart;          // ORDER_CLASS_Q (sort.cuh): a window's rounds sit on quarter-warp boundaries
    unsigned nlo, nhi;   // ... its class totals N_b, one byte each (classes 0..3 | 4..7)
    bool edge;     // the footprint sticks out of the p_rect grid
    bool low_rim;  // ... at the low end of an axis: its boxes start at coordinate -1 (no TMA reduction there)
};

// Active tiles, listed by k_tile_tables: {tile id, first slot, count, windows}.  The tile kernels
// run persistent warps that stride over this list, so empty tiles cost nothing.
__device__ __forceinline__ void tile_from_list(const Geo& g, const int4 e, TileCtx& tc, const unsigned char* __restrict__ tab = nullptr,
                                               int a = 0) {
    const int t = e.x;
    tc.tile = t;
    tc.base = e.y;
    tc.count = e.z;
    tc.quart = (e.w & ORDER_Q_FLAG) != 0;
    tc.windows = e.w & (ORDER_Q_FLAG - 1);
    tc.per = e.z / tc.windows;
    tc.extra = e.z - tc.per * tc.windows;
    tc.nlo = tc.nhi = 0u;
    if (tc.quart && tab) {   // (the kernels that walk windows pass the table; g2p's compact walk does not need it)
        const uint2 nb = __ldg(reinterpret_cast<const uint2*>(tab + static_cast<size_t>(a) * TAB_BYTES));
        tc.nlo = nb.x;
        tc.nhi = nb.y;
    }
    int tx = t % g.tdim[0];
    int r = t / g.tdim[0];
    int ty = r % g.tdim[1];
    int tz = r / g.tdim[1];
    tc.c0[0] = tx * T3::X;
    tc.c0[1] = ty * T3::Y;
    tc.c0[2] = tz * T3::Z;
    tc.low_rim = tc.c0[0] == 0 || tc.c0[1] == 0 || tc.c0[2] == 0;
    tc.edge = tc.low_rim || tc.c0[0] + T3::X + 1 > g.size[0] || tc.c0[1] + T3::Y + 1 > g.size[1] ||
              tc.c0[2] + T3::Z + 1 > g.size[2];
}

// Every cell of the tile lies in an a_rect block: then every particle the sort put into the tile is advanced by g2p
// (3d:263) and the per-particle block-key classification can be skipped (the tile's cells are the particles' cells).
__device__ __forceinline__ bool tile_all_active(const Geo& g, const TileCtx& tc) {
    bool ok = true;
    const int ext[3] = {T3::X, T3::Y, T3::Z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int k0 = block_key_of_cell(tc.c0[a] + g.org[a], g.res_i, g.res_shift);
        const int k1 = block_key_of_cell(tc.c0[a] + ext[a] - 1 + g.org[a], g.res_i, g.res_shift);
        ok = ok && k0 >= g.a_lo[a] && k1 < g.a_hi[a];
    }
    return ok;
}

// 1 / x as one MUFU.RCP (1 ulp); x > 0 and far from the denormal range here (node masses)
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// floor(pos) is one of the tile's cells
__device__ __forceinline__ bool in_tile(const Geo& g, const TileCtx& tc, const float* pos) {
    const int lx = rust_as_i32(floorf(pos[0])) - g.org[0] - tc.c0[0];
    const int ly = rust_as_i32(floorf(pos[1])) - g.org[1] - tc.c0[1];
    const int lz = rust_as_i32(floorf(pos[2])) - g.org[2] - tc.c0[2];
    return static_cast<unsigned>(lx) < static_cast<unsigned>(T3::X) && static_cast<unsigned>(ly) < static_cast<unsigned>(T3::Y) &&
           static_cast<unsigned>(lz) < static_cast<unsigned>(T3::Z);
}

// The tile kernels run one wave of resident warps over the active-tile list.  With a ticket counter (zeroed by the
// sort) a warp takes the next tile when it is done with one, so a warp that drew light tiles does more of them and
// the kernel's tail is one tile long; without one the list is dealt out with a fixed stride.  The first `fixed8`
// eighths of the list (rounded to whole rounds of the resident warps) can still be dealt with the fixed stride and
// only the rest handed out by tickets; in the ticket phase the next ticket is drawn at the START of a tile and looked
// at when the tile is done.  Measured at 2^24 particles (profiles/r02_scheduling_ab.md): k_g2p_tiled -6 % with tickets
// throughout; k_p2g_tiled -3.5 % with tickets throughout, -5 % with tickets for the second half of the list;
// k_mass_tiled gets SLOWER in proportion to the ticketed share (+70 % with tickets throughout, +17 % with tickets for
// the last eighth) and keeps the fixed stride: a returning atomic in flight shares the warp's load scoreboards, and
// with its short tiles and 48 warps per SM flushing with reductions every wait for a particle fetch then waits for
// the ticket as well (ncu: long-scoreboard stalls 2.2 -> 14.5 per issue).  Tickets that cover 2 or 4 consecutive tiles
// are slower than either (g2p 0.60 / 0.66 / 0.73 ms at 1 / 2 / 4: the tail grows with the batch); listing the tiles of
// 128 or more particles first, so that the tail is made of small pieces: k_p2g_tiled +5 % (light tiles running beside
// heavy ones is what keeps the pipes mixed).
struct TileWalk {
    int a;         // current list entry
    int raw;       // lane 0: the ticket drawn for the next entry
    int n_fixed;   // entries dealt with the fixed stride
};
__device__ __forceinline__ int draw_ticket(int* __restrict__ ticket, int lane) { return lane == 0 ? atomicAdd(ticket, 1) : 0; }
__device__ __forceinline__ void walk_begin(TileWalk& w, int* __restrict__ ticket, int lane, int first, int stride, int n_act, int fixed8) {
    w.n_fixed = ticket ? ((n_act / stride) * fixed8 / 8) * stride : n_act;
    w.a = first;
    w.raw = 0;
    if (w.a >= w.n_fixed && ticket) w.a = w.n_fixed + __shfl_sync(0xffffffffu, draw_ticket(ticket, lane), 0);
}
__device__ __forceinline__ void walk_prefetch(TileWalk& w, int* __restrict__ ticket, int lane) {   // at the top of the loop body
    if (ticket && w.a >= w.n_fixed) w.raw = draw_ticket(ticket, lane);
}
__device__ __forceinline__ void walk_next(TileWalk& w, int* __restrict__ ticket, int lane, int stride) {
    if (!ticket) {
        w.a += stride;
    } else if (w.a >= w.n_fixed) {
        w.a = w.n_fixed + __shfl_sync(0xffffffffu, w.raw, 0);
    } else {
        w.a += stride;
        if (w.a >= w.n_fixed) w.a = w.n_fixed + __shfl_sync(0xffffffffu, draw_ticket(ticket, lane), 0);   // leaving the fixed part
    }
}

// A lane's place in window w of its tile: `active` and the slot (relative to the tile's first).  `first` = slots of
// the windows before w; it is advanced to window w + 1, so the windows have to be visited in order (one call per
// window, w = 0, 1, ...).
//   ORDER_CLASS_Q: quarter warp k = round k of the window = the classes with N_b > w + k W, ascending, from lane 8k on
//   otherwise:     the window is the slot range [w per + min(w, extra), + per + (w < extra)), lane by lane
struct WinLane {
    bool active;
    int slot;
};
__device__ __forceinline__ int round_members(const TileCtx& tc, int thr) {   // #{b : N_b > thr}, thr < 128
    const unsigned t4 = static_cast<unsigned>(thr) * 0x01010101u;
    return (__popc(__vcmpgtu4(tc.nlo, t4)) + __popc(__vcmpgtu4(tc.nhi, t4))) >> 3;
}
__device__ __forceinline__ WinLane window_lane(const TileCtx& tc, int w, int lane, int& first) {
    WinLane r;
    r.active = false;
    r.slot = 0;
    if (w >= tc.windows) return r;
    if (tc.quart) {
        const int m0 = round_members(tc, w), m1 = round_members(tc, w + tc.windows),
                  m2 = round_members(tc, w + 2 * tc.windows), m3 = round_members(tc, w + 3 * tc.windows);
        const int k = lane >> 3, j = lane & 7;
        const int mk = k == 0 ? m0 : (k == 1 ? m1 : (k == 2 ? m2 : m3));
        const int before = k == 0 ? 0 : (k == 1 ? m0 : (k == 2 ? m0 + m1 : m0 + m1 + m2));
        r.active = j < mk;
        r.slot = first + before + j;
        first += m0 + m1 + m2 + m3;
    } else {
        const int len = tc.per + (w < tc.extra ? 1 : 0);
        r.active = lane < len;
        r.slot = first + lane;
        first += len;
    }
    return r;
}

// Parity tap: window * 32 + lane of every sorted slot, from the same window_lane() the tile kernels walk with.
__global__ void __launch_bounds__(128)
k_debug_windows(const __grid_constant__ Geo g, const int4* __restrict__ tiles, const int* __restrict__ n_active,
                const unsigned char* __restrict__ tab, int* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; a < *n_active; a += n_warps) {
        TileCtx tc;
        tile_from_list(g, tiles[a], tc, tab, a);
        if (tc.count == 0) continue;
        int first = 0;
        for (int w = 0; w < tc.windows; ++w) {
            const WinLane wl = window_lane(tc, w, lane, first);
            if (wl.active) out[tc.base + wl.slot] = w * 32 + lane;
        }
    }
}

// Footprint node k (0..599): shared-memory slot and global node index (-1 outside the grid).
__device__ __forceinline__ int footprint_slot(int k) {
    const int lz = k / (T3::NX * T3::NY);
    return k + lz * (T3::PLANE - T3::NX * T3::NY);
}
__device__ __forceinline__ int footprint_to_global(const Geo& g, const TileCtx& tc, int k) {
    int lx = k % T3::NX;
    int r = k / T3::NX;
    int ly = r % T3::NY;
    int lz = r / T3::NY;
    int x = tc.c0[0] - 1 + lx, y = tc.c0[1] - 1 + ly, z = tc.c0[2] - 1 + lz;
    if (k >= T3::NODES || x < 0 || y < 0 || z < 0 || x >= g.size[0] || y >= g.size[1] || z >= g.size[2])
        return -1;
    return node_addr(g, x, y, z);
}

// Row-wise walk over the footprint: step `it` (0..19), lanes 0..29 -> row 3*it + lane/10, x = lane%10.
// Returns the shared-memory slot; `gi` receives the global node index (-1 outside the grid / idle lane).
struct FootLane {
    int x, rsub;   // lane % 10, lane / 10 (3 for the two idle lanes)
};
__device__ __forceinline__ FootLane foot_lane(int lane) {
    FootLane f;
    f.rsub = lane / T3::NX;
    f.x = lane - f.rsub * T3::NX;
    return f;
}
__device__ __forceinline__ int foot_step(const Geo& g, const TileCtx& tc, const FootLane& f, int it, int& gi) {
    const int r = 3 * it + f.rsub;
    const int lz = (r * 205) >> 11;          // r / 10 for r < 64
    const int ly = r - lz * T3::NY;
    gi = -1;
    if (f.rsub < 3) {
        const int x = tc.c0[0] - 1 + f.x, y = tc.c0[1] - 1 + ly, z = tc.c0[2] - 1 + lz;
        if (!tc.edge || (x >= 0 && y >= 0 && z >= 0 && x < g.size[0] && y < g.size[1] && z < g.size[2]))
            gi = node_addr(g, x, y, z);
    }
    return f.x + T3::NX * ly + T3::PLANE * lz;
}
// Per-particle stencil in tile coordinates.
struct TStencil {
    float wx[3], wy[3], wz[3];   // zeroed outside the p_rect grid (3d:166-168)
    float cx, cy, cz;            // pos - (cell + 0.5)
    int node0;                   // shared-memory slot of stencil offset (0,0,0) in a float4 tile
    int node0q;                  // stencil column (0,0) in a mass tile: x + 10*y + 104*(lz >> 1)
    float wq[4];                 // z weights laid over the quad: nodes (lz & 1) .. (lz & 1) + 2
};

__device__ __forceinline__ void axis_weights(float c, float* w) {
    float m = 0.5f - c, p = 0.5f + c;
    w[0] = 0.5f * m * m;
    w[1] = 0.75f - c * c;
    w[2] = 0.5f * p * p;
}

__device__ __forceinline__ void tile_stencil(const Geo& g, const TileCtx& tc, float px, float py,
                                             float pz, TStencil& s) {
    float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    s.cx = px - (fx + 0.5f);
    s.cy = py - (fy + 0.5f);
    s.cz = pz - (fz + 0.5f);
    axis_weights(s.cx, s.wx);
    axis_weights(s.cy, s.wy);
    axis_weights(s.cz, s.wz);
    int rx = rust_as_i32(fx) - g.org[0], ry = rust_as_i32(fy) - g.org[1], rz = rust_as_i32(fz) - g.org[2];
    if (tc.edge) {   // warp-uniform: only tiles on the rim of the grid can reach outside it
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            int nx = rx - 1 + o, ny = ry - 1 + o, nz = rz - 1 + o;
            if (nx < 0 || nx >= g.size[0]) s.wx[o] = 0.0f;
            if (ny < 0 || ny >= g.size[1]) s.wy[o] = 0.0f;
            if (nz < 0 || nz >= g.size[2]) s.wz[o] = 0.0f;
        }
    }
    // the sort put this particle into this tile from the same floor(pos); the clamp only guards
    // shared memory against non-finite positions
    int lx = min(max(rx - tc.c0[0], 0), T3::X - 1);
    int ly = min(max(ry - tc.c0[1], 0), T3::Y - 1);
    int lz = min(max(rz - tc.c0[2], 0), T3::Z - 1);
    s.node0 = lx + T3::NX * ly + T3::PLANE * lz;
    s.node0q = lx + T3::NX * ly + T3::PLANE * (lz >> 1);
    const bool hi = (lz & 1) != 0;
    s.wq[0] = hi ? 0.0f : s.wz[0];
    s.wq[1] = hi ? s.wz[0] : s.wz[1];
    s.wq[2] = hi ? s.wz[1] : s.wz[2];
    s.wq[3] = hi ? s.wz[2] : 0.0f;
}

// ---- clear: zero the 8x8x4 node blocks marked dirty by the sort (clear_grid, 3d:136-146) ----------
// A block is dirty when a tile of its 3x3x3 neighbourhood holds particles now or held some in the
// previous substep (flags of two consecutive sorts are kept: `dirty_prev | dirty_now`).
__global__ void __launch_bounds__(256)
k_dirty_list(const __grid_constant__ Geo g, const unsigned char* __restrict__ dirty_now,
             unsigned char* __restrict__ dirty_prev, int* __restrict__ list, int* __restrict__ n_list, bool reset) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    bool d = false;
    if (t < g.n_tiles) {
        const bool was = dirty_prev[t] != 0;
        d = was || dirty_now[t] != 0;
        if (was && reset) dirty_prev[t] = 0;   // dirty_prev becomes the next sort's dirty_now
    }
    const unsigned m = __ballot_sync(0xffffffffu, d);
    const int lane = threadIdx.x & 31;
    int slot = 0;
    if (lane == 0 && m) slot = atomicAdd(n_list, __popc(m));
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (d) list[slot + __popc(m & ((1u << lane) - 1u))] = t;
}

// The tile kernels zero the blocks of the tiles they own on the way (k_mass_tiled: the node records of
// the tiles active in this sort; k_g2p_tiled: the node masses of the tiles active in the previous one,
// stamped in `gz`), so this kernel is left with the rim: dirty blocks whose own tile holds no particles.
__global__ void __launch_bounds__(128)
k_clear_tiles(const __grid_constant__ Geo g, const int* __restrict__ list, const int* __restrict__ n_list,
              float4* __restrict__ grid, float* __restrict__ gmass, const int* __restrict__ tile_base,
              const int* __restrict__ gz, const int* __restrict__ epoch_dev, bool fused, int what,
              const unsigned char* __restrict__ dirty_now) {
    const int epoch_prev = *epoch_dev;   // tiled substeps completed so far = the stamp k_g2p_tiled left in the last one
    const int lane = threadIdx.x & 31;
    const int n = *n_list;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; a < n; a += n_warps) {
        const int t = list[a];
        // what: 1 = node records, 2 = node masses, 3 = both (peer-halo slab runs clear them at different times)
        const bool do_grid = (what & 1) && (!fused || tile_base[t + 1] == tile_base[t]);   // else k_mass_tiled zeroes it
        const bool do_mass = (what & 2) && (!fused || gz[t] != epoch_prev);                // else k_g2p_tiled did
        // block-sparse: a block no stencil reaches any more (dirty only from the sort before) goes back to the
        // free list, zeroed (below, or already by k_g2p_tiled)
        const bool release = g.sp.blk && dirty_now && !dirty_now[t];
        if (!do_grid && !do_mass && !release) continue;
        const int tx = t % g.tdim[0], r = t / g.tdim[0], ty = r % g.tdim[1], tz = r / g.tdim[1];
        // 256 nodes: lane -> x = lane & 7, y = (lane >> 3) + 4*(j & 1), z = j >> 1
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int x = tx * T3::X + (lane & 7);
            const int y = ty * T3::Y + (lane >> 3) + 4 * (j & 1);
            const int z = tz * T3::Z + (j >> 1);
            if (x < g.size[0] && y < g.size[1] && z < g.size[2]) {
                const int gi = node_addr(g, x, y, z);
                if (gi >= 0) {
                    if (do_grid) grid[gi] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (do_mass) gmass[gi] = 0.0f;
                }
            }
        }
        if (release) {
            __syncwarp();
            if (lane == 0) {
                const int b = g.sp.blk[t];
                if (b > 0) {   // (block 0 is the overflow block of an exhausted pool: never recycled)
                    g.sp.blk[t] = -1;
                    g.sp.free_list[atomicAdd(&g.sp.scal[0], 1)] = b;
                }
            }
        }
    }
}

// Reductions into a neighbour GPU's memory (mapped through CUDA IPC): system scope, no return value
// (SASS REDG; a plain atomicAdd on these pointers compiled to the returning ATOMG and slowed the whole flush).
__device__ __forceinline__ void red_add_sys(float4* p, const float4 v) {
    asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void red_add_sys(float* p, const float v) {
    asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// Zero the tile's own 8x8x4 node block of `arr` (the nodes with the tile's cell indices).
template <typename T>
__device__ __forceinline__ void zero_own_block(const Geo& g, const TileCtx& tc, int lane, T* __restrict__ arr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int x = tc.c0[0] + (lane & 7);
        const int y = tc.c0[1] + (lane >> 3) + 4 * (j & 1);
        const int z = tc.c0[2] + (j >> 1);
        if (x < g.size[0] && y < g.size[1] && z < g.size[2]) arr[node_addr(g, x, y, z)] = T{};
    }
}

// ---- p2g 1: node masses ---------------------------------------------------------------------

template <bool PEER>   // PEER: slab run with the neighbours' arrays mapped (deposits into the shared planes go there too)
__global__ void __launch_bounds__(T3::THREADS, 12)   // 40 registers: 12 CTAs per SM (at 48 registers / 10 CTAs: +6 % time)
k_mass_tiled(const __grid_constant__ Geo g, const float4* __restrict__ P,
             const int* __restrict__ src, const int4* __restrict__ tiles,
             const int* __restrict__ n_active, float* __restrict__ gmass, float4* __restrict__ grid, PeerHalo ph,
             ParticleTex tq, const unsigned char* __restrict__ tab, int* __restrict__ ticket, int fixed8) {
    __shared__ float4 sm[T3::WARPS * T3::QSLOTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* tile = sm + warp * T3::QSLOTS;
    const int n_act = *n_active;
    const int n_warps = gridDim.x * T3::WARPS;
    TileWalk tw;
    for (walk_begin(tw, ticket, lane, blockIdx.x * T3::WARPS + warp, n_warps, n_act, fixed8); tw.a < n_act; walk_next(tw, ticket, lane, n_warps)) {
        walk_prefetch(tw, ticket, lane);
        const int a = tw.a;
        TileCtx tc;
        tile_from_list(g, __ldg(&tiles[a]), tc, tab, a);
        if (tc.count == 0) {          // a pseudo tile (ignored / dropped / migrated particles)
            continue;
        }
        for (int k = lane; k < T3::QSLOTS; k += 32) tile[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        // clear_grid for the node records of this tile's own block (nothing touches `grid` in this
        // kernel; "p2g 2" deposits into it next)
        zero_own_block(g, tc, lane, grid);
        // software pipeline: record one window ahead, index two windows ahead
        int first = 0;
        WinLane wl0 = window_lane(tc, 0, lane, first);
        WinLane wl1 = window_lane(tc, 1, lane, first);
        float4 p_next = make_float4(0.f, 0.f, 0.f, 0.f);
        if (wl0.active) p_next = fetch_f4(tq.P, P, fetch_i(tq.src, src, tc.base + wl0.slot));
        int i_next = wl1.active ? fetch_i(tq.src, src, tc.base + wl1.slot) : 0;
        for (int w = 0; w < tc.windows; ++w) {
            const bool active = wl0.active;
            const float4 p = p_next;
            if (wl1.active) p_next = fetch_f4(tq.P, P, i_next);
            wl0 = wl1;
            wl1 = window_lane(tc, w + 2, lane, first);
            if (wl1.active) i_next = fetch_i(tq.src, src, tc.base + wl1.slot);
            // Idle lanes skip the read-modify-writes altogether: a 128-bit access is served per quarter warp, and a
            // quarter none of whose lanes takes part costs no wavefront (ncu: an LDS.128 executed by mirrored idle
            // lanes always costs 4).  The lanes that do take part meet at a warp barrier of their own.
            const unsigned amask = __ballot_sync(0xffffffffu, active);
            if (active) {
                TStencil s;
                tile_stencil(g, tc, p.x, p.y, p.z, s);
                const float wqm[4] = {s.wq[0] * p.w, s.wq[1] * p.w, s.wq[2] * p.w, s.wq[3] * p.w};
                float4* n0 = tile + s.node0q;
#pragma unroll
                for (int oy = 0; oy < 3; ++oy)
#pragma unroll
                    for (int ox = 0; ox < 3; ++ox) {
                        const float wxy = s.wx[ox] * s.wy[oy];
                        float4* nd = n0 + ox + T3::NX * oy;
                        // the quad of this column: private to this lane within the window
                        float4 q4 = *nd;
                        q4.x += wxy * wqm[0];
                        q4.y += wxy * wqm[1];
                        q4.z += wxy * wqm[2];
                        q4.w += wxy * wqm[3];
                        *nd = q4;
                        __syncwarp(amask);
                    }
            }
            __syncwarp();
        }
        // flush: lane -> footprint column c = lane + 32*it (100 columns), six nodes along z
        float* peer_lo = (PEER && tc.c0[2] == g.slab_lo) ? ph.gmass[0] : nullptr;              // warp-uniform
        float* peer_hi = (PEER && tc.c0[2] + T3::Z == g.slab_hi) ? ph.gmass[1] : nullptr;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int c = lane + 32 * it;
            if (c < T3::NX * T3::NY) {
                const int ly = c / T3::NX, lx = c - ly * T3::NX;
                const int x = tc.c0[0] - 1 + lx, y = tc.c0[1] - 1 + ly;
                const float4 q0 = tile[c], q1 = tile[c + T3::PLANE];
                const float m6[6] = {q0.x, q0.y, q0.z + q1.x, q0.w + q1.y, q1.z, q1.w};
                if (!tc.edge || (x >= 0 && y >= 0 && x < g.size[0] && y < g.size[1])) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const int z = tc.c0[2] - 1 + k;
                        if (m6[k] != 0.0f && (!tc.edge || (z >= 0 && z < g.size[2]))) {
                            const int gi = node_addr(g, x, y, z);
                            atomicAdd(&gmass[gi], m6[k]);
                            // the two node planes a slab face shares: the neighbour's copy as well (NVLink)
                            if (PEER && k < 2 && peer_lo) red_add_sys(&peer_lo[gi], m6[k]);
                            if (PEER && k >= 4 && peer_hi) red_add_sys(&peer_hi[gi], m6[k]);
                        }
                    }
                }
            }
        }
        // no fence here: the kernel boundary and the system fences of k_peer_barrier order these deposits
        // before the flag the neighbour waits for
        __syncwarp();
    }
}

// ---- p2g 2: density, pressure, stress; fused momentum + force scatter ---------------------------

struct P2GSmem {
    float4 acc[T3::WARPS][T3::SLOTS];   // {momentum + force, mass} accumulators
    float4 mass[T3::WARPS][T3::QSLOTS]; // complete node masses (from k_mass_tiled) as z quads
};

struct PRec {   // one particle's streams
    float4 p, v, ca, cb;
    float cc;
};

__device__ __forceinline__ void load_prec(const Particles& q, const ParticleTex& tq, int i, bool ok, PRec& r) {
    if (ok) {
        r.p = fetch_f4(tq.P, q.P, i);
        r.v = fetch_f4(tq.V, q.V, i);
        r.ca = fetch_f4(tq.CA, q.CA, i);
        r.cb = fetch_f4(tq.CB, q.CB, i);
        r.cc = fetch_f(tq.CC, q.CC, i);
    }
}

template <bool PEER, bool TMA>   // TMA: the tile is flushed by tensor-memory-accelerator reductions (tm_grid)
__global__ void __launch_bounds__(T3::THREADS, 4)
k_p2g_tiled(const __grid_constant__ Geo g, Particles q, const int* __restrict__ src,
            const int4* __restrict__ tiles, const int* __restrict__ n_active,
            const float* __restrict__ gmass, float4* __restrict__ grid,
            float* __restrict__ dbg_density, float* __restrict__ dbg_pressure, PeerHalo ph,
            const __grid_constant__ CUtensorMap tm_grid, const __grid_constant__ CUtensorMap tm_mass, int tma_mass,
            ParticleTex tq, const unsigned char* __restrict__ tab, int* __restrict__ ticket, int fixed8) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    P2GSmem& sm = *reinterpret_cast<P2GSmem*>(smem_raw);
    __shared__ __align__(8) unsigned long long bars[T3::WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* acc = sm.acc[warp];
    float4* ms = sm.mass[warp];
    int load_parity = 0;
    if (TMA) {
        if (lane == 0) {
            mbar_init(&bars[warp], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    const int n_act = *n_active;
    const int n_warps = gridDim.x * T3::WARPS;
    TileWalk tw;
    for (walk_begin(tw, ticket, lane, blockIdx.x * T3::WARPS + warp, n_warps, n_act, fixed8); tw.a < n_act; walk_next(tw, ticket, lane, n_warps)) {
        walk_prefetch(tw, ticket, lane);
        const int a = tw.a;
        TileCtx tc;
        tile_from_list(g, __ldg(&tiles[a]), tc, tab, a);
        if (tc.count == 0) {          // a pseudo tile (ignored / dropped / migrated particles)
            continue;
        }
        int first = 0;
        WinLane wl0 = window_lane(tc, 0, lane, first);
        WinLane wl1 = window_lane(tc, 1, lane, first);
        PRec nxt;
        nxt.p = nxt.v = nxt.ca = nxt.cb = make_float4(0.f, 0.f, 0.f, 0.f);
        nxt.cc = 0.0f;
        load_prec(q, tq, wl0.active ? fetch_i(tq.src, src, tc.base + wl0.slot) : 0, wl0.active, nxt);
        int i_next = wl1.active ? fetch_i(tq.src, src, tc.base + wl1.slot) : 0;
        const FootLane fl = foot_lane(lane);
        // the node masses of the footprint as ONE tensor copy (box 12 x 10 x 6 floats) into the accumulator
        // tile, which is idle until the first window; unpacked into z quads below
        // (y / z may start at -1: loads zero-fill; x starts at c0[0] >= 0; the tensor ends one node early in x)
        const bool tma_load = TMA && tma_mass && tc.c0[0] + MBOX_X <= g.size[0];
        if (tma_load) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_expect_tx(&bars[warp], MBOX_FLOATS * 4);
                // the tensor's x coordinate is the node's x + 1 (base shifted for alignment): node c0 - 1 -> c0
                tma_load_box3(acc, &tm_mass, tc.c0[0], tc.c0[1] - 1, tc.c0[2] - 1, &bars[warp]);
            }
            mbar_wait(&bars[warp], load_parity);
            load_parity ^= 1;
            const float* st = reinterpret_cast<const float*>(acc);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int c = lane + 32 * it;
                if (c < T3::NX * T3::NY) {
                    const int ly = c / T3::NX, lx = c - ly * T3::NX;
                    const float* col = st + lx + MBOX_X * ly;
                    const float m0 = col[0], m1 = col[MBOX_X * T3::NY], m2 = col[2 * MBOX_X * T3::NY],
                                m3 = col[3 * MBOX_X * T3::NY], m4 = col[4 * MBOX_X * T3::NY], m5 = col[5 * MBOX_X * T3::NY];
                    ms[c] = make_float4(m0, m1, m2, m3);
                    ms[c + T3::PLANE] = make_float4(m2, m3, m4, m5);
                }
            }
            __syncwarp();
            for (int k = lane; k < T3::SLOTS; k += 32) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else
        {   // node masses of the footprint, column by column (six nodes along z -> two quads): all loads
            // in flight before the first store
            float mv[4][6];
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int c = lane + 32 * it;
                const int ly = c / T3::NX, lx = c - ly * T3::NX;
                const int x = tc.c0[0] - 1 + lx, y = tc.c0[1] - 1 + ly;
                const bool col_ok = c < T3::NX * T3::NY && (!tc.edge || (x >= 0 && y >= 0 && x < g.size[0] && y < g.size[1]));
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const int z = tc.c0[2] - 1 + k;
                    const bool ok = col_ok && (!tc.edge || (z >= 0 && z < g.size[2]));
                    const int gi = ok ? node_addr(g, x, y, z) : -1;
                    mv[it][k] = gi >= 0 ? __ldg(&gmass[gi]) : 0.0f;
                }
            }
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int c = lane + 32 * it;
                if (c < T3::NX * T3::NY) {
                    ms[c] = make_float4(mv[it][0], mv[it][1], mv[it][2], mv[it][3]);
                    ms[c + T3::PLANE] = make_float4(mv[it][2], mv[it][3], mv[it][4], mv[it][5]);
                }
            }
            for (int k = lane; k < T3::SLOTS; k += 32) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();

        for (int w = 0; w < tc.windows; ++w) {
            const bool active = wl0.active;
            const int d = tc.base + wl0.slot;   // sorted slot
            const PRec cur = nxt;
            load_prec(q, tq, i_next, wl1.active, nxt);        // prefetch the next window
            wl0 = wl1;
            wl1 = window_lane(tc, w + 2, lane, first);
            if (wl1.active) i_next = fetch_i(tq.src, src, tc.base + wl1.slot);
            // Idle lanes skip the shared-memory work altogether: a 128-bit access is served per quarter warp, and a
            // quarter none of whose lanes takes part costs no wavefront (ncu: an LDS.128 executed by mirrored idle
            // lanes always costs 4).  The lanes that do take part meet at a warp barrier of their own.
            const unsigned amask = __ballot_sync(0xffffffffu, active);
            if (active) {
                TStencil s;
                tile_stencil(g, tc, cur.p.x, cur.p.y, cur.p.z, s);
                // density = sum_i m_i w_ip (3d:198-215): one quad of node masses per stencil column,
                // summed z -> x -> y
                float density = 0.0f;
#pragma unroll
                for (int oy = 0; oy < 3; ++oy) {
                    const float4* row = ms + s.node0q + T3::NX * oy;
                    float rs = 0.0f;
#pragma unroll
                    for (int ox = 0; ox < 3; ++ox) {
                        const float4 q4 = row[ox];
                        rs += (q4.x * s.wq[0] + q4.y * s.wq[1] + q4.z * s.wq[2] + q4.w * s.wq[3]) * s.wx[ox];
                    }
                    density += rs * s.wy[oy];
                }
                const float m = cur.p.w;
                const float volume = m * __frcp_rn(density);
                const float pressure = tait_pressure_fast(g, density);
                if (dbg_density) dbg_density[d] = density;
                if (dbg_pressure) dbg_pressure[d] = pressure;
                // M = m C + T,  T = -4 V (-p I + mu (C + C^T)) dt   (3d:222-225), column-major
                const float C[9] = {cur.ca.x, cur.ca.y, cur.ca.z, cur.ca.w, cur.cb.x, cur.cb.y, cur.cb.z, cur.cb.w, cur.cc};
                const float s1 = -4.0f * volume * g.dt;
                float M[9];
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        float stress = g.mu * (C[3 * c + r] + C[3 * r + c]);
                        if (c == r) stress -= pressure;
                        M[3 * c + r] = m * C[3 * c + r] + s1 * stress;
                    }
                // value at offset o: w * (b + ox*M0 + oy*M1 + oz*M2), b = m v + M * (-1 - c)
                const float dx0 = -1.0f - s.cx, dy0 = -1.0f - s.cy, dz0 = -1.0f - s.cz;
                float b[3];
                b[0] = m * cur.v.x + M[0] * dx0 + M[3] * dy0 + M[6] * dz0;
                b[1] = m * cur.v.y + M[1] * dx0 + M[4] * dy0 + M[7] * dz0;
                b[2] = m * cur.v.z + M[2] * dx0 + M[5] * dy0 + M[8] * dz0;
                // per-z factors: node k gets wz_k * A + (k * wz_k) * G with A = wxy*c, G = wxy*M2
                const float q1 = s.wz[1], q2 = 2.0f * s.wz[2];
                float4* n0 = acc + s.node0;
#pragma unroll
                for (int oy = 0; oy < 3; ++oy) {
                    const float r0 = b[0] + oy * M[3], r1 = b[1] + oy * M[4], r2 = b[2] + oy * M[5];
#pragma unroll
                    for (int ox = 0; ox < 3; ++ox) {
                        const float wxy = s.wx[ox] * s.wy[oy];
                        const float A0 = wxy * (r0 + ox * M[0]), A1 = wxy * (r1 + ox * M[1]), A2 = wxy * (r2 + ox * M[2]);
                        const float G0 = wxy * M[6], G1 = wxy * M[7], G2 = wxy * M[8];
                        const float mw = wxy * m;
                        float4* nd = n0 + ox + T3::NX * oy;
                        // three nodes along z: private to this lane within the window
                        float4 a0 = nd[0], a1 = nd[T3::PLANE], a2 = nd[2 * T3::PLANE];
                        a0.x += s.wz[0] * A0;  a0.y += s.wz[0] * A1;  a0.z += s.wz[0] * A2;  a0.w += s.wz[0] * mw;
                        a1.x += s.wz[1] * A0 + q1 * G0;  a1.y += s.wz[1] * A1 + q1 * G1;
                        a1.z += s.wz[1] * A2 + q1 * G2;  a1.w += s.wz[1] * mw;
                        a2.x += s.wz[2] * A0 + q2 * G0;  a2.y += s.wz[2] * A1 + q2 * G1;
                        a2.z += s.wz[2] * A2 + q2 * G2;  a2.w += s.wz[2] * mw;
                        nd[0] = a0;
                        nd[T3::PLANE] = a1;
                        nd[2 * T3::PLANE] = a2;
                        __syncwarp(amask);
                    }
                }
            }
            __syncwarp();
        }
        if (TMA && !tc.low_rim) {   // (a reduction box must not start at a negative coordinate; overhang at the far end is fine)
            // six 10x10 planes as tensor reductions (cp.reduce.async.bulk.tensor .add, SASS UTMAREDG): no LDS, no
            // per-node REDG, no index arithmetic.  Every lane orders its own accumulator stores (generic proxy)
            // before the async proxy, then the warp meets, then one lane issues.
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int lz = 0; lz < T3::NZ; ++lz)
                    tma_reduce_add_plane(&tm_grid, tc.c0[0] - 1, tc.c0[1] - 1, tc.c0[2] - 1 + lz, acc + lz * T3::PLANE);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (!PEER) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // acc may be reused
            }
        } else {
#pragma unroll 4
            for (int it = 0; it < FOOT_STEPS; ++it) {
                int gi;
                const int sl = foot_step(g, tc, fl, it, gi);
                if (gi >= 0) {
                    const float4 v = acc[sl];
                    if (v.w != 0.0f || v.x != 0.0f || v.y != 0.0f || v.z != 0.0f) atomicAdd(&grid[gi], v);
                }
            }
        }
        if (PEER) {
            // the two node planes a slab face shares: the neighbour's copy as well (red.add over NVLink)
            float4* peer_lo = tc.c0[2] == g.slab_lo ? ph.grid[0] : nullptr;              // warp-uniform
            float4* peer_hi = tc.c0[2] + T3::Z == g.slab_hi ? ph.grid[1] : nullptr;
            if (peer_lo || peer_hi) {
                for (int it = 0; it < FOOT_STEPS; ++it) {
                    int gi;
                    const int sl = foot_step(g, tc, fl, it, gi);
                    if (gi < 0) continue;
                    float4* dst = sl < 2 * T3::PLANE ? peer_lo : (sl >= 4 * T3::PLANE ? peer_hi : nullptr);
                    if (!dst) continue;
                    const float4 v = acc[sl];
                    if (v.w != 0.0f || v.x != 0.0f || v.y != 0.0f || v.z != 0.0f) red_add_sys(&dst[gi], v);
                }
            }
            if (TMA && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
    }
    if (TMA && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // reductions performed before the warp exits
}

// ---- update + g2p -----------------------------------------------------------------------------

// Slab runs: the particles that leave this rank's slab are packed here for the neighbour rank.
struct SlabBufs {
    float* rec[2];   // 17 words per record (16 f32 + id), through the lower / upper face
    int cap;
};

// COUNT: also start the next substep's neighbour search (sort.cuh): every particle of the tile
// gets its new bucket; the ones that stay in this tile are ranked with shared-memory integer
// atomics (native ATOMS.ADD), the few that change tile or are dropped go to the immigrant list.
// g2p has no write conflicts, so it walks the tile's slots 32 at a time, every lane busy, whatever the windows are.
// (Measured, tools/experiments: walking window by window instead — no bank conflicts under ORDER_CLASS_Q, but a fifth
// more iterations — costs 12 %; five CTAs per SM at 96 registers cost 10 %.)
template <bool COUNT, bool TMA>   // TMA: the footprint is loaded by the tensor memory accelerator (tm_grid)
__global__ void __launch_bounds__(T3::THREADS, 4)
k_g2p_tiled(const __grid_constant__ Geo g, Particles q, Particles qn, const int* __restrict__ src,
            const int4* __restrict__ tiles, const int* __restrict__ n_active,
            const float4* __restrict__ grid, const float* __restrict__ mouse, SortTables st, SlabBufs sb,
            float* __restrict__ gmass, int* __restrict__ gz, const int* __restrict__ epoch_dev,
            const __grid_constant__ CUtensorMap tm_grid, ParticleTex tq, int* __restrict__ ticket, int fixed8) {
    const int epoch = *epoch_dev + 1;   // this substep's number (k_tail advances the counter after this kernel)
    if (mouse && mouse[2] == 0.0f) mouse = nullptr;   // {x, y, present}: the pointer itself never changes (CUDA graphs)
    __shared__ __align__(128) float4 sm[T3::WARPS * T3::SLOTS];
    __shared__ int scnt_all[T3::WARPS * TILE_CELLS];
    __shared__ __align__(8) unsigned long long bars[T3::WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* vt = sm + warp * T3::SLOTS;
    int* scnt = scnt_all + warp * TILE_CELLS;
    const int n_act = *n_active;
    const int n_warps = gridDim.x * T3::WARPS;
    int tma_parity = 0;
    if (TMA) {
        if (lane == 0) {
            mbar_init(&bars[warp], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    TileWalk tw;
    for (walk_begin(tw, ticket, lane, blockIdx.x * T3::WARPS + warp, n_warps, n_act, fixed8); tw.a < n_act; walk_next(tw, ticket, lane, n_warps)) {
        walk_prefetch(tw, ticket, lane);
        const int a = tw.a;
        TileCtx tc;
        tile_from_list(g, __ldg(&tiles[a]), tc);
        if (tc.count == 0) {          // a pseudo tile (ignored / dropped / migrated particles)
            if (COUNT && lane == 0) st.imm_cnt[a] = 0;
            continue;
        }
        // q = state before this substep (storage order, read through src);
        // qn = state after it, written at the sorted slot.  g2p has no write conflicts, so it walks
        // the tile's slots 32 at a time regardless of the window structure.
        const bool all_active = tile_all_active(g, tc);
        int i_cur = lane < tc.count ? fetch_i(tq.src, src, tc.base + lane) : 0;
        float4 p_next = make_float4(0.f, 0.f, 0.f, 0.f);
        float id_next = 0.0f;            // the particle id travels in V.w: fetched with the position, a window ahead
        if (lane < tc.count) {
            p_next = fetch_f4(tq.P, q.P, i_cur);
            id_next = tq.V ? tex1Dfetch<float4>(tq.V, i_cur).w : __ldg(&q.V[i_cur].w);
        }
        int i_next = 32 + lane < tc.count ? fetch_i(tq.src, src, tc.base + 32 + lane) : 0;
        int n_leave = 0;
        if (COUNT) {
#pragma unroll
            for (int j = 0; j < TILE_CELLS / 32; ++j) scnt[lane + 32 * j] = 0;
        }
        // clear_grid for the node masses of this tile's own block, one substep ahead: nothing reads
        // `gmass` between "p2g 2" and the next "p2g 1"; the stamp tells k_clear_tiles to skip the block
        zero_own_block(g, tc, lane, gmass);
        if (lane == 0) gz[tc.tile] = epoch;
        // node records of the footprint: cp.async (LDGSTS) straight into shared memory, all 19 per
        // lane in flight, zero-filled outside the p_rect grid; then update_grid in place
        const FootLane fl = foot_lane(lane);
        if (TMA) {   // every tile, the rim included: loads zero-fill whatever lies outside the tensor
            // six 10x10 planes by the tensor memory accelerator (SASS UTMALDG): no LSU wavefronts; every
            // lane orders its previous reads and writes of the tile before the async proxy, then the warp meets
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_expect_tx(&bars[warp], T3::NZ * FOOT_PLANE_BYTES);
#pragma unroll
                for (int lz = 0; lz < T3::NZ; ++lz)
                    tma_load_plane(vt + lz * T3::PLANE, &tm_grid, tc.c0[0] - 1, tc.c0[1] - 1, tc.c0[2] - 1 + lz, &bars[warp]);
            }
            __syncwarp();
            mbar_wait(&bars[warp], tma_parity);
            tma_parity ^= 1;
        } else {
#pragma unroll
            for (int it = 0; it < FOOT_STEPS; ++it) {
                int gi;
                const int sl = foot_step(g, tc, fl, it, gi);
                if (fl.rsub < 3) {
                    const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(vt + sl));
                    const float4* gp = grid + (gi >= 0 ? gi : 0);
                    const int bytes = gi >= 0 ? 16 : 0;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gp), "r"(bytes) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        // update_grid in place (3d:253-256): a linear sweep over the tile's 624 slots — no footprint index arithmetic;
        // the four pad slots behind each plane are never gathered, whatever they hold
#pragma unroll 4
        for (int k = lane; k < T3::SLOTS; k += 32) {
            float4 nd = vt[k];
            if (nd.w > 0.0f) {   // one reciprocal (MUFU.RCP, 1 ulp), three multiplies
                const float inv = rcp_approx(nd.w);
                nd.x = nd.x * inv + g.dtg[0];
                nd.y = nd.y * inv + g.dtg[1];
                nd.z = nd.z * inv + g.dtg[2];
                vt[k] = nd;
            }
        }
        __syncwarp();

        for (int it = 0; it < tc.count; it += 32) {
            const bool active = it + lane < tc.count;
            const int d = tc.base + it + lane;   // sorted slot = index in the new buffer
            const int i = i_cur;                 // storage index in the old buffer
            const float4 p = p_next;
            const float idw = id_next;
            i_cur = i_next;
            if (it + 32 + lane < tc.count) {   // prefetch the next window
                p_next = fetch_f4(tq.P, q.P, i_next);
                id_next = tq.V ? tex1Dfetch<float4>(tq.V, i_next).w : __ldg(&q.V[i_next].w);
            }
            if (it + 64 + lane < tc.count) i_next = fetch_i(tq.src, src, d + 64);
            float pos[3] = {p.x, p.y, p.z};
            // g2p walks a_rect blocks only (3d:263); warp-uniform short cut for tiles that lie inside a_rect
            const bool advance = active && (all_active || classify_pos<3>(g, pos) == CLS_ACTIVE);
            if (advance) {
                TStencil s;
                tile_stencil(g, tc, p.x, p.y, p.z, s);
                // S = sum w v ; Dk = sum w v (o_k - 1) ; B col k = Dk - S c_k
                float S[3] = {0.f, 0.f, 0.f}, Dx[3] = {0.f, 0.f, 0.f}, Dy[3] = {0.f, 0.f, 0.f}, Dz[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int oz = 0; oz < 3; ++oz) {
                    float Pz[3] = {0.f, 0.f, 0.f}, Pdx[3] = {0.f, 0.f, 0.f}, Pdy[3] = {0.f, 0.f, 0.f};
#pragma unroll
                    for (int oy = 0; oy < 3; ++oy) {
                        const float4* row = vt + s.node0 + T3::NX * oy + T3::PLANE * oz;
                        const float4 n0 = row[0], n1 = row[1], n2 = row[2];
                        float a0[3] = {n0.x * s.wx[0], n0.y * s.wx[0], n0.z * s.wx[0]};
                        float a2[3] = {n2.x * s.wx[2], n2.y * s.wx[2], n2.z * s.wx[2]};
                        float rs[3] = {a0[0] + n1.x * s.wx[1] + a2[0], a0[1] + n1.y * s.wx[1] + a2[1],
                                       a0[2] + n1.z * s.wx[1] + a2[2]};
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            const float rsw = rs[r] * s.wy[oy];
                            Pz[r] += rsw;
                            Pdx[r] += (a2[r] - a0[r]) * s.wy[oy];
                            if (oy == 0) Pdy[r] -= rsw;
                            if (oy == 2) Pdy[r] += rsw;
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const float pz = Pz[r] * s.wz[oz];
                        S[r] += pz;
                        Dx[r] += Pdx[r] * s.wz[oz];
                        Dy[r] += Pdy[r] * s.wz[oz];
                        if (oz == 0) Dz[r] -= pz;
                        if (oz == 2) Dz[r] += pz;
                    }
                }
                float vel[3] = {S[0], S[1], S[2]};
                float B[9];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    B[r] = Dx[r] - S[r] * s.cx;
                    B[3 + r] = Dy[r] - S[r] * s.cy;
                    B[6 + r] = Dz[r] - S[r] * s.cz;
                }
                integrate_particle<3>(g, pos, vel, mouse);
                // a particle that is still in a cell of this (a_rect) tile cannot have left p_rect (3d:356-366)
                if (!(all_active && in_tile(g, tc, pos)) && left_p_rect<3>(g, pos)) pos[0] = __int_as_float(0x7f800000);   // dropped: tombstone
                qn.P[d] = make_float4(pos[0], pos[1], pos[2], p.w);
                qn.V[d] = make_float4(vel[0], vel[1], vel[2], idw);
                qn.CA[d] = make_float4(4.0f * B[0], 4.0f * B[1], 4.0f * B[2], 4.0f * B[3]);
                qn.CB[d] = make_float4(4.0f * B[4], 4.0f * B[5], 4.0f * B[6], 4.0f * B[7]);
                qn.CC[d] = 4.0f * B[8];
            } else if (active) {   // frozen halo particle: carried over unchanged (3d:149 vs 3d:263)
                qn.P[d] = p;
                qn.V[d] = __ldg(&q.V[i]);
                qn.CA[d] = __ldg(&q.CA[i]);
                qn.CB[d] = __ldg(&q.CB[i]);
                qn.CC[d] = __ldg(&q.CC[i]);
            }
            if (COUNT) {
                // bucket of the (possibly moved) particle; frozen halo particles keep theirs
                int cls = -1, bucket = 0;
                if (active) {
                    // most particles stay inside their tile: inside an a_rect tile the new bucket is just the cell's
                    // place in the tile (no block keys, no rect tests); everything else takes the general rule
                    const int lx = rust_as_i32(floorf(pos[0])) - g.org[0] - tc.c0[0];
                    const int ly = rust_as_i32(floorf(pos[1])) - g.org[1] - tc.c0[1];
                    const int lz = rust_as_i32(floorf(pos[2])) - g.org[2] - tc.c0[2];
                    if (all_active && static_cast<unsigned>(lx) < static_cast<unsigned>(T3::X) &&
                        static_cast<unsigned>(ly) < static_cast<unsigned>(T3::Y) && static_cast<unsigned>(lz) < static_cast<unsigned>(T3::Z)) {
                        bucket = (tc.tile << 8) + local_cell_3d(lx, ly, lz);
                        cls = CLS_ACTIVE;
                    } else {
                        bucket = bucket_of<3>(g, make_float4(pos[0], pos[1], pos[2], 0.f), cls);
                    }
                }
                const bool stays = active && (bucket >> 8) == tc.tile;
                const bool leaves = active && !stays;
                if (active) st.gcell[d] = bucket;
                if (stays) st.rank[d] = atomicAdd(&scnt[bucket & (TILE_CELLS - 1)], 1);
                if (g.slab_on) {
                    // left this rank's slab (bucket_of said "migrated"): hand the record to the neighbour
                    int side = -1;
                    if (advance && bucket == migrated_bucket(g))
                        side = rust_as_i32(floorf(pos[2])) - g.org[2] < g.slab_lo ? 0 : 1;
#pragma unroll
                    for (int sd = 0; sd < 2; ++sd) {
                        const unsigned mm = __ballot_sync(0xffffffffu, side == sd);
                        if (mm == 0) continue;
                        int slot = 0;
                        if (lane == 0) slot = atomicAdd(&st.scal[SCAL_MIG_LO + sd], __popc(mm));
                        slot = __shfl_sync(0xffffffffu, slot, 0) + __popc(mm & ((1u << lane) - 1u));
                        if (side == sd) {
                            if (slot < sb.cap) {
                                float* r = sb.rec[sd] + static_cast<size_t>(slot) * MIG_WORDS;
                                const float4 a = qn.P[d], v = qn.V[d], ca = qn.CA[d], cb = qn.CB[d];
                                r[0] = a.x; r[1] = a.y; r[2] = a.z;
                                r[3] = v.x; r[4] = v.y; r[5] = v.z;
                                r[6] = ca.x; r[7] = ca.y; r[8] = ca.z; r[9] = ca.w;
                                r[10] = cb.x; r[11] = cb.y; r[12] = cb.z; r[13] = cb.w;
                                r[14] = qn.CC[d];
                                r[15] = a.w;
                                r[16] = v.w;   // id bits
                            } else {
                                st.scal[SCAL_MIG_OVERFLOW] = 1;
                            }
                        }
                    }
                }
                // leavers are listed inside this tile's own slot range: no global counter
                const unsigned lm = __ballot_sync(0xffffffffu, leaves);
                if (leaves) st.imm_list[tc.base + n_leave + __popc(lm & ((1u << lane) - 1u))] = d;
                n_leave += __popc(lm);
            }
        }
        __syncwarp();
        if (COUNT) {
            // residents per cell -> count[], their total -> tile_total[] (plain stores: this warp owns
            // the tile; immigrants are added by k_immigrants afterwards)
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
        }
        __syncwarp();
    }
}

}  // namespace fluid
