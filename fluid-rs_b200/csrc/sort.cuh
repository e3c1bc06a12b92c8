// sort.cuh — uniform-grid neighbour search: cell keys, counting sort, reorder, cellStart tables.
//
// GPU stand-in for the reference's `particles_mul: AHashMap<block key, Vec<Particle>>` and its
// per-block migration mailboxes `swap_mul` (3d:52,56,104-108,345-380).
//
// Buckets.  The grid is cut into tiles of 256 cells (8x8x4 in 3D, 16x16 in 2D); bucket id =
// tile * 256 + cell-in-tile.  Two pseudo tiles follow the real ones: tile n_tiles collects the
// particles the reference ignores (block key outside p_rect), tile n_tiles + 1 the ones it dropped
// in migration (3d:356-366).
//
// Every substep the particles are brought into the order (tile, rank in cell, cell):
//   counts   count[bucket], tile_total[tile] and per particle (bucket, rank)
//            - steady state (3D tiled path): g2p counts the particles that stay in their tile in
//              SHARED memory (native integer atomics) and lists the few that change tile;
//              k_immigrants / k_tail give those their rank with global atomics;
//            - cold start (after add_particles / set_rect, and the generic path): k_classify_all
//              does it for every particle with global atomics;
//   scan     exclusive scan over the TILE totals only (hand-written three-kernel scan, no CUB);
//   tables   one warp per tile: cell offsets inside the tile (cellStart), the window tables of the
//            chosen order, the active-tile list; resets count[] for reuse;
//   src      one thread per particle: its slot in closed form from those tables -> src[slot];
//   reorder  none: the tile kernels gather through src[] and g2p writes at the sorted slots.
#pragma once

#include "common.cuh"

namespace fluid {

struct Particles {
    float4* P;    // pos.xyz, mass
    float4* V;    // vel.xyz, id (int bits)
    float4* CA;   // C[0..3]   (column-major affine matrix, 3d:39)
    float4* CB;   // C[4..7]   (3D only)
    float*  CC;   // C[8]      (3D only)
};

// The same streams (and src[]) as linear TEXTURES.  The tile kernels gather particle records through src[]: 32
// lanes hit ~17 different 128-byte lines per 128-bit load, and every line is a wavefront of the LSU pipe — the
// pipe that binds those kernels (20 % of k_p2g_tiled's LSU wavefronts were these gathers).  A texture fetch takes
// the TEX pipe instead, which the kernels leave idle.  All zero = not available (the kernels use __ldg then).
struct ParticleTex {
    cudaTextureObject_t P, V, CA, CB, CC, src;
};
__device__ __forceinline__ float4 fetch_f4(cudaTextureObject_t t, const float4* __restrict__ p, int i) {
    return t ? tex1Dfetch<float4>(t, i) : __ldg(&p[i]);
}
__device__ __forceinline__ float fetch_f(cudaTextureObject_t t, const float* __restrict__ p, int i) {
    return t ? tex1Dfetch<float>(t, i) : __ldg(&p[i]);
}
__device__ __forceinline__ int fetch_i(cudaTextureObject_t t, const int* __restrict__ p, int i) {
    return t ? tex1Dfetch<int>(t, i) : __ldg(&p[i]);
}

// Everything the counting kernels write.
struct SortTables {
    int* gcell;        // per particle: bucket
    int* rank;         // per particle: rank inside its bucket
    int* count;        // per bucket (n_cells_pad + 512)
    int* tile_total;   // per tile (+2 pseudo tiles)
    int* imm_list;     // particles that changed tile in g2p: tile a's leavers sit at [first slot of a, +imm_cnt[a])
    int* imm_cnt;      // per active-tile-list entry: number of leavers
    int* scal;         // device scalars (SCAL_*)
};

// Peer-memory halo of a slab run (NVLink P2P, mapped through CUDA IPC): side 0 = the rank below, 1 = above.
// The tile kernels add every deposit that falls into the two node planes an interface shares into the
// neighbour's arrays as well, so both ranks hold complete sums with no plane exchange.
struct PeerHalo {
    float4* grid[2];            // neighbour's node records (same geometry, same indexing) or nullptr
    float* gmass[2];            // neighbour's node masses
    unsigned char* dirty[2][2]; // neighbour's two dirty-block flag arrays
    float* mig[2];              // neighbour's receive buffer for the particles this rank hands over (header record first)
    int* flag[2];               // neighbour's arrival flag for this rank (peer barrier)
};

constexpr int TILE_CELLS = 256;
constexpr int N_PSEUDO = 3;   // pseudo tiles behind the real ones: ignored, dropped, migrated away
constexpr int SCAL_N_ACTIVE = 0, SCAL_N_IMM = 1, SCAL_MIG_LO = 2, SCAL_MIG_HI = 3, SCAL_MIG_OVERFLOW = 4;
constexpr int SCAL_N_CAND = 5, SCAL_N_DIRTY = 6, SCAL_N_DIRTY2 = 7;
constexpr int SCAL_TICKET = 8;   // + 0, 1, 2: next list entry of k_mass_tiled / k_p2g_tiled / k_g2p_tiled (zeroed with the rest by the sort)
constexpr int SCAL_COUNT = 16;
constexpr int MIG_WORDS = 17;   // packed migrant record: 16 f32 + id

__device__ __forceinline__ bool is_tombstone(float x) { return isinf(x) && x > 0.0f; }

__device__ __forceinline__ int limbo_bucket(const Geo& g) { return g.n_cells_pad; }
__device__ __forceinline__ int dropped_bucket(const Geo& g) { return g.n_cells_pad + TILE_CELLS; }
__device__ __forceinline__ int migrated_bucket(const Geo& g) { return g.n_cells_pad + 2 * TILE_CELLS; }

// Bucket and class of a position (exact integer rules, common.cuh).
template <int DIM>
__device__ __forceinline__ int bucket_of(const Geo& g, const float4 p, int& cls) {
    if (is_tombstone(p.x)) {
        cls = CLS_DROPPED;
        return dropped_bucket(g);
    }
    const float pos[3] = {p.x, p.y, p.z};
    int cell[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < DIM; ++a) cell[a] = rust_as_i32(floorf(pos[a]));
    cls = classify_cells<DIM>(g, cell);
    if (cls == CLS_LIMBO) return limbo_bucket(g);
    int rel[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < DIM; ++a)   // key in p_rect => cell inside the grid; the clamp guards memory only
        rel[a] = min(max(cell[a] - g.org[a], 0), g.size[a] - 1);
    if (DIM == 3 && g.slab_on && (rel[2] < g.slab_lo || rel[2] >= g.slab_hi)) {
        cls = CLS_LIMBO;   // another rank's slab: not ours to deposit or advance
        return migrated_bucket(g);
    }
    return tiled_cell_index<DIM>(g, rel);
}

// One global-atomic count: rank inside the bucket, tile total aggregated per warp.
__device__ __forceinline__ void count_global(const SortTables& t, int i, int bucket, bool valid) {
    const int lane = threadIdx.x & 31;
    const int tile = valid ? (bucket >> 8) : (-1 - lane);
    if (valid) {
        t.gcell[i] = bucket;
        t.rank[i] = atomicAdd(&t.count[bucket], 1);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, tile);
    if (valid && lane == __ffs(peers) - 1) atomicAdd(&t.tile_total[tile], __popc(peers));
}

// Cold start: every particle through global atomics.
template <int DIM>
__global__ void __launch_bounds__(256)
k_classify_all(const __grid_constant__ Geo g, const float4* __restrict__ P, int n, SortTables t,
               int* __restrict__ class_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int cls = -1, bucket = 0;
    if (i < n) bucket = bucket_of<DIM>(g, P[i], cls);
    count_global(t, i, bucket, i < n);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        unsigned m = __ballot_sync(0xffffffffu, cls == c);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&class_count[c], __popc(m));
    }
}

// Steady state: the particles g2p listed because they left their tile (or were dropped).  g2p keeps
// one list per tile inside the tile's own slot range (no global counter, no same-address atomics);
// persistent warps walk the tile list of the sort that g2p ran on.
__global__ void __launch_bounds__(256)
k_immigrants(SortTables t, const int4* __restrict__ tiles, int n_tiles_listed_max) {
    const int n_list = min(t.scal[SCAL_N_ACTIVE], n_tiles_listed_max);
    const int lane = threadIdx.x & 31;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; a < n_list; a += n_warps) {
        const int n_imm = t.imm_cnt[a];
        if (n_imm == 0) continue;
        const int first = tiles[a].y;
        for (int j0 = 0; j0 < n_imm; j0 += 32) {
            const bool valid = j0 + lane < n_imm;
            int i = 0, bucket = 0;
            if (valid) {
                i = t.imm_list[first + j0 + lane];
                bucket = t.gcell[i];
            }
            count_global(t, i, bucket, valid);
        }
    }
}

// Steady state: the tail of the array (ignored and dropped particles) is not covered by any tile.
// It copies those records from the old buffer (through src) into their slots of the new one, as
// g2p does for the tiles, and counts them.
template <int DIM>
__global__ void __launch_bounds__(256)
k_tail(const __grid_constant__ Geo g, Particles from, Particles to, const int* __restrict__ src,
       const int* __restrict__ n_deposit, const int* __restrict__ n_end, int n, SortTables t, int* __restrict__ epoch_dev) {
    // runs right after k_g2p_tiled: the substep counter both compare their block stamps with moves on
    if (epoch_dev && blockIdx.x == 0 && threadIdx.x == 0) *epoch_dev += 1;
    const int first = *n_deposit;
    // slab runs carry only the ignored particles over: dropped and migrated ones end here
    if (n_end) n = min(n, *n_end);
    const int len = n - first;
    const int stride = gridDim.x * blockDim.x;
    const int rounds = (len + stride - 1) / stride;
    for (int r = 0; r < rounds; ++r) {
        const int d = first + r * stride + blockIdx.x * blockDim.x + threadIdx.x;
        const bool valid = d < n;
        int cls = -1, bucket = 0;
        if (valid) {
            const int i = src[d];
            const float4 p = from.P[i];
            to.P[d] = p;
            to.V[d] = from.V[i];
            to.CA[d] = from.CA[i];
            if (DIM == 3) {
                to.CB[d] = from.CB[i];
                to.CC[d] = from.CC[i];
            }
            bucket = bucket_of<DIM>(g, p, cls);
        }
        count_global(t, d, bucket, valid);
    }
}

// ---- exclusive scan (used over the tile totals) -----------------------------------------------

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;   // 2048 values per block

__device__ __forceinline__ int warp_inclusive_scan(int v) {
    int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Block-wide exclusive scan of one value per thread; returns the exclusive prefix and the
// block total through `total`.  blockDim.x <= 1024.
__device__ __forceinline__ int block_exclusive_scan(int v, int& total) {
    __shared__ int warp_sums[32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = warp_inclusive_scan(v);
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int nw = (blockDim.x + 31) >> 5;
        int s = lane < nw ? warp_sums[lane] : 0;
        s = warp_inclusive_scan(s);
        warp_sums[lane] = s;
    }
    __syncthreads();
    int base = wid ? warp_sums[wid - 1] : 0;
    total = warp_sums[((blockDim.x + 31) >> 5) - 1];
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_partial(const int* __restrict__ in, int m, int* __restrict__ block_sums) {
    int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
    int s = 0;
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < m) s += in[base + k];
    int total;
    block_exclusive_scan(s, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
k_scan_sums(int* __restrict__ block_sums, int nb) {
    int carry = 0;
    for (int base = 0; base < nb; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < nb ? block_sums[i] : 0;
        int total;
        int ex = block_exclusive_scan(v, total);
        if (i < nb) block_sums[i] = carry + ex;
        carry += total;
    }
}

// out[k] = exclusive prefix, out[m] = grand total; the input is zeroed for the next round.
// The indices of the non-zero inputs (tiles that hold particles) are appended to cand[] so that the
// per-tile kernel can run persistent warps over a dense list instead of one CTA per (mostly empty) tile.
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_final(int* __restrict__ in, int m, const int* __restrict__ block_sums, int* __restrict__ out,
             int* __restrict__ cand, int* __restrict__ n_cand) {
    int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < m) ? in[base + k] : 0;
        s += v[k];
    }
    int total;
    int ex = block_exclusive_scan(s, total) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < m) {
            out[base + k] = ex;
            in[base + k] = 0;
        }
        ex += v[k];
    }
    if (base <= m - 1 && m - 1 < base + SCAN_ITEMS) out[m] = ex;
    if (cand) {
        int nz = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) nz += v[k] > 0;
        const int lane = threadIdx.x & 31;
        const int inc = warp_inclusive_scan(nz);
        const int warp_total = __shfl_sync(0xffffffffu, inc, 31);
        int slot = 0;
        if (lane == 31 && warp_total) slot = atomicAdd(n_cand, warp_total);
        slot = __shfl_sync(0xffffffffu, slot, 31) + inc - nz;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k)
            if (v[k] > 0) cand[slot++] = base + k;
    }
}

// ---- per-tile order ---------------------------------------------------------------------------
//
// ORDER_CELL      (cell, rank): plain cell order (2D / generic path).
// ORDER_CLASS_RR  3D tiled path.  Cells are numbered class-major (local_cell_3d: bank class, column,
//                 z), so the cell-sorted sequence q = 0..N-1 of a tile lists class 0's columns first,
//                 then class 1's, ...  It is dealt round robin into
//                     W = max(ceil(N/32), max particles in one column)  windows:  window = q mod W.
//                 * Two particles of one (x,y) column are less than W apart in q, so a window never
//                   holds two particles of the same column: for a fixed stencil offset the 32 lanes of
//                   a window update 32 different node columns (and the three nodes along z of a lane
//                   are private to it) — the shared-memory read-modify-writes of phases_tiled.cuh
//                   need no atomics and no conflict passes.
//                 * Every window receives floor or ceil(N_b / W) particles of each bank class b; inside
//                   the window they are merged round robin over the classes, so the 8 lanes of a
//                   quarter warp hit 8 different 16-byte bank groups of the float4 node tile.
//                 The tile list carries W: {tile, first slot, N, W}; window w holds the slots
//                 [w*(N/W) + min(w, N%W), + N/W + (w < N%W)).
//                 (2D tiled path; 3D tiles too crowded for ORDER_CLASS_Q.)
// ORDER_CLASS_Q   3D tiled path.  A 128-bit shared-memory access is served per QUARTER warp, one wavefront per
//                 distinct address in the quarter's fullest bank group, so what a tile costs per access is the
//                 number of quarter warps it occupies plus their collisions, and no order can do with fewer
//                 than max(ceil(N/8), N_max) of them, N_max = the fullest bank class (tools/order_lab.py).  This
//                 order meets that bound with no collision at all: with q_b = a particle's place in its class's
//                 (column, z) sequence and
//                     W = max(fullest column, ceil(N_max / 4))  windows:
//                     window = q_b mod W,   round = q_b div W (0..3),
//                 round k of a window is quarter warp k, its members the classes with N_b > window + k*W in
//                 ascending order, one lane each from lane 8k on.  So a quarter never holds two particles of one
//                 class (8 different 16-byte bank groups, whatever the stencil offset), every class starts at
//                 window 0 (the windows in which the fullest class has its ceil(N_b/W) members are the same for
//                 all classes: sum over windows of the rounds in use = N_max), and two particles of a column are
//                 still less than W apart in q_b: no two in one window.  Lanes behind a round's members idle
//                 (an idle quarter costs no wavefront).  Slots stay compact: window after window, round after
//                 round.  The tile list carries W | ORDER_Q_FLAG; the class totals N_b (one byte each, N_b <= 4W
//                 <= 128), N_b div W and N_b mod W sit in the tile's table row.
enum TileOrder : int { ORDER_CELL = 0, ORDER_CLASS_RR = 1, ORDER_CLASS_Q = 2 };
constexpr int ORDER_Q_FLAG = 1 << 30;   // in tiles[a].w and tile_info[t].x beside W (<= 32; the other orders store any W > 0 or -W there)

// floor(x / w) for 0 <= x < 8192, 1 <= w <= 32 without an integer division: (x + 0.5) / w is at least
// 0.5/32 away from an integer while the float product is off by less than 2^-10.
__device__ __forceinline__ int small_div(int x, float inv_w) {
    return __float2int_rz((static_cast<float>(x) + 0.5f) * inv_w);
}

constexpr int PERM_WARPS = 4;
constexpr int PERM_MAX_W = 32;   // windows per tile covered by the in-window merge table
constexpr int TAB_BYTES = PERM_MAX_W * 8;   // per listed tile: members of class b in window w, one byte each

// What k_build_src needs to place a particle of a tile: {W (0 = plain cell order), tile-list entry}
// W < 0: |W| windows but no class merge (more than PERM_MAX_W windows or >= 8192 particles).

// One warp per tile that holds particles: cellStart of its 256 cells, the number of windows W, the
// class-in-window table, the active-tile list entry and the dirty marks of the node blocks the tile's
// particles can reach; count[] is reset for the next round.  The per-particle slots are computed by
// k_build_src from these tables (closed form), so no per-slot permutation is stored.
template <int ORDER>
__global__ void __launch_bounds__(PERM_WARPS * 32)
k_tile_tables(const __grid_constant__ Geo g, int* __restrict__ count, const int* __restrict__ tile_base,
              int* __restrict__ cell_off, int2* __restrict__ tile_info, unsigned char* __restrict__ tab,
              int4* __restrict__ tiles, int* __restrict__ scal, unsigned char* __restrict__ dirty,
              const int* __restrict__ cand, PeerHalo ph) {
    const int lane = threadIdx.x & 31;
    // persistent warps over the list of tiles that hold particles (k_scan_final)
    const int n_cand = scal[SCAL_N_CAND];
    if (blockIdx.x == 0 && threadIdx.x == 0) scal[SCAL_N_ACTIVE] = n_cand;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    // Software pipeline over this warp's tiles: the three dependent loads (list entry -> tile_base ->
    // counts) of the next tile are in flight while the current one is processed.
    const int a0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int t_nx = a0 < n_cand ? cand[a0] : 0;
    int t_nn = a0 + n_warps < n_cand ? cand[a0 + n_warps] : 0;
    int base_nx = 0, end_nx = 0;
    int4 ca_nx = make_int4(0, 0, 0, 0), cb_nx = ca_nx;
    if (a0 < n_cand) {
        base_nx = tile_base[t_nx];
        end_nx = tile_base[t_nx + 1];
        const int4* cp = reinterpret_cast<const int4*>(count + t_nx * TILE_CELLS + lane * 8);
        ca_nx = cp[0];
        cb_nx = cp[1];
    }
    for (int a = a0; a < n_cand; a += n_warps) {
        const int t = t_nx;
        const int base = base_nx;
        const int n_t = end_nx - base_nx;
        const int4 a4 = ca_nx, b4 = cb_nx;
        t_nx = t_nn;
        if (a + n_warps < n_cand) {
            base_nx = tile_base[t_nx];
            end_nx = tile_base[t_nx + 1];
            const int4* cp = reinterpret_cast<const int4*>(count + t_nx * TILE_CELLS + lane * 8);
            ca_nx = cp[0];
            cb_nx = cp[1];
        }
        if (a + 2 * n_warps < n_cand) t_nn = cand[a + 2 * n_warps];
        const int c_first = t * TILE_CELLS;
        if (t >= g.n_tiles) {   // pseudo tiles: a single bucket, slots stay in rank order
            if (lane == 0) {
                cell_off[c_first] = base;
                count[c_first] = 0;
                tiles[a] = make_int4(t, base, 0, 1);   // listed, but nothing for the tile kernels to do
                tile_info[t] = make_int2(0, a);
            }
            continue;
        }
        if (dirty && lane < 27) {
            // the particles of this tile deposit into the node blocks of its 3x3x3 tile neighbourhood:
            // mark them for k_clear_tiles (clear_grid, 3d:136-146, only where something was written)
            const int tx = t % g.tdim[0], r = t / g.tdim[0], ty = r % g.tdim[1], tz = r / g.tdim[1];
            const int nx = tx + lane % 3 - 1, ny = ty + (lane / 3) % 3 - 1, nz = tz + lane / 9 - 1;
            if (nx >= 0 && ny >= 0 && nz >= 0 && nx < g.tdim[0] && ny < g.tdim[1] && nz < g.tdim[2]) {
                const int blk = (nz * g.tdim[1] + ny) * g.tdim[0] + nx;
                dirty[blk] = 1;
                // block-sparse node storage: the first tile to reach a block takes one from the free list
                if (g.sp.blk && atomicCAS(&g.sp.blk[blk], -1, -2) == -1) {
                    const int top = atomicSub(&g.sp.scal[0], 1);
                    int b = 0;                      // pool exhausted: the overflow block, and an error for the host
                    if (top > 0) b = g.sp.free_list[top - 1];
                    else {
                        atomicAdd(&g.sp.scal[0], 1);
                        g.sp.scal[1] = 1;
                    }
                    g.sp.blk[blk] = b;              // consumed by later kernels only (kernel boundary orders it)
                }
                // a tile on a slab face also deposits into the neighbour's copy of the two shared node
                // planes (block layers tz-1, tz below / tz, tz+1 above): the neighbour has to clear them
                if (ph.dirty[0][0] && tz * Tile<3>::Z == g.slab_lo && nz <= tz) {
                    ph.dirty[0][0][blk] = 1;
                    ph.dirty[0][1][blk] = 1;
                }
                if (ph.dirty[1][0] && (tz + 1) * Tile<3>::Z == g.slab_hi && nz >= tz) {
                    ph.dirty[1][0][blk] = 1;
                    ph.dirty[1][1][blk] = 1;
                }
            }
        }
        // lane owns 8 consecutive cells (3D: two (x,y) columns of one bank class, 4 cells each)
        int cnt[8];
        {
            int4* cp = reinterpret_cast<int4*>(count + c_first + lane * 8);
            cnt[0] = a4.x; cnt[1] = a4.y; cnt[2] = a4.z; cnt[3] = a4.w;
            cnt[4] = b4.x; cnt[5] = b4.y; cnt[6] = b4.z; cnt[7] = b4.w;
            cp[0] = make_int4(0, 0, 0, 0);      // count[] is all zero again outside a sort
            cp[1] = make_int4(0, 0, 0, 0);
        }
        int mine = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) mine += cnt[j];
        const int q_first = warp_inclusive_scan(mine) - mine;   // cell-sorted offset of my first cell
        {
            int ex = base + q_first;
            int st[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                st[j] = ex;
                ex += cnt[j];
            }
            int4* op = reinterpret_cast<int4*>(cell_off + c_first + lane * 8);   // cellStart
            op[0] = make_int4(st[0], st[1], st[2], st[3]);
            op[1] = make_int4(st[4], st[5], st[6], st[7]);
        }
        if (ORDER == ORDER_CELL) {
            if (lane == 0) {
                tiles[a] = make_int4(t, base, n_t, (n_t + 31) / 32);
                tile_info[t] = make_int2(0, a);
            }
            continue;
        }
        const int col_max = __reduce_max_sync(0xffffffffu, max(cnt[0] + cnt[1] + cnt[2] + cnt[3], cnt[4] + cnt[5] + cnt[6] + cnt[7]));
        if (ORDER == ORDER_CLASS_Q) {
            int n_cls = mine;                       // class totals N_b (lanes 4b..4b+3 hold class b)
            n_cls += __shfl_xor_sync(0xffffffffu, n_cls, 1);
            n_cls += __shfl_xor_sync(0xffffffffu, n_cls, 2);
            const int n_max = __reduce_max_sync(0xffffffffu, n_cls);
            const int w_q = max(col_max, (n_max + 3) >> 2);
            const int nb = __shfl_sync(0xffffffffu, n_cls, 4 * (lane & 7));
            if (w_q <= PERM_MAX_W) {                // (then N_b <= 128: one byte each; more crowded tiles get ORDER_CLASS_RR)
                if (lane == 0) {
                    tiles[a] = make_int4(t, base, n_t, w_q | ORDER_Q_FLAG);
                    tile_info[t] = make_int2(w_q | ORDER_Q_FLAG, a);
                }
                if (lane < 8) {
                    unsigned char* row = tab + static_cast<size_t>(a) * TAB_BYTES;
                    const int d = small_div(nb, 1.0f / static_cast<float>(w_q));
                    row[lane] = static_cast<unsigned char>(nb);
                    row[8 + lane] = static_cast<unsigned char>(d);
                    row[16 + lane] = static_cast<unsigned char>(nb - d * w_q);
                }
                continue;
            }
        }
        const int w_count = max((n_t + 31) / 32, col_max);
        const bool merge = w_count <= PERM_MAX_W && n_t < 8192;
        if (lane == 0) {
            tiles[a] = make_int4(t, base, n_t, w_count);   // list slot = candidate index: no atomics
            tile_info[t] = make_int2(merge ? w_count : -w_count, a);
        }
        if (!merge) continue;
        // class totals N_b (lanes 4b..4b+3 hold class b) and class starts S_b
        int n_cls = mine;
        n_cls += __shfl_xor_sync(0xffffffffu, n_cls, 1);
        n_cls += __shfl_xor_sync(0xffffffffu, n_cls, 2);
        const int s_cls = __shfl_sync(0xffffffffu, q_first, lane & ~3);   // start of my class
        const float inv_w = 1.0f / static_cast<float>(w_count);
        // tab[w*8 + b] = members of class b in window w: those q in [S_b, S_b + N_b) with q = w (mod W)
        const int b = lane & 7;
        const int nb = __shfl_sync(0xffffffffu, n_cls, 4 * b);
        const int sb = __shfl_sync(0xffffffffu, s_cls, 4 * b);
        const int sb_mod = sb - small_div(sb, inv_w) * w_count;
        unsigned char* my_tab = tab + static_cast<size_t>(a) * TAB_BYTES;
        for (int w = lane >> 3; w < w_count; w += 4) {
            int off = w - sb_mod;
            if (off < 0) off += w_count;
            my_tab[w * 8 + b] = static_cast<unsigned char>(off < nb ? small_div(nb - off + w_count - 1, inv_w) : 0);
        }
    }
}

// Sorted slot -> storage index.  The particle streams are never reordered by a separate pass:
// the tile kernels gather through src[], and g2p writes the advanced particles straight into
// their sorted slots of the other buffer, so storage order is always last substep's sorted
// order and the gathers stay nearly coalesced.
//
// Slot of a particle with (bucket, rank): q = cellStart[bucket] - first slot of the tile + rank is its
// position in the tile's class-major cell order; window w = q mod W; inside the window the classes are
// merged round robin (ORDER_CLASS_RR above): with k = my index among my class's members of the window,
//     pos = sum_b min(n_b, k) + #{b < my class : n_b > k},   n_b = tab[w][b].
// Slot of one particle from its bucket, cellStart[bucket] + rank and its tile's {W, list entry} (formulas above).
__device__ __forceinline__ int build_src_slot(int bucket, int q_abs, int2 info, const int* __restrict__ cell_off,
                                              const int* __restrict__ tile_base, const unsigned char* __restrict__ tab) {
    if (info.x == 0) return q_abs;   // plain cell order (pseudo tiles, 2D particle-per-thread path)
    const int t = bucket >> 8;
    const int base = __ldg(&tile_base[t]);
    const int n_t = __ldg(&tile_base[t + 1]) - base;
    const int q = q_abs - base;
    if (info.x > 0 && (info.x & ORDER_Q_FLAG)) {   // ORDER_CLASS_Q
        const int w_count = info.x & 0xff;
        const float inv_w = __fdividef(1.0f, static_cast<float>(w_count));
        const int cls = (bucket & (TILE_CELLS - 1)) >> 5;
        const int qb = q_abs - __ldg(&cell_off[(t << 8) + (cls << 5)]);   // place in my class's (column, z) sequence: < N_b <= 128
        const int k = small_div(qb, inv_w);          // round
        const int w = qb - k * w_count;              // window
        const uint2* row = reinterpret_cast<const uint2*>(tab + static_cast<size_t>(info.y) * TAB_BYTES);
        const uint2 nb = __ldg(&row[0]), db = __ldg(&row[1]), mb = __ldg(&row[2]);
        // slots of the windows before w: sum_b (N_b div W) * w + min(N_b mod W, w)
        const unsigned w4 = static_cast<unsigned>(w) * 0x01010101u;
        int pos = w * static_cast<int>(__vsadu4(db.x, 0u) + __vsadu4(db.y, 0u)) +
                  static_cast<int>(__vsadu4(__vminu4(mb.x, w4), 0u) + __vsadu4(__vminu4(mb.y, w4), 0u));
        // rounds before mine in this window: #{b : N_b > w + k' W}, k' < k
#pragma unroll
        for (int kk = 0; kk < 3; ++kk)
            if (kk < k) {
                const unsigned t4 = static_cast<unsigned>(w + kk * w_count) * 0x01010101u;
                pos += (__popc(__vcmpgtu4(nb.x, t4)) + __popc(__vcmpgtu4(nb.y, t4))) >> 3;
            }
        // my place in the round: the classes below mine that reach it
        const unsigned q4 = static_cast<unsigned>(qb) * 0x01010101u;
        const unsigned long long gt = (static_cast<unsigned long long>(__vcmpgtu4(nb.y, q4)) << 32) | __vcmpgtu4(nb.x, q4);
        const unsigned long long below = cls == 0 ? 0ull : (~0ull >> (64 - 8 * cls));
        pos += __popcll(gt & below) >> 3;
        return base + pos;
    }
    if (info.x < 0) {    // too many windows for the merge table: round robin only
        const int w_count = -info.x;
        const int per = n_t / w_count, extra = n_t - per * w_count;
        const int pos = q / w_count, w = q - pos * w_count;
        return base + w * per + min(w, extra) + pos;
    }
    const int w_count = info.x;
    // (MUFU.RCP: 1 ulp off at most, far inside small_div's margin of 0.5/32 on a quotient below 8192)
    const float inv_w = __fdividef(1.0f, static_cast<float>(w_count));
    const int per = small_div(n_t, inv_w), extra = n_t - per * w_count;
    const int w = q - small_div(q, inv_w) * w_count;
    const int cls = (bucket & (TILE_CELLS - 1)) >> 5;
    const int s_cls = __ldg(&cell_off[(t << 8) + (cls << 5)]) - base;
    const int s_mod = s_cls - small_div(s_cls, inv_w) * w_count;
    int off = w - s_mod;
    if (off < 0) off += w_count;
    const int k = small_div(q - (s_cls + off), inv_w);   // my index among my class in window w
    const uint2 row = __ldg(reinterpret_cast<const uint2*>(tab + static_cast<size_t>(info.y) * TAB_BYTES + w * 8));
    const unsigned kk = static_cast<unsigned>(min(k, 255)) * 0x01010101u;
    // sum_b min(n_b, k): per-byte minimum, then the byte sum
    int pos = __vsadu4(__vminu4(row.x, kk), 0u) + __vsadu4(__vminu4(row.y, kk), 0u);
    // #{b < cls : n_b > k}
    const unsigned gx = __vcmpgtu4(row.x, kk), gy = __vcmpgtu4(row.y, kk);   // 0xff per byte where n_b > k
    const unsigned long long gt = (static_cast<unsigned long long>(gy) << 32) | gx;
    const unsigned long long below = cls == 0 ? 0ull : (~0ull >> (64 - 8 * cls));
    pos += __popcll(gt & below) >> 3;
    return base + w * per + min(w, extra) + pos;
}

// BUILD_SRC_PER_THREAD particles per thread: the kernel is a chain of dependent table look-ups (bucket -> cellStart ->
// tile entry -> class start -> window row), so independent chains in one thread are what hides their latency.
constexpr int BUILD_SRC_PER_THREAD = 4;
__global__ void __launch_bounds__(256, 8)   // 32 registers: the look-up chains want the occupancy (at 39 registers / 6 CTAs: +15 % time)
k_build_src(int n, const int* __restrict__ gcell, const int* __restrict__ rank,
            const int* __restrict__ cell_off, const int* __restrict__ tile_base,
            const int2* __restrict__ tile_info, const unsigned char* __restrict__ tab, int* __restrict__ src) {
    const int i0 = blockIdx.x * (256 * BUILD_SRC_PER_THREAD) + threadIdx.x;
    int bucket[BUILD_SRC_PER_THREAD], q_abs[BUILD_SRC_PER_THREAD];
    int2 info[BUILD_SRC_PER_THREAD];
#pragma unroll
    for (int k = 0; k < BUILD_SRC_PER_THREAD; ++k) {
        const int i = i0 + 256 * k;
        bucket[k] = i < n ? gcell[i] : 0;
        q_abs[k] = i < n ? rank[i] : 0;
    }
#pragma unroll
    for (int k = 0; k < BUILD_SRC_PER_THREAD; ++k) {
        q_abs[k] += __ldg(&cell_off[bucket[k]]);            // (bucket 0 for the lanes behind the end: a valid address)
        info[k] = __ldg(&tile_info[bucket[k] >> 8]);
    }
#pragma unroll
    for (int k = 0; k < BUILD_SRC_PER_THREAD; ++k) {
        const int i = i0 + 256 * k;
        if (i < n) src[build_src_slot(bucket[k], q_abs[k], info[k], cell_off, tile_base, tab)] = i;
    }
}

// Physical gather (only used to compact dropped particles away and for the steady-state tail).
template <int DIM>
__global__ void __launch_bounds__(256)
k_gather_range(Particles from, Particles to, const int* __restrict__ src, int first, int n) {
    int d = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    const int i = src[d];
    to.P[d] = from.P[i];
    to.V[d] = from.V[i];
    to.CA[d] = from.CA[i];
    if (DIM == 3) {
        to.CB[d] = from.CB[i];
        to.CC[d] = from.CC[i];
    }
}

}  // namespace fluid
