// common.cuh — geometry, integer cell/key rules and the quadratic stencil shared by all kernels.
//
// Reference: GossiperLoturot/fluid-rs src/3d_multi.rs ("3d:") and src/2d_multi.rs ("2d:").
// Integer outputs (cell = floor(pos) 3d:153, block key = div_euclid 3d:398-401) follow the
// reference bit for bit: IEEE division and fmodf, no fast-math.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace fluid {

// Tile of cells that one warp owns in the tiled kernels and that defines the sort order: 8 x 8 (x, y) columns,
// 4 cells deep in 3D, 1 in 2D.  A tile always owns 256 buckets (8 bank classes x 8 columns x 4 depth slots); in 2D
// only depth slot 0 of a column is used, so the 2D scenes (4 particles per cell, 2d:24) put about as many
// particles into a tile as the 3D ones (1 per cell) and the sort tables and window order are the same code.
template <int DIM> struct Tile;
template <> struct Tile<3> { static constexpr int X = 8, Y = 8, Z = 4, CELLS = 256; };
template <> struct Tile<2> { static constexpr int X = 8, Y = 8, Z = 1, CELLS = 256; };

// Block-sparse node storage (the reference keeps a hash map of blocks and a touched list so that cost follows
// the fluid, not the domain: 3d:52-55, 89-96, 136-146).  With `blk` set, the node arrays are a POOL of 8x8x4
// node blocks; blk[tile] is the pool block of a tile's node block (-1: none yet).  Blocks are taken from the
// free list when a tile's 3x3x3 neighbourhood first holds particles (k_tile_tables) and returned, zeroed, when
// it no longer does (k_clear_tiles).  scal[0] = free blocks left, scal[1] = pool exhausted (error).
struct SparsePool {
    int* blk;
    int* free_list;
    int* scal;
};

struct Geo {
    int dim;
    SparsePool sp;     // blk == nullptr: dense node arrays in the reference's layout
    int org[3];        // grid origin cell = p_rect.0 * grid_res            (3d:169)
    int size[3];       // grid size in cells = (p_rect.1 - p_rect.0) * res  (3d:94)
    int a_lo[3], a_hi[3], p_lo[3], p_hi[3];   // block-key rects (3d:80-86)
    int tdim[3];       // tiles per axis (size rounded up to the tile shape)
    int n_tiles;
    int n_cells_pad;   // n_tiles * 256: length of the tiled cell index space
    int guard;         // guard nodes before/after the grid array (zero-weight overreach)
    int slab_on;            // z-slab decomposition active
    int slab_lo, slab_hi;   // owned cells [lo, hi) along z, relative to the grid origin
    int res_i;         // grid_res
    int res_shift;     // log2(grid_res) when it is a power of two, else -1
    float res_f;       // grid_res as f32 (3d:399)
    float dt, rest_density, mu, stiffness, power, mouse_r2, pclamp;
    float inv_rest;    // 1 / rest_density
    int power_is_4;    // eos_power == 4 exactly: (x*x)*(x*x) instead of powf
    float dtg[3];      // dt * gravity (3d:255)
    float clip_lo[3], clip_hi[3];
    float wall_lo[3], wall_hi[3];   // clip +- damp (3d:321-324)
};

// ---- exact integer rules ----------------------------------------------------------------

// Rust `f as i32`: truncating, saturating, NaN -> 0.  __float2int_rz saturates and maps NaN to 0.
__device__ __forceinline__ int rust_as_i32(float f) { return __float2int_rz(f); }

// f32::div_euclid (3d:399): q = trunc(a / b); if a % b < 0 { q - 1 } (b > 0 here).
__device__ __forceinline__ float rust_div_euclid(float a, float b) {
    float q = truncf(__fdiv_rn(a, b));
    if (fmodf(a, b) < 0.0f) q = (b > 0.0f) ? q - 1.0f : q + 1.0f;
    return q;
}

__device__ __forceinline__ int block_key(float p, float res_f) {
    return rust_as_i32(rust_div_euclid(p, res_f));
}

// Same key from the integer cell: for every f32 pos and integer grid_res,
//   key_from_pos(pos) == floor_div(floor(pos) as i32, grid_res)
// because the rounded quotient pos/res never crosses a block face (ulp(res*k)/res > ulp(k)/2);
// the CPU suite checks the identity one ulp either side of every face
// (test_key_equals_floor_div_of_cell) and the GPU suite checks this function against block_key()
// on the device (test_fast_key_matches_exact_key).  Saturated cells (|pos| >= 2^31, inf, the drop tombstone) give keys
// near +-2^27, outside any rect set_rect accepts, exactly like the exact rule.  The hot kernels
// use this (shift or integer division) instead of IEEE division + fmodf per axis.
__device__ __forceinline__ int block_key_of_cell(const int cell, const int res_i, const int res_shift) {
    if (res_shift >= 0) return cell >> res_shift;          // arithmetic shift = floor division
    int q = cell / res_i;
    if ((cell % res_i) < 0) --q;
    return q;
}

enum ParticleClass : int { CLS_ACTIVE = 0, CLS_FROZEN = 1, CLS_LIMBO = 2, CLS_DROPPED = 3 };

// Class of a particle from its position alone: the reference stores a particle in the block
// whose key is key_from_pos(pos) (3d:104-108, 347-366), advances a_rect blocks (3d:263),
// deposits p_rect blocks (3d:149) and never touches the rest.
template <int DIM>
__device__ __forceinline__ int classify(const Geo& g, const float* pos, int* key) {
    bool in_a = true, in_p = true;
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
        int k = block_key(pos[a], g.res_f);
        key[a] = k;
        in_a = in_a && (k >= g.a_lo[a]) && (k < g.a_hi[a]);
        in_p = in_p && (k >= g.p_lo[a]) && (k < g.p_hi[a]);
    }
    return in_a ? CLS_ACTIVE : (in_p ? CLS_FROZEN : CLS_LIMBO);
}

// classify() for the hot kernels: keys from the integer cells (see block_key_of_cell).
template <int DIM>
__device__ __forceinline__ int classify_cells(const Geo& g, const int* cell) {
    bool in_a = true, in_p = true;
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
        int k = block_key_of_cell(cell[a], g.res_i, g.res_shift);
        in_a = in_a && (k >= g.a_lo[a]) && (k < g.a_hi[a]);
        in_p = in_p && (k >= g.p_lo[a]) && (k < g.p_hi[a]);
    }
    return in_a ? CLS_ACTIVE : (in_p ? CLS_FROZEN : CLS_LIMBO);
}

template <int DIM>
__device__ __forceinline__ int classify_pos(const Geo& g, const float* pos) {
    int cell[3];
#pragma unroll
    for (int a = 0; a < DIM; ++a) cell[a] = rust_as_i32(floorf(pos[a]));
    return classify_cells<DIM>(g, cell);
}

// Cell index inside an 8x8x4 tile: class-major.  The bank class of a cell is the 16-byte bank
// group (x + 2y) mod 8 of its node column in the shared-memory tile (index x + 10y + 104z); cells
// are numbered class, then y (which fixes the column inside the class), then z: 32 cells per class,
// 4 per column.  k_tile_tables and k_build_src build windows that hold each class evenly (sort.cuh).
__device__ __forceinline__ int local_cell_3d(int lx, int ly, int lz) {
    const int cls = (lx + 2 * ly) & 7;
    return lz + 4 * (ly + 8 * cls);
}

// Tiled cell index (sort key).  rel = cell - origin, clamped into the grid by the caller.
template <int DIM>
__device__ __forceinline__ int tiled_cell_index(const Geo& g, const int* rel) {
    using T = Tile<DIM>;
    int tx = rel[0] / T::X, lx = rel[0] - tx * T::X;
    int ty = rel[1] / T::Y, ly = rel[1] - ty * T::Y;
    if (DIM == 2) {
        return (ty * g.tdim[0] + tx) * T::CELLS + local_cell_3d(lx, ly, 0);
    }
    int tz = rel[2] / T::Z, lz = rel[2] - tz * T::Z;
    return ((tz * g.tdim[1] + ty) * g.tdim[0] + tx) * T::CELLS + local_cell_3d(lx, ly, lz);
}

// Reference linear node index x + y*sx + z*sx*sy (3d:169-172).
template <int DIM>
__device__ __forceinline__ int ref_cell_index(const Geo& g, const int* rel) {
    int idx = rel[0] + rel[1] * g.size[0];
    if (DIM == 3) idx += rel[2] * g.size[0] * g.size[1];
    return idx;
}

// Index of node (x, y, z) (relative to the grid origin, inside the grid) in the node arrays: the reference's
// linear index x + y*sx + z*sx*sy behind the guard (3d:169-172), or, block-sparse, 256 * pool block + the node's
// place in its 8x8x4 block; -1 if that block has no storage (it cannot receive a deposit then: every block an
// active tile's stencils reach is allocated by the sort).
__device__ __forceinline__ int node_addr(const Geo& g, int x, int y, int z) {
    if (!g.sp.blk) return g.guard + x + (y + z * g.size[1]) * g.size[0];
    const int t = ((z >> 2) * g.tdim[1] + (y >> 3)) * g.tdim[0] + (x >> 3);
    const int b = __ldg(&g.sp.blk[t]);
    return b < 0 ? -1 : (b << 8) + (x & 7) + ((y & 7) << 3) + ((z & 3) << 6);
}

// ---- quadratic stencil (3d:153-161, 390-396) ----------------------------------------------

template <int DIM>
struct Stencil {
    int base[DIM];      // node of offset 0, relative to the grid origin: cell - 1 - org
    float w[DIM][3];    // per-axis weights, zeroed where the node is outside the p_rect grid
    float d[DIM][3];    // x_node - x_particle per axis = (o - 1) - c
};

template <int DIM>
__device__ __forceinline__ void make_stencil(const Geo& g, const float* pos, Stencil<DIM>& s) {
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
        float fl = floorf(pos[a]);
        int cell = rust_as_i32(fl);
        float c = pos[a] - (fl + 0.5f);
        float m = 0.5f - c, p = 0.5f + c;
        s.w[a][0] = 0.5f * m * m;
        s.w[a][1] = 0.75f - c * c;
        s.w[a][2] = 0.5f * p * p;
        s.d[a][0] = -1.0f - c;
        s.d[a][1] = -c;
        s.d[a][2] = 1.0f - c;
        int b = cell - 1 - g.org[a];
        s.base[a] = b;
        // The reference skips nodes outside the p_rect grid (3d:166-168); a zero weight drops
        // the same terms from every sum.
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            int n = b + o;
            if (n < 0 || n >= g.size[a]) s.w[a][o] = 0.0f;
        }
    }
}

// Tait pressure with the reference's lower clamp (3d:217-220): max(clamp, B*((rho/rho0)^g - 1)).
__device__ __forceinline__ float tait_pressure(const Geo& g, float density) {
    float eos = g.stiffness * (powf(__fdiv_rn(density, g.rest_density), g.power) - 1.0f);
    return fmaxf(g.pclamp, eos);
}

// Same, for the hot kernel: the reference's default exponent 4 (3d:27) is two multiplies, and
// rho / rho0 becomes rho * (1 / rho0); both differ from powf / true division by at most a couple
// of ulp, far inside the 1e-5 parity bound.  Any other exponent takes the powf path.
__device__ __forceinline__ float tait_pressure_fast(const Geo& g, float density) {
    const float x = density * g.inv_rest;
    float xp;
    if (g.power_is_4) {
        const float x2 = x * x;
        xp = x2 * x2;
    } else {
        xp = powf(x, g.power);
    }
    return fmaxf(g.pclamp, g.stiffness * (xp - 1.0f));
}

}  // namespace fluid
