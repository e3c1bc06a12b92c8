// sort.cuh — uniform-grid neighbour search: cell keys, counting sort, reorder, cellStart table.
//
// GPU stand-in for the reference's `particles_mul: AHashMap<block key, Vec<Particle>>` and its
// per-block migration mailboxes (3d:52,56,104-108,345-380).  Every substep:
//   1. k_classify_count: block key (div_euclid, 3d:398-401) -> class; cell = floor(pos)
//      (3d:153) -> tiled cell index; rank = atomicAdd(count[cell], 1)
//   2. exclusive scan of count -> cellStart (hand-written 3-kernel scan, no CUB)
//   3. k_reorder: dst = cellStart[cell] + rank; all SoA streams move to the other buffer
// Buckets n_cells_pad and n_cells_pad+1 collect the particles the reference ignores (key
// outside p_rect) and the ones it dropped in migration (3d:356-366).
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

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;   // 2048 counts per block

__device__ __forceinline__ bool is_tombstone(float x) { return isinf(x) && x > 0.0f; }

template <int DIM>
__global__ void __launch_bounds__(256)
k_classify_count(const __grid_constant__ Geo g, const float4* __restrict__ P, int n,
                 int* __restrict__ cell_idx, int* __restrict__ rank, int* __restrict__ count,
                 int* __restrict__ class_count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int cls = -1;
    if (i < n) {
        float4 p = P[i];
        float pos[3] = {p.x, p.y, p.z};
        int bucket;
        if (is_tombstone(p.x)) {
            cls = CLS_DROPPED;
            bucket = g.n_cells_pad + 1;
        } else {
            int cell[3] = {0, 0, 0};
#pragma unroll
            for (int a = 0; a < DIM; ++a) cell[a] = rust_as_i32(floorf(pos[a]));
            cls = classify_cells<DIM>(g, cell);
            if (cls == CLS_LIMBO) {
                bucket = g.n_cells_pad;
            } else {
                int rel[3] = {0, 0, 0};
#pragma unroll
                for (int a = 0; a < DIM; ++a)   // in p_rect => inside the grid; clamp guards smem only
                    rel[a] = min(max(cell[a] - g.org[a], 0), g.size[a] - 1);
                bucket = tiled_cell_index<DIM>(g, rel);
            }
        }
        cell_idx[i] = bucket;
        rank[i] = atomicAdd(&count[bucket], 1);
    }
    // class counters: one atomic per warp and class
    unsigned full = 0xffffffffu;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        unsigned m = __ballot_sync(full, cls == c);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&class_count[c], __popc(m));
    }
}

// ---- exclusive scan over the count array -----------------------------------------------

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
k_scan_partial(const int* __restrict__ count, int m, int* __restrict__ block_sums) {
    int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
    int s = 0;
    if (base + SCAN_ITEMS <= m) {
        const int4* q = reinterpret_cast<const int4*>(count + base);
        int4 a = q[0], b = q[1];
        s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    } else {
        for (int k = 0; k < SCAN_ITEMS; ++k)
            if (base + k < m) s += count[base + k];
    }
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

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_final(const int* __restrict__ count, int m, const int* __restrict__ block_sums,
             int* __restrict__ start) {
    int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < m) ? count[base + k] : 0;
        s += v[k];
    }
    int total;
    int ex = block_exclusive_scan(s, total) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < m) start[base + k] = ex;
        ex += v[k];
    }
    // start[m] = grand total (written by the thread that owns the last element)
    if (base <= m - 1 && m - 1 < base + SCAN_ITEMS) start[m] = ex;
}

// ---- reorder ---------------------------------------------------------------------------

template <int DIM>
__global__ void __launch_bounds__(256)
k_reorder(Particles src, Particles dst, int n, const int* __restrict__ cell_idx,
          const int* __restrict__ rank, const int* __restrict__ start) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d = start[cell_idx[i]] + rank[i];
    dst.P[d] = src.P[i];
    dst.V[d] = src.V[i];
    dst.CA[d] = src.CA[i];
    if (DIM == 3) {
        dst.CB[d] = src.CB[i];
        dst.CC[d] = src.CC[i];
    }
}

}  // namespace fluid
