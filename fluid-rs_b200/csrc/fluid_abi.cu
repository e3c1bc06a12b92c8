// fluid_abi.cu — C ABI (include/fluid_b200.h) and host orchestration of the step engine.
//
// Drop-in for fluid-rs's `Simulation` (src/3d_multi.rs:50-134,383-387; 2d_multi.rs same
// lines).  All state is device resident; the host only sequences kernels on one stream.
// There is no CPU fallback: without a CUDA device fluid_create fails with FLUID_ERR_NO_DEVICE.
#include "../../include/fluid_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "phases_generic.cuh"
#include "phases_tiled.cuh"
#include "phases_tiled2d.cuh"
#include "sort.cuh"

using namespace fluid;

namespace {

thread_local std::string g_last_error;

fluid_status fail(fluid_status st, const std::string& msg) {
    g_last_error = msg;
    return st;
}

#define CU_TRY(expr)                                                                      \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            (void)cudaGetLastError();                                                     \
            return fail(_e == cudaErrorMemoryAllocation ? FLUID_ERR_OUT_OF_MEMORY         \
                                                        : FLUID_ERR_CUDA,                 \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));              \
        }                                                                                 \
    } while (0)

#define ST_TRY(expr)                          \
    do {                                      \
        fluid_status _s = (expr);             \
        if (_s != FLUID_OK) return _s;        \
    } while (0)

inline int rec_floats(int dim) { return 2 * dim + dim * dim + 1; }
inline unsigned blocks_for(int64_t n, int threads) {
    return static_cast<unsigned>((n + threads - 1) / threads);
}

// host copy of the device integer rules (set_rect needs key_from_pos on the host, 3d:80-81)
int host_block_key(float p, float res) {
    float q = std::trunc(p / res);
    if (std::fmod(p, res) < 0.0f) q = res > 0.0f ? q - 1.0f : q + 1.0f;
    if (q != q) return 0;
    if (q >= 2147483648.0f) return INT_MAX;
    if (q <= -2147483648.0f) return INT_MIN;
    return static_cast<int>(q);
}

constexpr int N_EVENTS = 6;   // sort | clear | p2g1 | p2g2 | g2p | end
constexpr int PROFILE_POOL = 1024;   // substeps between two drains (a drain waits for the stream: keep it out of short timed regions)

}  // namespace

struct fluid_sim {
    fluid_config cfg{};
    int dim = 3;
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;

    Geo geo{};
    bool rect_set = false;

    int64_t n = 0;          // particle slots in use (live + not yet compacted tombstones)
    int64_t cap = 0;
    Particles buf[2]{};
    ParticleTex tex[2]{};    // the same arrays (and src) as linear textures: gathers through the TEX pipe (sort.cuh)
    bool tex_on = true;      // FLUID_B200_TEX=0 keeps the gathers on __ldg
    int cur = 0;
    // neighbour search (sort.cuh)
    int* gcell = nullptr;    // per particle: bucket (tile * 256 + cell in tile)
    int* rank = nullptr;     // per particle: rank inside its bucket
    int* src = nullptr;      // sorted slot -> storage index in buf[cur]
    int* imm_list = nullptr; // particles that changed tile in the last g2p (per-tile lists)
    int* imm_cnt = nullptr;  // leavers per active-tile-list entry
    int4* tiles = nullptr;   // active tile list {tile, first slot, count, 0}, rebuilt by every sort
    int* scal = nullptr;     // device scalars: [0] active tiles, [1] immigrants
    int* cell_off = nullptr; // per bucket: first cell-sorted slot (cellStart)
    int* tile_total = nullptr;
    int* tile_base = nullptr;    // exclusive scan of tile_total; [n_tiles] = number of p_rect particles
    int* cand = nullptr;         // tiles that hold particles (from the scan), input list of k_tile_tables
    int* dirty_list = nullptr;   // node blocks to clear this substep
    unsigned char* dirty[2] = {nullptr, nullptr};   // node blocks touched by the current / previous sort
    int dirty_cur = 0;
    int2* tile_info = nullptr;   // per tile: {windows W (0 = plain cell order, < 0 = no class merge), tile-list entry}
    unsigned char* tab = nullptr;   // per tile-list entry: TAB_BYTES of class-in-window counts (k_tile_tables)
    alignas(64) CUtensorMap tm_grid;   // the node grid as a TMA tensor {4 floats, x, y, z}, box = one 10x10 footprint plane
    alignas(64) CUtensorMap tm_mass;   // the node masses {x + 1, y, z} (base one float early: 16-byte aligned), box 12 x 10 x 6
    bool tma_mass = false;
    bool tma = false;            // tm_grid is valid (3D grids; FLUID_B200_NO_TMA=1 keeps the LDGSTS / REDG paths)
    PeerHalo peer{};             // neighbours' grids mapped through CUDA IPC (peer-memory halo), or all null
    void* peer_base[2][6] = {{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr},
                             {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}};   // opened IPC mappings
    float* mig_recv[2] = {nullptr, nullptr};   // records the neighbours write here over NVLink (header record first)
    int* flags = nullptr;        // [0], [1]: arrival flags the lower / upper neighbour writes (peer barrier); [2]: time-out
    int barrier_epoch = 0;
    bool p2p = false;            // deposits into the shared node planes go to the neighbour directly
    int* gz = nullptr;           // per tile: substep number in which k_g2p_tiled zeroed its node-mass block
    int* d_epoch = nullptr;      // device: tiled substeps completed so far (compared with gz; advanced by k_tail)
    bool grid_clean = false;     // every node outside the dirty blocks is zero
    bool sorted_valid = false;   // arrays are in tile order for the current positions
    bool counts_pending = false; // the last g2p left buckets / ranks / counts for the next sort
    int tile_order = ORDER_CLASS_RR;  // window order of the 3D tiled path (FLUID_B200_ORDER=q: ORDER_CLASS_Q, sort.cuh)
    int fixed8[3] = {0, 4, 0};        // ... after this many eighths of the list dealt with the fixed stride (FLUID_B200_FIXED8=mpg digits)
    int dyn_tiles = 6;                // tile kernels that take their tiles from a ticket counter: 1 k_mass_tiled, 2 k_p2g_tiled, 4 k_g2p_tiled (FLUID_B200_DYN=mask; 0: fixed stride)
    int sm_count = 148;
    unsigned grid_mass = 0, grid_p2g = 0, grid_g2p = 0;   // persistent grids: SMs x resident CTAs
    unsigned grid2_mass = 0, grid2_p2g = 0, grid2_g2p = 0;   // ... of the 2D tile kernels

    int* count = nullptr;    // per bucket, (n_tiles + 2) * 256; all zero outside a sort
    int* block_sums = nullptr;
    int64_t n_scan_blocks = 0;
    int* class_count = nullptr;   // 4 ints (device)

    float4* grid = nullptr;      // {momentum.xyz, mass} per node, reference layout + guards
    int64_t sparse_blocks = 0;   // > 0: block-sparse node storage with this many pool blocks (takes effect at set_rect)
    int64_t pool_blocks = 0;     // pool blocks of the current grid (0 = dense)
    int64_t node_alloc = 0;      // elements of `grid` / `gmass`: dense nodes + guards, or 256 * pool blocks
    int* h_pool = nullptr;       // pinned: {free blocks, exhausted flag}
    float4* grid2 = nullptr;     // second node buffer of the resident small-scene kernel (allocated on first use)
    long long* fx[2] = {nullptr, nullptr};   // deterministic mode: 4 fixed-point sums per node (phases_generic.cuh)
    bool det = false;            // FLUID_B200_DETERMINISTIC=1 / fluid_set_deterministic: order-independent node sums
    bool grid_is_fixed = false;  // the last substep left its node sums in fx[0], not in `grid`
    int64_t resident_max = 16384;   // step() of at most this many particles runs as ONE cooperative launch
    bool coop = false;           // device supports cooperative launches
    unsigned long long* d_stamps = nullptr;   // %globaltimer stamps of the resident kernel's last substep
    bool resident_timed = false; // the phase timers of the last substep are in d_stamps, not in events
    float* gmass = nullptr;      // node masses alone (p2g 1 output, read by p2g 2), same layout
    int64_t grid_nodes = 0;      // reference node count (without guards)
    bool tiled = true;           // 3D: tiled sm_100a kernels; FLUID_B200_GENERIC=1 forces the generic ones
    float* d_mouse = nullptr;    // {x, y, present}
    float h_mouse[4] = {0, 0, 0, 0};
    bool mouse_on = false;

    float* d_stage = nullptr;    // staging for host<->device record copies
    int64_t stage_floats = 0;
    int* d_stage_ids = nullptr;
    int64_t stage_ids = 0;
    int* d_counter = nullptr;

    cudaEvent_t ev[N_EVENTS]{};
    bool timers_recorded = false;
    cudaEvent_t* last_ev = nullptr;    // event set of the last timed substep
    cudaEvent_t* cur_ev = nullptr;     // event set of the substep in progress (split substeps)
    bool cur_timed = false;
    // z-slab decomposition
    float* mig_rec[2] = {nullptr, nullptr};   // a one-record header {count}, then the packed records of the particles leaving through z_lo / z_hi
    int* d_status = nullptr;                  // device: scal[0..7], tile_base[n_tiles .. n_tiles + 3] gathered for one read-back
    int* h_status = nullptr;                  // pinned host copy: + the headers of the two received buffers
    int mig_cap = 0;
    bool has_nb[2] = {false, false};
    float* halo_mass_recv[2] = {nullptr, nullptr};
    float4* halo_node_recv[2] = {nullptr, nullptr};
    long long* halo_fx_recv[2] = {nullptr, nullptr};   // deterministic mode: the neighbour's fixed-point planes
    // profile mode: a pool of event sets, drained into sums when full or when read
    bool profiling = false;
    std::vector<cudaEvent_t> pool;     // PROFILE_POOL * N_EVENTS
    int pool_used = 0;
    double prof_sum[N_EVENTS - 1] = {0, 0, 0, 0, 0};
    int64_t prof_substeps = 0;

    // CUDA graphs of one steady-state substep, one per (particle buffer, dirty-flag buffer) parity
    bool graphs_on = true;                  // FLUID_B200_GRAPH=0 launches every kernel individually
    cudaGraphExec_t graph[4] = {nullptr, nullptr, nullptr, nullptr};
    int graph_nodes[4] = {0, 0, 0, 0};      // kernels per graph (launch accounting)
    int64_t graph_n = -1;                   // the particle count the graphs were captured for
    int64_t graph_min = 0;                  // (scenes of at most resident_max particles take the resident kernel instead)

    int64_t launches = 0;
    int32_t next_id = 0;
    int64_t dropped_total = 0;
};

namespace {

void graphs_drop(fluid_sim* s);   // (defined with the substep)

void free_particles(Particles& p) {
    cudaFree(p.P);
    cudaFree(p.V);
    cudaFree(p.CA);
    cudaFree(p.CB);
    cudaFree(p.CC);
    p = Particles{};
}

fluid_status alloc_particles(Particles& p, int64_t cap) {
    CU_TRY(cudaMalloc(&p.P, cap * sizeof(float4)));
    CU_TRY(cudaMalloc(&p.V, cap * sizeof(float4)));
    CU_TRY(cudaMalloc(&p.CA, cap * sizeof(float4)));
    CU_TRY(cudaMalloc(&p.CB, cap * sizeof(float4)));
    CU_TRY(cudaMalloc(&p.CC, cap * sizeof(float)));
    return FLUID_OK;
}

void free_textures(fluid_sim* s) {
    for (int b = 0; b < 2; ++b) {
        cudaTextureObject_t* t[6] = {&s->tex[b].P, &s->tex[b].V, &s->tex[b].CA, &s->tex[b].CB, &s->tex[b].CC, &s->tex[b].src};
        for (int k = 0; k < 6; ++k) {
            if (*t[k] && !(b == 1 && k == 5)) cudaDestroyTextureObject(*t[k]);   // (src is one object, held by both)
            *t[k] = 0;
        }
    }
    (void)cudaGetLastError();
}

template <typename T>
cudaTextureObject_t linear_texture(const T* ptr, int64_t n, int device) {
    const cudaChannelFormatDesc fmt = cudaCreateChannelDesc<T>();
    size_t max_elems = 0;
    if (cudaDeviceGetTexture1DLinearMaxWidth(&max_elems, &fmt, device) != cudaSuccess || static_cast<size_t>(n) > max_elems) {
        (void)cudaGetLastError();
        return 0;
    }
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = const_cast<T*>(ptr);
    rd.res.linear.desc = fmt;
    rd.res.linear.sizeInBytes = static_cast<size_t>(n) * sizeof(T);
    cudaTextureDesc td{};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t t = 0;
    if (cudaCreateTextureObject(&t, &rd, &td, nullptr) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return t;
}

// all or nothing: the kernels test one handle per stream, but a partial set would only complicate the fallback
void make_textures(fluid_sim* s) {
    free_textures(s);
    if (!s->tex_on || s->dim != 3 || s->cap <= 0) return;
    const cudaTextureObject_t ts = linear_texture(s->src, s->cap, s->device);
    bool ok = ts != 0;
    for (int b = 0; b < 2 && ok; ++b) {
        s->tex[b].P = linear_texture(s->buf[b].P, s->cap, s->device);
        s->tex[b].V = linear_texture(s->buf[b].V, s->cap, s->device);
        s->tex[b].CA = linear_texture(s->buf[b].CA, s->cap, s->device);
        s->tex[b].CB = linear_texture(s->buf[b].CB, s->cap, s->device);
        s->tex[b].CC = linear_texture(s->buf[b].CC, s->cap, s->device);
        s->tex[b].src = ts;
        ok = s->tex[b].P && s->tex[b].V && s->tex[b].CA && s->tex[b].CB && s->tex[b].CC;
    }
    if (!ok) {
        if (ts && !s->tex[0].src) cudaDestroyTextureObject(ts);
        free_textures(s);
    }
}

fluid_status ensure_capacity(fluid_sim* s, int64_t want) {
    if (want <= s->cap) return FLUID_OK;
    if (want > INT_MAX - 1024) return fail(FLUID_ERR_INVALID_ARG, "more than 2^31 particles per handle");
    int64_t cap = std::max<int64_t>(want, s->cap + s->cap / 2);
    cap = std::min<int64_t>((cap + 1023) / 1024 * 1024, INT_MAX - 1024);
    // every new buffer first; the handle is only touched once all of them exist
    Particles nb[2]{};
    int* tabs[4] = {nullptr, nullptr, nullptr, nullptr};   // gcell, rank, imm_list, src
    fluid_status st = FLUID_OK;
    for (int b = 0; b < 2 && st == FLUID_OK; ++b) st = alloc_particles(nb[b], cap);
    for (int k = 0; k < 4 && st == FLUID_OK; ++k)
        if (cudaMalloc(&tabs[k], cap * sizeof(int)) != cudaSuccess) {
            (void)cudaGetLastError();
            st = fail(FLUID_ERR_OUT_OF_MEMORY, "ensure_capacity: device allocation failed");
        }
    if (st == FLUID_OK && s->n > 0) {
        const Particles& o = s->buf[s->cur];
        cudaError_t e = cudaMemcpyAsync(nb[0].P, o.P, s->n * sizeof(float4), cudaMemcpyDeviceToDevice, s->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(nb[0].V, o.V, s->n * sizeof(float4), cudaMemcpyDeviceToDevice, s->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(nb[0].CA, o.CA, s->n * sizeof(float4), cudaMemcpyDeviceToDevice, s->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(nb[0].CB, o.CB, s->n * sizeof(float4), cudaMemcpyDeviceToDevice, s->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(nb[0].CC, o.CC, s->n * sizeof(float), cudaMemcpyDeviceToDevice, s->stream);
        if (e != cudaSuccess) st = fail(FLUID_ERR_CUDA, std::string("ensure_capacity: ") + cudaGetErrorString(e));
    }
    if (st == FLUID_OK && cudaStreamSynchronize(s->stream) != cudaSuccess) st = fail(FLUID_ERR_CUDA, "ensure_capacity: copy failed");
    if (st != FLUID_OK) {   // the handle keeps its old buffers and capacity
        for (int b = 0; b < 2; ++b) free_particles(nb[b]);
        for (int k = 0; k < 4; ++k) cudaFree(tabs[k]);
        (void)cudaGetLastError();
        return st;
    }
    graphs_drop(s);   // the captured launches hold the old pointers
    free_textures(s);
    for (int b = 0; b < 2; ++b) free_particles(s->buf[b]);
    cudaFree(s->gcell);
    cudaFree(s->rank);
    cudaFree(s->imm_list);
    cudaFree(s->src);
    s->buf[0] = nb[0];
    s->buf[1] = nb[1];
    s->cur = 0;
    s->gcell = tabs[0];
    s->rank = tabs[1];
    s->imm_list = tabs[2];
    s->src = tabs[3];
    s->sorted_valid = s->counts_pending = false;
    s->cap = cap;
    make_textures(s);
    return FLUID_OK;
}

fluid_status ensure_stage(fluid_sim* s, int64_t floats, int64_t ids) {
    if (floats > s->stage_floats) {
        cudaFree(s->d_stage);
        s->d_stage = nullptr;
        s->stage_floats = 0;
        CU_TRY(cudaMalloc(&s->d_stage, floats * sizeof(float)));
        s->stage_floats = floats;
    }
    if (ids > s->stage_ids) {
        cudaFree(s->d_stage_ids);
        s->d_stage_ids = nullptr;
        s->stage_ids = 0;
        CU_TRY(cudaMalloc(&s->d_stage_ids, ids * sizeof(int)));
        s->stage_ids = ids;
    }
    return FLUID_OK;
}

// ---- small utility kernels -------------------------------------------------------------

template <int DIM>
__global__ void k_unpack_records(const float* __restrict__ rec, const int* __restrict__ ids,
                                 int base_id, int n, Particles dst, int offset) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int RF = 2 * DIM + DIM * DIM + 1;
    const float* r = rec + static_cast<int64_t>(i) * RF;
    int id = ids ? ids[i] : base_id + i;
    int d = offset + i;
    if (DIM == 3) {
        dst.P[d] = make_float4(r[0], r[1], r[2], r[15]);
        dst.V[d] = make_float4(r[3], r[4], r[5], __int_as_float(id));
        dst.CA[d] = make_float4(r[6], r[7], r[8], r[9]);
        dst.CB[d] = make_float4(r[10], r[11], r[12], r[13]);
        dst.CC[d] = r[14];
    } else {
        dst.P[d] = make_float4(r[0], r[1], 0.0f, r[8]);
        dst.V[d] = make_float4(r[2], r[3], 0.0f, __int_as_float(id));
        dst.CA[d] = make_float4(r[4], r[5], r[6], r[7]);
    }
}

template <int DIM>
__device__ __forceinline__ void pack_record(const Particles& q, int i, float* r, int* id) {
    float4 p = q.P[i], v = q.V[i], a = q.CA[i];
    if (DIM == 3) {
        float4 b = q.CB[i];
        r[0] = p.x; r[1] = p.y; r[2] = p.z;
        r[3] = v.x; r[4] = v.y; r[5] = v.z;
        r[6] = a.x; r[7] = a.y; r[8] = a.z; r[9] = a.w;
        r[10] = b.x; r[11] = b.y; r[12] = b.z; r[13] = b.w;
        r[14] = q.CC[i];
        r[15] = p.w;
    } else {
        r[0] = p.x; r[1] = p.y; r[2] = v.x; r[3] = v.y;
        r[4] = a.x; r[5] = a.y; r[6] = a.z; r[7] = a.w;
        r[8] = p.w;
    }
    *id = __float_as_int(v.w);
}

// iter_particle (3d:383-387): every particle stored in an a_rect block, compacted.
template <int DIM>
__global__ void k_pack_active(const __grid_constant__ Geo g, Particles q, int n,
                              float* __restrict__ rec, int* __restrict__ ids,
                              int* __restrict__ counter, int capacity) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool take = false;
    if (i < n) {
        float4 p = q.P[i];
        int cls = -1;
        (void)bucket_of<DIM>(g, p, cls);   // same rule as the sort (tombstones, rects, slab ownership)
        take = cls == CLS_ACTIVE;
    }
    unsigned m = __ballot_sync(0xffffffffu, take);
    int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (take) {
        int d = base + __popc(m & ((1u << lane) - 1));
        if (d < capacity) {
            constexpr int RF = 2 * DIM + DIM * DIM + 1;
            int id;
            pack_record<DIM>(q, i, rec + static_cast<int64_t>(d) * RF, &id);
            if (ids) ids[d] = id;
        }
    }
}

// ids / cell / key / reference cell index of the sorted p_rect particles (parity taps).
template <int DIM>
__global__ void k_debug_keys(const __grid_constant__ Geo g, Particles q, const int* __restrict__ src,
                             const int* __restrict__ n_deposit, int* __restrict__ ids,
                             int* __restrict__ cell, int* __restrict__ key,
                             int* __restrict__ ref_index) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;   // sorted slot
    if (i >= *n_deposit) return;
    const int j = src[i];                            // storage index
    float4 p = q.P[j];
    float pos[3] = {p.x, p.y, p.z};
    int k[3], rel[3] = {0, 0, 0};
    classify<DIM>(g, pos, k);
    for (int a = 0; a < DIM; ++a) {
        int c = rust_as_i32(floorf(pos[a]));
        if (cell) cell[i * DIM + a] = c;
        // exact rule (3d:398-401); if the hot kernels' integer rule ever disagreed, report the
        // integer rule's key so the bit-exact parity tests fail loudly
        int kf = block_key_of_cell(c, g.res_i, g.res_shift);
        if (key) key[i * DIM + a] = (kf == k[a]) ? k[a] : kf;
        rel[a] = c - g.org[a];
    }
    if (ids) ids[i] = __float_as_int(q.V[j].w);
    if (ref_index) ref_index[i] = ref_cell_index<DIM>(g, rel);
}

// ---- the substep -----------------------------------------------------------------------

struct DebugTaps {
    int* ids = nullptr;
    int* cell = nullptr;
    int* key = nullptr;
    float* density = nullptr;
    float* pressure = nullptr;
};

SortTables sort_tables(fluid_sim* s) {
    SortTables t;
    t.gcell = s->gcell;
    t.rank = s->rank;
    t.count = s->count;
    t.tile_total = s->tile_total;
    t.imm_list = s->imm_list;
    t.imm_cnt = s->imm_cnt;
    t.scal = s->scal;
    return t;
}

// scan over the tile totals, per-tile order, reorder (count[], tile_total[], gcell[], rank[] are filled)
template <int DIM>
fluid_status sort_finish(fluid_sim* s) {
    const int n = static_cast<int>(s->n);
    const int m = s->geo.n_tiles + N_PSEUDO;
    const unsigned nb = static_cast<unsigned>(s->n_scan_blocks);
    CU_TRY(cudaMemsetAsync(s->scal, 0, SCAL_COUNT * sizeof(int), s->stream));
    k_scan_partial<<<nb, SCAN_THREADS, 0, s->stream>>>(s->tile_total, m, s->block_sums);
    k_scan_sums<<<1, 1024, 0, s->stream>>>(s->block_sums, static_cast<int>(nb));
    k_scan_final<<<nb, SCAN_THREADS, 0, s->stream>>>(s->tile_total, m, s->block_sums, s->tile_base, s->cand,
                                                    s->scal + SCAL_N_CAND);
    // persistent warps over the candidate list: one resident wave of small CTAs
    const unsigned pb = std::min<unsigned>(blocks_for(static_cast<int64_t>(m) * 32, PERM_WARPS * 32),
                                           static_cast<unsigned>(s->sm_count * 16));
    if (DIM == 3 || s->tiled) {   // (2D tiles are 8 x 8 columns of depth 1: the same order and tables)
        s->dirty_cur ^= 1;   // the buffer k_clear_tiles emptied last substep
        if (DIM == 3 && s->tile_order == ORDER_CLASS_Q)
            k_tile_tables<ORDER_CLASS_Q><<<pb, PERM_WARPS * 32, 0, s->stream>>>(s->geo, s->count, s->tile_base, s->cell_off, s->tile_info, s->tab,
                                                                               s->tiles, s->scal, s->dirty[s->dirty_cur], s->cand, s->peer);
        else
            k_tile_tables<ORDER_CLASS_RR><<<pb, PERM_WARPS * 32, 0, s->stream>>>(s->geo, s->count, s->tile_base, s->cell_off, s->tile_info, s->tab,
                                                                                s->tiles, s->scal, s->dirty[s->dirty_cur], s->cand, s->peer);
    } else {
        k_tile_tables<ORDER_CELL><<<pb, PERM_WARPS * 32, 0, s->stream>>>(s->geo, s->count, s->tile_base, s->cell_off, s->tile_info, s->tab,
                                                                        s->tiles, s->scal, nullptr, s->cand, PeerHalo{});
    }
    s->launches += 4;
    if (n > 0) {
        k_build_src<<<blocks_for(n, 256 * BUILD_SRC_PER_THREAD), 256, 0, s->stream>>>(n, s->gcell, s->rank, s->cell_off, s->tile_base, s->tile_info, s->tab, s->src);
        ++s->launches;
    }
    CU_TRY(cudaGetLastError());
    s->sorted_valid = true;
    s->counts_pending = false;
    return FLUID_OK;
}

// cold start: every particle classified and counted with global atomics
template <int DIM>
fluid_status sort_cold(fluid_sim* s) {
    const int n = static_cast<int>(s->n);
    const int64_t buckets = static_cast<int64_t>(s->geo.n_tiles + N_PSEUDO) * TILE_CELLS;
    CU_TRY(cudaMemsetAsync(s->count, 0, buckets * sizeof(int), s->stream));
    CU_TRY(cudaMemsetAsync(s->tile_total, 0, (s->geo.n_tiles + N_PSEUDO) * sizeof(int), s->stream));
    CU_TRY(cudaMemsetAsync(s->class_count, 0, 4 * sizeof(int), s->stream));
    if (n > 0) {
        k_classify_all<DIM><<<blocks_for(n, 256), 256, 0, s->stream>>>(s->geo, s->buf[s->cur].P, n, sort_tables(s),
                                                                      s->class_count);
        ++s->launches;
    }
    return sort_finish<DIM>(s);
}

// steady state: g2p already counted the particles that stayed in their tile
template <int DIM>
fluid_status sort_steady(fluid_sim* s) {
    k_immigrants<<<s->sm_count * 8, 256, 0, s->stream>>>(sort_tables(s), s->tiles, s->geo.n_tiles + N_PSEUDO);
    ++s->launches;
    return sort_finish<DIM>(s);
}

template <int DIM>
fluid_status ensure_sorted(fluid_sim* s) {
    if (!s->sorted_valid) return sort_cold<DIM>(s);
    if (s->counts_pending) return sort_steady<DIM>(s);
    return FLUID_OK;
}

fluid_status profile_drain(fluid_sim* s) {
    if (s->pool_used == 0) return FLUID_OK;
    CU_TRY(cudaEventSynchronize(s->pool[(s->pool_used - 1) * N_EVENTS + N_EVENTS - 1]));
    for (int k = 0; k < s->pool_used; ++k)
        for (int i = 0; i < N_EVENTS - 1; ++i) {
            float ms = 0.0f;
            CU_TRY(cudaEventElapsedTime(&ms, s->pool[k * N_EVENTS + i], s->pool[k * N_EVENTS + i + 1]));
            s->prof_sum[i] += ms * 1e-3;
        }
    s->prof_substeps += s->pool_used;
    s->pool_used = 0;
    return FLUID_OK;
}

// Peer-halo slab runs: the neighbour's "p2g 1" adds into this rank's node masses right after its own sort,
// possibly before this rank's next clear.  So the node masses of the rim blocks are zeroed here, at the end
// of the substep (after g2p, before the migrant exchange both ranks wait on); the blocks of the active
// tiles were zeroed by k_g2p_tiled itself (stamp gz == epoch).  The list is the union of both dirty flag
// arrays (this substep's sort, the neighbour's marks, which go into both); nothing is reset (the next clear
// of the node records needs them).
fluid_status clear_mass_rim(fluid_sim* s, bool fused) {
    k_dirty_list<<<blocks_for(s->geo.n_tiles, 256), 256, 0, s->stream>>>(s->geo, s->dirty[0], s->dirty[1], s->dirty_list,
                                                                        s->scal + SCAL_N_DIRTY2, false);
    k_clear_tiles<<<std::min<unsigned>(blocks_for(static_cast<int64_t>(s->geo.n_tiles) * 32, 128),
                                      static_cast<unsigned>(s->sm_count * 16)), 128, 0, s->stream>>>(
        s->geo, s->dirty_list, s->scal + SCAL_N_DIRTY2, s->grid, s->gmass, s->tile_base, s->gz, s->d_epoch, fused, 2, nullptr);
    s->launches += 2;
    CU_TRY(cudaGetLastError());
    return FLUID_OK;
}

// deterministic mode: the fixed-point node sums (count = 1: generic kernels, 2: the resident kernel's two buffers)
fluid_status ensure_fixed(fluid_sim* s, int count) {
    const int64_t n_alloc = s->node_alloc;
    for (int b = 0; b < count; ++b)
        if (!s->fx[b]) CU_TRY(cudaMalloc(&s->fx[b], n_alloc * 4 * sizeof(long long)));
    return FLUID_OK;
}

// step() / substeps() of a small scene as ONE cooperative launch (k_substeps_resident, phases_generic.cuh).
bool resident_eligible(const fluid_sim* s) {
    return s->coop && s->rect_set && s->n > 0 && s->n <= s->resident_max && !s->geo.slab_on && !s->profiling &&
           !s->pool_blocks;
}

template <int DIM, bool DET>
fluid_status resident_launch(fluid_sim* s, const float* d_mouse, int n_substeps, unsigned blocks) {
    const Geo geo = s->geo;
    Particles q = s->buf[s->cur];
    int n = static_cast<int>(s->n);
    NodeGrid g0{s->grid, s->fx[0]}, g1{s->grid2, s->fx[1]};
    unsigned long long* stamps = s->d_stamps;
    void* args[] = {const_cast<Geo*>(&geo), &q, &n, &g0, &g1, &d_mouse, &n_substeps, &stamps};
    CU_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&k_substeps_resident<DIM, DET>), dim3(blocks), dim3(128), args, 0,
                                       s->stream));
    return FLUID_OK;
}

fluid_status substeps_resident(fluid_sim* s, const float* d_mouse, int n_substeps) {
    if (n_substeps <= 0) return FLUID_OK;
    const int64_t n_alloc = s->node_alloc;
    if (!s->grid2) CU_TRY(cudaMalloc(&s->grid2, n_alloc * sizeof(float4)));
    if (!s->d_stamps) CU_TRY(cudaMalloc(&s->d_stamps, 8 * sizeof(unsigned long long)));
    if (s->det) ST_TRY(ensure_fixed(s, 2));
    // both node buffers clean at the start (the kernel only clears what it wrote itself)
    if (s->det) {
        CU_TRY(cudaMemsetAsync(s->fx[0], 0, n_alloc * 4 * sizeof(long long), s->stream));
        CU_TRY(cudaMemsetAsync(s->fx[1], 0, n_alloc * 4 * sizeof(long long), s->stream));
    } else {
        CU_TRY(cudaMemsetAsync(s->grid, 0, n_alloc * sizeof(float4), s->stream));
        CU_TRY(cudaMemsetAsync(s->grid2, 0, n_alloc * sizeof(float4), s->stream));
    }
    const unsigned blocks = blocks_for(s->n, 128);
    if (s->dim == 3) ST_TRY(s->det ? (resident_launch<3, true>(s, d_mouse, n_substeps, blocks)) : (resident_launch<3, false>(s, d_mouse, n_substeps, blocks)));
    else ST_TRY(s->det ? (resident_launch<2, true>(s, d_mouse, n_substeps, blocks)) : (resident_launch<2, false>(s, d_mouse, n_substeps, blocks)));
    ++s->launches;
    s->grid_is_fixed = s->det;
    s->grid_clean = false;                              // the tiled path wipes the grid before it runs again
    s->sorted_valid = s->counts_pending = false;        // positions moved; nothing was sorted
    s->timers_recorded = true;
    s->resident_timed = true;
    return FLUID_OK;
}

// One substep = clear -> p2g 1 -> p2g 2 -> update -> g2p (3d:111-133).  `phases` selects the parts
// (slab runs exchange halo planes between them): 1 = sort + clear + p2g 1, 2 = p2g 2, 4 = update + g2p.
template <int DIM>
fluid_status substep_impl(fluid_sim* s, const float* d_mouse, bool timed, const DebugTaps* dbg, int phases = 7) {
    if (!s->rect_set || s->n == 0) return FLUID_OK;   // no blocks to walk (3d:149 over an empty rect)
    const int n = static_cast<int>(s->n);
    const int* n_dep = s->tile_base + s->geo.n_tiles;
    const bool tiled = s->tiled;   // 3D: phases_tiled.cuh, 2D: phases_tiled2d.cuh
    // slab runs split the substep at the halo exchanges: the tiled 3D path, or (deterministic mode) the
    // particle-per-thread kernels with fixed-point node sums
    const bool slab_generic = DIM == 3 && !tiled && s->det && s->geo.slab_on;
    if (phases != 7 && !(tiled && DIM == 3) && !slab_generic)
        return fail(FLUID_ERR_STATE, "split substeps need the tiled 3D path or the deterministic mode");
    const unsigned tb2 = blocks_for(s->geo.n_tiles, T2::WARPS);
    if (phases & 1) {
        s->cur_ev = s->ev;
        s->cur_timed = timed;
        if (s->profiling) {
            if (s->pool_used == PROFILE_POOL) ST_TRY(profile_drain(s));
            s->cur_ev = &s->pool[s->pool_used * N_EVENTS];
            ++s->pool_used;
            s->cur_timed = true;
        }
    }
    cudaEvent_t* ev = s->cur_ev;
    timed = s->cur_timed;
    const int64_t n_alloc = s->node_alloc;
    const unsigned tb = blocks_for(s->geo.n_tiles, T3::WARPS);   // never more CTAs than tiles / 4
    const int* n_act = s->scal + SCAL_N_ACTIVE;
    if (phases & 1) {
        if (timed) CU_TRY(cudaEventRecord(ev[0], s->stream));
        ST_TRY(ensure_sorted<DIM>(s));
        Particles q = s->buf[s->cur];
        if (dbg && (dbg->ids || dbg->cell || dbg->key)) {
            k_debug_keys<DIM><<<blocks_for(n, 256), 256, 0, s->stream>>>(s->geo, q, s->src, n_dep, dbg->ids, dbg->cell,
                                                                        dbg->key, nullptr);
            ++s->launches;
        }
        if (timed) CU_TRY(cudaEventRecord(ev[1], s->stream));
        // clear_grid (3d:136-146)
        if (tiled) {
            if (!s->grid_clean) {   // first substep after set_rect / a generic-path substep: wipe everything once
                CU_TRY(cudaMemsetAsync(s->grid, 0, n_alloc * sizeof(float4), s->stream));
                CU_TRY(cudaMemsetAsync(s->gmass, 0, n_alloc * sizeof(float), s->stream));
                CU_TRY(cudaMemsetAsync(s->dirty[s->dirty_cur ^ 1], 0, s->geo.n_tiles, s->stream));
                if (s->pool_blocks) {
                    // block-sparse: every block back on the free list, then the blocks this sort's stencils reach
                    // (k_tile_tables marked them in dirty[dirty_cur] before the pool existed in this state)
                    return fail(FLUID_ERR_STATE, "internal: sparse pool must be reset before the sort");
                }
                s->grid_clean = true;
            } else {
                // only the node blocks the previous or the coming deposits can touch
                k_dirty_list<<<blocks_for(s->geo.n_tiles, 256), 256, 0, s->stream>>>(
                    s->geo, s->dirty[s->dirty_cur], s->dirty[s->dirty_cur ^ 1], s->dirty_list, s->scal + SCAL_N_DIRTY, true);
                k_clear_tiles<<<std::min<unsigned>(blocks_for(static_cast<int64_t>(s->geo.n_tiles) * 32, 128),
                                                  static_cast<unsigned>(s->sm_count * 16)), 128, 0, s->stream>>>(
                    s->geo, s->dirty_list, s->scal + SCAL_N_DIRTY, s->grid, s->gmass, s->tile_base, s->gz, s->d_epoch, true,
                    s->p2p ? 1 : 3,    // peer-halo runs clear the node masses at the end of the substep instead
                    s->dirty[s->dirty_cur]);
                s->launches += 2;
            }
            if (timed) CU_TRY(cudaEventRecord(ev[2], s->stream));
            if (DIM == 2)
                k_mass_tiled2<<<std::min(tb2, s->grid2_mass), T2::THREADS, 0, s->stream>>>(s->geo, q.P, s->src, s->tiles, n_act, s->gmass,
                                                                                          s->grid);
            else if (s->p2p)
                k_mass_tiled<true><<<std::min(tb, s->grid_mass), T3::THREADS, 0, s->stream>>>(s->geo, q.P, s->src, s->tiles, n_act,
                                                                                            s->gmass, s->grid, s->peer, s->tex[s->cur], s->tab, (s->dyn_tiles & 1) ? s->scal + SCAL_TICKET : nullptr, s->fixed8[0]);
            else
                k_mass_tiled<false><<<std::min(tb, s->grid_mass), T3::THREADS, 0, s->stream>>>(s->geo, q.P, s->src, s->tiles, n_act,
                                                                                             s->gmass, s->grid, s->peer, s->tex[s->cur], s->tab, (s->dyn_tiles & 1) ? s->scal + SCAL_TICKET : nullptr, s->fixed8[0]);
            ++s->launches;
            if (timed) CU_TRY(cudaEventRecord(ev[3], s->stream));
        } else {
            if (s->det) {
                ST_TRY(ensure_fixed(s, slab_generic ? 2 : 1));
                CU_TRY(cudaMemsetAsync(s->fx[0], 0, n_alloc * 4 * sizeof(long long), s->stream));
                if (slab_generic) CU_TRY(cudaMemsetAsync(s->fx[1], 0, n_alloc * 4 * sizeof(long long), s->stream));
            } else {
                CU_TRY(cudaMemsetAsync(s->grid, 0, n_alloc * sizeof(float4), s->stream));
            }
            s->grid_is_fixed = s->det;
            s->grid_clean = false;
            if (timed) CU_TRY(cudaEventRecord(ev[2], s->stream));
            const NodeGrid ng{s->grid, s->fx[0]};
            if (s->det) k_p2g1_generic<DIM, true><<<blocks_for(n, 128), 128, 0, s->stream>>>(s->geo, q, s->src, n_dep, ng);
            else k_p2g1_generic<DIM, false><<<blocks_for(n, 128), 128, 0, s->stream>>>(s->geo, q, s->src, n_dep, ng);
            ++s->launches;
            if (timed) CU_TRY(cudaEventRecord(ev[3], s->stream));
        }
    }
    if (phases & 2) {
        Particles q = s->buf[s->cur];
        if (tiled && DIM == 2) {
            k_p2g_tiled2<<<std::min(tb2, s->grid2_p2g), T2::THREADS, 0, s->stream>>>(s->geo, q, s->src, s->tiles, n_act, s->gmass, s->grid,
                                                                                    dbg ? dbg->density : nullptr,
                                                                                    dbg ? dbg->pressure : nullptr);
        } else if (tiled) {
            const unsigned gp = std::min(tb, s->grid_p2g);
            float* dd = dbg ? dbg->density : nullptr;
            float* dp = dbg ? dbg->pressure : nullptr;
#define P2G_LAUNCH(PEER, TMA)                                                                                       \
    k_p2g_tiled<PEER, TMA><<<gp, T3::THREADS, sizeof(P2GSmem), s->stream>>>(s->geo, q, s->src, s->tiles, n_act, s->gmass, \
                                                                           s->grid, dd, dp, s->peer, s->tm_grid, s->tm_mass,     \
                                                                           s->tma_mass ? 1 : 0, s->tex[s->cur], s->tab, (s->dyn_tiles & 2) ? s->scal + SCAL_TICKET + 1 : nullptr, s->fixed8[1])
            if (s->p2p && s->tma) P2G_LAUNCH(true, true);
            else if (s->p2p) P2G_LAUNCH(true, false);
            else if (s->tma) P2G_LAUNCH(false, true);
            else P2G_LAUNCH(false, false);
#undef P2G_LAUNCH
        }
        else {
            const NodeGrid ng{s->grid, s->fx[0]};
            // slab runs: the force deposits go to a buffer of their own, so that the second halo exchange carries
            // this phase's deposits only (the first one already completed the mass / momentum planes)
            const NodeGrid ng_out{s->grid, slab_generic ? s->fx[1] : s->fx[0]};
            float* dd = dbg ? dbg->density : nullptr;
            float* dp = dbg ? dbg->pressure : nullptr;
            if (s->det) k_p2g2_generic<DIM, true><<<blocks_for(n, 128), 128, 0, s->stream>>>(s->geo, q, s->src, n_dep, ng, ng_out, dd, dp);
            else k_p2g2_generic<DIM, false><<<blocks_for(n, 128), 128, 0, s->stream>>>(s->geo, q, s->src, n_dep, ng, ng_out, dd, dp);
        }
        ++s->launches;
        if (timed) CU_TRY(cudaEventRecord(ev[4], s->stream));
    }
    if (phases & 4) {
        Particles q = s->buf[s->cur];
        if (tiled) {
            // g2p also counts the particles for the next substep's neighbour search and writes the new
            // state at the sorted slots of the other buffer (no reorder pass)
            Particles qn = s->buf[s->cur ^ 1];
            SlabBufs sb{};
            // behind the header record; with the neighbours' buffers mapped the records go there directly (NVLink)
            for (int sd = 0; sd < 2; ++sd) {
                float* dst = s->peer.mig[sd] ? s->peer.mig[sd] : s->mig_rec[sd];
                sb.rec[sd] = dst ? dst + MIG_WORDS : nullptr;
            }
            sb.cap = s->mig_cap;
            if (DIM == 2)
                k_g2p_tiled2<<<std::min(tb2, s->grid2_g2p), T2::THREADS, 0, s->stream>>>(s->geo, q, qn, s->src, s->tiles, n_act, s->grid,
                                                                                        d_mouse, sort_tables(s), s->gmass, s->gz, s->d_epoch);
            else {
#define G2P_LAUNCH(TMA)                                                                                                       \
    k_g2p_tiled<true, TMA><<<std::min(tb, s->grid_g2p), T3::THREADS, 0, s->stream>>>(                                          \
        s->geo, q, qn, s->src, s->tiles, n_act, s->grid, d_mouse, sort_tables(s), sb, s->gmass, s->gz, s->d_epoch, s->tm_grid, \
        s->tex[s->cur], (s->dyn_tiles & 4) ? s->scal + SCAL_TICKET + 2 : nullptr, s->fixed8[2])
                if (s->tma) G2P_LAUNCH(true);
                else G2P_LAUNCH(false);
#undef G2P_LAUNCH
            }
            // ignored (and, outside slab runs, dropped) particles sit behind the tiles: carried over
            // and counted here; a slab run ends dropped / migrated particles at this point
            const int* n_end = s->geo.slab_on ? s->tile_base + s->geo.n_tiles + 1 : nullptr;
            k_tail<DIM><<<s->sm_count, 256, 0, s->stream>>>(s->geo, q, qn, s->src, n_dep, n_end, n, sort_tables(s), s->d_epoch);
            s->launches += 2;
            if (s->p2p) ST_TRY(clear_mass_rim(s, true));
            s->cur ^= 1;
            s->counts_pending = true;
        } else {
            const NodeGrid ng{s->grid, s->fx[0]};
            const long long* more = slab_generic ? s->fx[1] : nullptr;
            if (s->det) k_g2p_generic<DIM, true><<<blocks_for(n, 128), 128, 0, s->stream>>>(s->geo, q, s->src, n_dep, ng, more, d_mouse);
            else k_g2p_generic<DIM, false><<<blocks_for(n, 128), 128, 0, s->stream>>>(s->geo, q, s->src, n_dep, ng, more, d_mouse);
            ++s->launches;
            if (slab_generic) {   // hand the particles that left the slab to the neighbours (behind the header records)
                k_pack_migrants_generic<<<blocks_for(n, 256), 256, 0, s->stream>>>(
                    s->geo, q, s->src, n_dep, s->mig_rec[0] ? s->mig_rec[0] + MIG_WORDS : nullptr,
                    s->mig_rec[1] ? s->mig_rec[1] + MIG_WORDS : nullptr, s->mig_cap, s->scal);
                ++s->launches;
            }
            s->sorted_valid = false;   // positions moved; the generic path re-sorts from scratch
        }
        if (timed) {
            CU_TRY(cudaEventRecord(ev[5], s->stream));
            s->last_ev = ev;
            s->timers_recorded = true;
            s->resident_timed = false;
        }
    }
    CU_TRY(cudaGetLastError());
    return FLUID_OK;
}

fluid_status substep(fluid_sim* s, const float* d_mouse, bool timed, const DebugTaps* dbg) {
    return s->dim == 3 ? substep_impl<3>(s, d_mouse, timed, dbg) : substep_impl<2>(s, d_mouse, timed, dbg);
}

// ---- CUDA graphs: a steady-state tiled substep is the same dozen launches every time ---------------------
// (same grids, same pointers up to the two buffer parities; the substep number and the mouse position live
// in device memory).  The graph removes the launch gaps between its small sort / clear kernels.
void graphs_drop(fluid_sim* s) {
    for (int k = 0; k < 4; ++k) {
        if (s->graph[k]) cudaGraphExecDestroy(s->graph[k]);
        s->graph[k] = nullptr;
        s->graph_nodes[k] = 0;
    }
    s->graph_n = -1;
}

bool graph_eligible(const fluid_sim* s) {
    return s->graphs_on && s->tiled && s->rect_set && s->n >= s->graph_min && s->sorted_valid &&
           s->counts_pending && s->grid_clean && !s->profiling && !s->geo.slab_on;
}

// one untimed steady-state substep through the graph of the current parity (captured on first use)
fluid_status substep_graphed(fluid_sim* s, const float* d_mouse) {
    if (s->graph_n != s->n) graphs_drop(s);   // launch geometry depends on the particle count
    s->graph_n = s->n;
    const int key = s->cur * 2 + s->dirty_cur;
    if (!s->graph[key]) {
        const int64_t before = s->launches;
        cudaGraph_t g = nullptr;
        CU_TRY(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        const fluid_status st = substep(s, d_mouse, false, nullptr);   // records the launches, flips the host state
        const cudaError_t ce = cudaStreamEndCapture(s->stream, &g);
        if (st != FLUID_OK || ce != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            (void)cudaGetLastError();
            s->graphs_on = false;                     // fall back to plain launches for good
            s->sorted_valid = s->counts_pending = false;   // the captured substep never ran: re-sort from the positions
            return st != FLUID_OK ? st : fail(FLUID_ERR_CUDA, "CUDA graph capture of the substep failed");
        }
        cudaGraphExec_t ge = nullptr;
        const cudaError_t ci = cudaGraphInstantiate(&ge, g, 0);
        cudaGraphDestroy(g);
        if (ci != cudaSuccess) {
            (void)cudaGetLastError();
            s->graphs_on = false;
            s->sorted_valid = s->counts_pending = false;
            return fail(FLUID_ERR_CUDA, "cudaGraphInstantiate failed");
        }
        s->graph[key] = ge;
        s->graph_nodes[key] = static_cast<int>(s->launches - before);
        CU_TRY(cudaGraphLaunch(ge, s->stream));       // the capture did not execute anything
        return FLUID_OK;
    }
    CU_TRY(cudaGraphLaunch(s->graph[key], s->stream));
    // what substep_impl does to the host state
    s->dirty_cur ^= 1;
    s->cur ^= 1;
    s->sorted_valid = true;
    s->counts_pending = true;
    s->launches += s->graph_nodes[key];
    return FLUID_OK;
}

// The device copy is {x, y, present}: kernels always get the same pointer (a captured CUDA graph keeps working
// whether or not the caller passes a mouse position) and test the flag.
fluid_status upload_mouse(fluid_sim* s, const float* mouse_xy, const float** d_mouse) {
    *d_mouse = s->d_mouse;
    const bool on = mouse_xy != nullptr;
    if (!on && !s->mouse_on) return FLUID_OK;   // already {., ., 0} on the device
    s->h_mouse[0] = on ? mouse_xy[0] : 0.0f;
    s->h_mouse[1] = on ? mouse_xy[1] : 0.0f;
    s->h_mouse[2] = on ? 1.0f : 0.0f;
    // (pageable source: the runtime stages the 12 bytes before the call returns, so h_mouse may be reused at once)
    CU_TRY(cudaMemcpyAsync(s->d_mouse, s->h_mouse, 3 * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    s->mouse_on = on;
    return FLUID_OK;
}

// Pull the class counters of the last sort; compact away tombstones (they sit at the tail).
fluid_status refresh_counts(fluid_sim* s, int64_t counts[4]) {
    counts[0] = counts[1] = counts[2] = 0;
    counts[3] = s->dropped_total;
    if (s->n == 0) return FLUID_OK;
    if (!s->rect_set) {   // before set_rect every key is outside the (empty) rects
        counts[2] = s->n;
        return FLUID_OK;
    }
    ST_TRY(s->dim == 3 ? sort_cold<3>(s) : sort_cold<2>(s));
    int h[4];
    CU_TRY(cudaMemcpyAsync(h, s->class_count, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(cudaStreamSynchronize(s->stream));
    if (h[3] > 0) {
        // dropped particles are the last h[3] slots of the SORTED order; bring storage into that
        // order once and cut them off
        const int n = static_cast<int>(s->n);
        Particles from = s->buf[s->cur], to = s->buf[s->cur ^ 1];
        if (s->dim == 3) k_gather_range<3><<<blocks_for(n, 256), 256, 0, s->stream>>>(from, to, s->src, 0, n);
        else k_gather_range<2><<<blocks_for(n, 256), 256, 0, s->stream>>>(from, to, s->src, 0, n);
        ++s->launches;
        CU_TRY(cudaGetLastError());
        s->cur ^= 1;
        s->sorted_valid = false;
    }
    s->dropped_total += h[3];
    s->n -= h[3];
    counts[0] = h[0];
    counts[1] = h[1];
    counts[2] = h[2];
    counts[3] = s->dropped_total;
    return FLUID_OK;
}

}  // namespace

namespace {
template <typename T>
__device__ __forceinline__ bool nonzero(const T& v);
template <> __device__ __forceinline__ bool nonzero<float>(const float& v) { return v != 0.0f; }
template <> __device__ __forceinline__ bool nonzero<float4>(const float4& v) {
    return v.x != 0.0f || v.y != 0.0f || v.z != 0.0f || v.w != 0.0f;
}
__device__ __forceinline__ void add_to(float& a, const float& b) { a += b; }
__device__ __forceinline__ void add_to(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

// own planes += the neighbour's partial sums; a + b == b + a, so both ranks end with the same bits.
template <typename T>
__global__ void k_accumulate_planes(const __grid_constant__ Geo g, T* __restrict__ own, const T* __restrict__ recv,
                                    int64_t n, int z_first, unsigned char* __restrict__ dirty) {
    const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const T v = recv[e];
    if (!nonzero<T>(v)) return;
    T a = own[e];
    add_to(a, v);
    own[e] = a;
    const int plane = g.size[0] * g.size[1];
    const int z = z_first + static_cast<int>(e / plane);
    const int r = static_cast<int>(e % plane);
    const int x = r % g.size[0], y = r / g.size[0];
    dirty[((z / T3::Z) * g.tdim[1] + y / T3::Y) * g.tdim[0] + x / T3::X] = 1;   // cleared next substep
}

__global__ void k_append_migrants(const __grid_constant__ Geo g, const float* __restrict__ rec, int m, Particles to,
                                  int first, SortTables t, bool count, int* __restrict__ n_lost) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = j < m;
    int cls = -1, bucket = 0;
    const int d = first + j;
    if (valid) {
        const float* r = rec + static_cast<size_t>(j) * MIG_WORDS;
        const float4 p = make_float4(r[0], r[1], r[2], r[15]);
        to.P[d] = p;
        to.V[d] = make_float4(r[3], r[4], r[5], r[16]);
        to.CA[d] = make_float4(r[6], r[7], r[8], r[9]);
        to.CB[d] = make_float4(r[10], r[11], r[12], r[13]);
        to.CC[d] = r[14];
        bucket = bucket_of<3>(g, p, cls);
        // a record that belongs to neither this rank nor the rank it came from (it crossed a whole slab in one
        // substep, or was appended by the caller outside the slab) would vanish at the next substep: count it
        if (n_lost && bucket == migrated_bucket(g)) atomicAdd(n_lost, 1);
    }
    if (count) count_global(t, d, bucket, valid);   // `count` is uniform over the grid
}
}  // namespace

__global__ void k_bump(int* counter) { *counter += 1; }

// block-sparse pool: every tile without storage, blocks 1 .. n-1 on the free list (block 0 = overflow block)
__global__ void k_pool_init(SparsePool sp, int n_tiles, int n_blocks) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tiles) sp.blk[i] = -1;
    if (i < n_blocks - 1) sp.free_list[i] = n_blocks - 1 - i;     // popped from the top: 1, 2, 3, ...
    if (i == 0) {
        sp.scal[0] = n_blocks - 1;
        sp.scal[1] = 0;
    }
}

// fluid_read_grid from either layout (reference order out; nodes without storage are zero)
__global__ void k_export_grid_any(const __grid_constant__ Geo g, const float4* __restrict__ grid, int n_nodes,
                                  float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const int x = i % g.size[0], r = i / g.size[0], y = r % g.size[1], z = r / g.size[1];
    const int gi = node_addr(g, x, y, z);
    float4 nd = gi >= 0 ? grid[gi] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (nd.w > 0.0f) {
        const float inv = 1.0f;
        (void)inv;
        nd.x = __fdiv_rn(nd.x, nd.w) + g.dtg[0];
        nd.y = __fdiv_rn(nd.y, nd.w) + g.dtg[1];
        nd.z = __fdiv_rn(nd.z, nd.w) + g.dtg[2];
    }
    out[i * 4 + 0] = nd.x;
    out[i * 4 + 1] = nd.y;
    out[i * 4 + 2] = nd.z;
    out[i * 4 + 3] = nd.w;
}

// `draw`'s binning (3d:472-481): console_xy = (pos.xy / viewport * console) as ivec2
template <int DIM>
__global__ void k_render_frame(const __grid_constant__ Geo g, Particles q, int n, float vx, float vy, int cols, int rows,
                               int* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = q.P[i];
    int cls = -1;
    (void)bucket_of<DIM>(g, p, cls);
    if (cls != CLS_ACTIVE) return;   // iter_particle yields a_rect blocks only (3d:383-387)
    const int cx = rust_as_i32(__fmul_rn(__fdiv_rn(p.x, vx), static_cast<float>(cols)));
    const int cy = rust_as_i32(__fmul_rn(__fdiv_rn(p.y, vy), static_cast<float>(rows)));
    if (cx < 0 || cy < 0 || cx >= cols || cy >= rows) return;
    atomicAdd(&counts[cy * cols + cx], 1);
}

// =========================================================================================
extern "C" {

int32_t fluid_abi_version(void) { return FLUID_B200_ABI_VERSION; }

const char* fluid_last_error(void) { return g_last_error.c_str(); }

const char* fluid_phase_label(int32_t phase) {
    static const char* labels[FLUID_NUM_PHASES] = {"clear", "p2g 1", "p2g 2", "update", "g2p"};
    return (phase >= 0 && phase < FLUID_NUM_PHASES) ? labels[phase] : "";
}

fluid_status fluid_config_default(int32_t dim, fluid_config* out) {
    if (!out || (dim != 2 && dim != 3)) return fail(FLUID_ERR_INVALID_ARG, "fluid_config_default: dim must be 2 or 3");
    std::memset(out, 0, sizeof(*out));
    out->dim = dim;
    out->dt = dim == 2 ? 0.032f : 0.066f;
    out->iterations = static_cast<int32_t>(1.0 / 0.032);
    out->grid_res = dim == 2 ? 32 : 16;
    out->gravity[1] = 0.3f;
    out->rest_density = dim == 2 ? 4.0f : 1.0f;
    out->dynamic_viscosity = 0.1f;
    out->eos_stiffness = 10.0f;
    out->eos_power = 4.0f;
    out->mouse_radius = 10.0f;
    for (int a = 0; a < 3; ++a) {
        out->clip_min[a] = 0.0f;
        out->clip_max[a] = 64.0f;
    }
    out->boundary_damp_dist = 3.0f;
    out->pressure_clamp = dim == 2 ? -0.0f : -0.1f;
    return FLUID_OK;
}

fluid_status fluid_create(const fluid_config* cfg, int32_t device, fluid_sim** out) {
    if (!cfg || !out) return fail(FLUID_ERR_INVALID_ARG, "fluid_create: null argument");
    if (cfg->dim != 2 && cfg->dim != 3) return fail(FLUID_ERR_INVALID_ARG, "fluid_create: dim must be 2 or 3");
    if (cfg->grid_res <= 0) return fail(FLUID_ERR_INVALID_ARG, "fluid_create: grid_res must be positive");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        (void)cudaGetLastError();
        return fail(FLUID_ERR_NO_DEVICE, "fluid_create: no CUDA device (this engine has no CPU fallback)");
    }
    if (device < 0 || device >= n_dev) return fail(FLUID_ERR_INVALID_ARG, "fluid_create: bad device index");
    CU_TRY(cudaSetDevice(device));
    fluid_sim* s = new (std::nothrow) fluid_sim;
    if (!s) return fail(FLUID_ERR_OUT_OF_MEMORY, "fluid_create: host allocation failed");
    s->cfg = *cfg;
    s->dim = cfg->dim;
    s->device = device;
    cudaError_t ce = cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaMalloc(&s->d_mouse, 4 * sizeof(float));
    if (ce == cudaSuccess) ce = cudaMemset(s->d_mouse, 0, 4 * sizeof(float));
    if (ce == cudaSuccess) ce = cudaMalloc(&s->d_epoch, sizeof(int));
    if (ce == cudaSuccess) ce = cudaMemset(s->d_epoch, 0, sizeof(int));
    if (ce == cudaSuccess) ce = cudaMalloc(&s->class_count, 4 * sizeof(int));
    if (ce == cudaSuccess) ce = cudaMalloc(&s->d_counter, sizeof(int));
    if (ce == cudaSuccess) ce = cudaMalloc(&s->scal, SCAL_COUNT * sizeof(int));
    if (ce == cudaSuccess) ce = cudaMemset(s->scal, 0, SCAL_COUNT * sizeof(int));
    for (int i = 0; i < N_EVENTS && ce == cudaSuccess; ++i) ce = cudaEventCreate(&s->ev[i]);
    if (ce != cudaSuccess) {
        fluid_destroy(s);
        return fail(FLUID_ERR_CUDA, std::string("fluid_create: ") + cudaGetErrorString(ce));
    }
    s->stream = s->own_stream;
    const char* force_generic = std::getenv("FLUID_B200_GENERIC");
    s->tiled = !(force_generic && force_generic[0] == '1');
    if (const char* e = std::getenv("FLUID_B200_GRAPH")) s->graphs_on = e[0] != '0';
    if (const char* e = std::getenv("FLUID_B200_TEX")) s->tex_on = e[0] != '0';
    if (const char* e = std::getenv("FLUID_B200_ORDER")) s->tile_order = (e[0] == 'q') ? ORDER_CLASS_Q : ORDER_CLASS_RR;
    if (const char* e = std::getenv("FLUID_B200_DYN")) s->dyn_tiles = std::atoi(e) & 7;
    if (const char* e = std::getenv("FLUID_B200_FIXED8"))
        for (int k = 0; k < 3 && e[k] >= '0' && e[k] <= '8'; ++k) s->fixed8[k] = e[k] - '0';
    if (const char* e = std::getenv("FLUID_B200_SPARSE_BLOCKS")) {
        if (cfg->dim == 3) s->sparse_blocks = std::max<long long>(std::atoll(e), 0);
    }
    if (const char* e = std::getenv("FLUID_B200_DETERMINISTIC")) {
        s->det = e[0] == '1';
        if (s->det) s->tiled = false;
    }
    {
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        s->coop = coop != 0;
        if (const char* e = std::getenv("FLUID_B200_RESIDENT_MAX")) s->resident_max = std::atoll(e);
    }

    cudaFuncSetAttribute(k_p2g_tiled<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(P2GSmem)));
    cudaFuncSetAttribute(k_p2g_tiled<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(P2GSmem)));
    cudaFuncSetAttribute(k_p2g_tiled<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(P2GSmem)));
    cudaFuncSetAttribute(k_p2g_tiled<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(P2GSmem)));
    {   // persistent grids: one wave of resident CTAs per kernel (148 SMs x occupancy)
        cudaDeviceProp prop{};
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) s->sm_count = prop.multiProcessorCount;
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_mass_tiled<false>, T3::THREADS, 0);
        s->grid_mass = static_cast<unsigned>(s->sm_count * std::max(occ, 1));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_p2g_tiled<false, true>, T3::THREADS, sizeof(P2GSmem));
        s->grid_p2g = static_cast<unsigned>(s->sm_count * std::max(occ, 1));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_g2p_tiled<true, true>, T3::THREADS, 0);
        s->grid_g2p = static_cast<unsigned>(s->sm_count * std::max(occ, 1));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_mass_tiled2, T2::THREADS, 0);
        s->grid2_mass = static_cast<unsigned>(s->sm_count * std::max(occ, 1));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_p2g_tiled2, T2::THREADS, 0);
        s->grid2_p2g = static_cast<unsigned>(s->sm_count * std::max(occ, 1));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_g2p_tiled2, T2::THREADS, 0);
        s->grid2_g2p = static_cast<unsigned>(s->sm_count * std::max(occ, 1));
        // a cooperative launch needs every CTA resident: one particle per thread
        int occ_r = 1;
        if (cfg->dim == 3) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_r, k_substeps_resident<3, true>, 128, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_r, k_substeps_resident<2, true>, 128, 0);
        s->resident_max = std::min<int64_t>(s->resident_max, static_cast<int64_t>(s->sm_count) * std::max(occ_r, 1) * 128);
        (void)cudaGetLastError();
    }
    *out = s;
    return FLUID_OK;
}

// unmap a neighbour's arrays (peer-memory halo)
static void close_peer(fluid_sim* s, int side) {
    for (int k = 0; k < 6; ++k) {
        if (s->peer_base[side][k]) cudaIpcCloseMemHandle(s->peer_base[side][k]);
        s->peer_base[side][k] = nullptr;
    }
    s->peer.grid[side] = nullptr;
    s->peer.gmass[side] = nullptr;
    s->peer.dirty[side][0] = s->peer.dirty[side][1] = nullptr;
    s->peer.mig[side] = nullptr;
    s->peer.flag[side] = nullptr;
    s->p2p = s->peer.grid[0] || s->peer.grid[1];
}

fluid_status fluid_destroy(fluid_sim* s) {
    if (s) {
        close_peer(s, 0);
        close_peer(s, 1);
    }
    if (!s) return FLUID_OK;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    graphs_drop(s);
    free_textures(s);
    for (int b = 0; b < 2; ++b) free_particles(s->buf[b]);
    cudaFree(s->gcell);
    cudaFree(s->rank);
    cudaFree(s->imm_list);
    cudaFree(s->src);
    cudaFree(s->gmass);
    cudaFree(s->tiles);
    cudaFree(s->scal);
    cudaFree(s->count);
    cudaFree(s->cell_off);
    cudaFree(s->tile_total);
    cudaFree(s->tile_base);
    cudaFree(s->block_sums);
    cudaFree(s->dirty[0]);
    cudaFree(s->dirty[1]);
    cudaFree(s->cand);
    cudaFree(s->dirty_list);
    cudaFree(s->imm_cnt);
    cudaFree(s->gz);
    cudaFree(s->tile_info);
    cudaFree(s->tab);
    cudaFree(s->d_status);
    if (s->h_status) cudaFreeHost(s->h_status);
    cudaFree(s->flags);
    cudaFree(s->mig_recv[0]);
    cudaFree(s->mig_recv[1]);
    for (int sd = 0; sd < 2; ++sd) {
        cudaFree(s->mig_rec[sd]);
        cudaFree(s->halo_mass_recv[sd]);
        cudaFree(s->halo_node_recv[sd]);
        cudaFree(s->halo_fx_recv[sd]);
    }
    cudaFree(s->class_count);
    cudaFree(s->geo.sp.blk);
    cudaFree(s->geo.sp.free_list);
    cudaFree(s->geo.sp.scal);
    if (s->h_pool) cudaFreeHost(s->h_pool);
    cudaFree(s->grid);
    cudaFree(s->grid2);
    cudaFree(s->fx[0]);
    cudaFree(s->fx[1]);
    cudaFree(s->d_stamps);
    cudaFree(s->d_mouse);
    cudaFree(s->d_epoch);
    cudaFree(s->d_stage);
    cudaFree(s->d_stage_ids);
    cudaFree(s->d_counter);
    for (int i = 0; i < N_EVENTS; ++i)
        if (s->ev[i]) cudaEventDestroy(s->ev[i]);
    for (cudaEvent_t e : s->pool) cudaEventDestroy(e);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    (void)cudaGetLastError();
    delete s;
    return FLUID_OK;
}

fluid_status fluid_set_stream(fluid_sim* s, void* cuda_stream) {
    if (!s) return fail(FLUID_ERR_INVALID_ARG, "fluid_set_stream: null handle");
    CU_TRY(cudaSetDevice(s->device));
    CU_TRY(cudaStreamSynchronize(s->stream));
    graphs_drop(s);
    s->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : s->own_stream;
    return FLUID_OK;
}

fluid_status fluid_synchronize(fluid_sim* s) {
    if (!s) return fail(FLUID_ERR_INVALID_ARG, "fluid_synchronize: null handle");
    CU_TRY(cudaSetDevice(s->device));
    CU_TRY(cudaStreamSynchronize(s->stream));
    return FLUID_OK;
}

namespace {
// cuTensorMapEncodeTiled through the runtime's driver entry point: no link against libcuda
using EncodeTiled = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
bool make_grid_map(const Geo& g, float4* first_node, CUtensorMap* out) {
    if (const char* e = std::getenv("FLUID_B200_NO_TMA")) {
        if (e[0] == '1') return false;
    }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        return false;
    }
    const cuuint64_t dims[4] = {4, static_cast<cuuint64_t>(g.size[0]), static_cast<cuuint64_t>(g.size[1]),
                                static_cast<cuuint64_t>(g.size[2])};
    const cuuint64_t strides[3] = {16, 16ull * g.size[0], 16ull * g.size[0] * g.size[1]};   // bytes, dims 1..3
    const cuuint32_t box[4] = {4, T3::NX, T3::NY, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return reinterpret_cast<EncodeTiled>(fn)(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, first_node, dims, strides, box, estr,
                                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// node masses: x rows of g.size[0] floats.  TMA wants a 16-byte aligned base and strides: needs size[0] % 4 == 0,
// and the tensor starts one float before node 0 (g.guard - 1 is a multiple of 4 then), so x_tensor = x_node + 1.
bool make_mass_map(const Geo& g, float* gmass_alloc, CUtensorMap* out) {
    if (const char* e = std::getenv("FLUID_B200_NO_TMA")) {
        if (e[0] == '1') return false;
    }
    if (g.size[0] % 4 != 0 || (g.guard - 1) % 4 != 0) return false;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        return false;
    }
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(g.size[0]), static_cast<cuuint64_t>(g.size[1]),
                                static_cast<cuuint64_t>(g.size[2])};
    const cuuint64_t strides[2] = {4ull * g.size[0], 4ull * g.size[0] * g.size[1]};
    const cuuint32_t box[3] = {MBOX_X, T3::NY, T3::NZ};
    const cuuint32_t estr[3] = {1, 1, 1};
    return reinterpret_cast<EncodeTiled>(fn)(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, gmass_alloc + g.guard - 1, dims, strides, box,
                                             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

fluid_status fluid_set_rect(fluid_sim* s, const float* mn, const float* mx) {
    if (!s || !mn || !mx) return fail(FLUID_ERR_INVALID_ARG, "fluid_set_rect: null argument");
    CU_TRY(cudaSetDevice(s->device));
    const int D = s->dim;
    Geo g{};
    g.dim = D;
    const float res = static_cast<float>(s->cfg.grid_res);
    int64_t nodes = 1, tiles = 1;
    const int tshape[3] = {D == 3 ? Tile<3>::X : Tile<2>::X, D == 3 ? Tile<3>::Y : Tile<2>::Y,
                           D == 3 ? Tile<3>::Z : 1};
    for (int a = 0; a < 3; ++a) {
        if (a < D) {
            int kmin = host_block_key(mn[a], res), kmax = host_block_key(mx[a], res);
            if (kmax < kmin) return fail(FLUID_ERR_INVALID_ARG, "fluid_set_rect: max < min");
            g.a_lo[a] = kmin;
            g.a_hi[a] = kmax + 1;
            g.p_lo[a] = kmin - 1;
            g.p_hi[a] = kmax + 2;
            int64_t sz = static_cast<int64_t>(g.p_hi[a] - g.p_lo[a]) * s->cfg.grid_res;
            int64_t org = static_cast<int64_t>(g.p_lo[a]) * s->cfg.grid_res;
            if (sz > (1 << 20) || org < INT_MIN / 2 || org > INT_MAX / 2)
                return fail(FLUID_ERR_INVALID_ARG, "fluid_set_rect: rect too large");
            g.size[a] = static_cast<int>(sz);
            g.org[a] = static_cast<int>(org);
        } else {
            g.size[a] = 1;
            g.a_hi[a] = g.p_hi[a] = 1;
        }
        g.tdim[a] = (g.size[a] + tshape[a] - 1) / tshape[a];
        nodes *= g.size[a];
        tiles *= g.tdim[a];
    }
    // the reference indexes the grid with i32 (3d:170-172); keep the same limit
    if (nodes > INT_MAX / 2 || tiles * 256 > INT_MAX / 2)
        return fail(FLUID_ERR_INVALID_ARG, "fluid_set_rect: grid exceeds 2^30 nodes");
    g.n_tiles = static_cast<int>(tiles);
    g.n_cells_pad = static_cast<int>(tiles * 256);
    g.guard = 1 + g.size[0] + (D == 3 ? g.size[0] * g.size[1] : 0);
    g.slab_on = 0;   // set_rect resets the decomposition: call fluid_slab_set (and the IPC import) again
    close_peer(s, 0);
    close_peer(s, 1);
    g.slab_lo = 0;
    g.slab_hi = 0;
    g.res_f = res;
    g.res_i = s->cfg.grid_res;
    g.res_shift = -1;
    for (int b = 0; b < 30; ++b)
        if (s->cfg.grid_res == (1 << b)) g.res_shift = b;
    g.dt = s->cfg.dt;
    g.rest_density = s->cfg.rest_density;
    g.mu = s->cfg.dynamic_viscosity;
    g.stiffness = s->cfg.eos_stiffness;
    g.power = s->cfg.eos_power;
    g.mouse_r2 = s->cfg.mouse_radius * s->cfg.mouse_radius;
    g.pclamp = s->cfg.pressure_clamp;
    g.inv_rest = 1.0f / s->cfg.rest_density;
    g.power_is_4 = s->cfg.eos_power == 4.0f ? 1 : 0;
    for (int a = 0; a < 3; ++a) {
        g.dtg[a] = s->cfg.dt * s->cfg.gravity[a];
        g.clip_lo[a] = s->cfg.clip_min[a];
        g.clip_hi[a] = s->cfg.clip_max[a];
        g.wall_lo[a] = s->cfg.clip_min[a] + s->cfg.boundary_damp_dist;
        g.wall_hi[a] = s->cfg.clip_max[a] - s->cfg.boundary_damp_dist;
    }

    CU_TRY(cudaStreamSynchronize(s->stream));
    graphs_drop(s);
    cudaFree(s->grid);
    cudaFree(s->grid2);
    cudaFree(s->fx[0]);
    cudaFree(s->fx[1]);
    s->grid2 = nullptr;
    s->fx[0] = s->fx[1] = nullptr;
    s->grid_is_fixed = false;
    cudaFree(s->gmass);
    cudaFree(s->tiles);
    cudaFree(s->count);
    cudaFree(s->cell_off);
    cudaFree(s->tile_total);
    cudaFree(s->tile_base);
    cudaFree(s->block_sums);
    cudaFree(s->dirty[0]);
    cudaFree(s->dirty[1]);
    cudaFree(s->cand);
    cudaFree(s->dirty_list);
    cudaFree(s->imm_cnt);
    cudaFree(s->gz);
    cudaFree(s->tile_info);
    cudaFree(s->tab);
    s->cand = s->dirty_list = s->imm_cnt = s->gz = nullptr;
    s->tile_info = nullptr;
    s->tab = nullptr;
    s->dirty[0] = s->dirty[1] = nullptr;
    s->grid_clean = false;
    s->grid = nullptr;
    s->gmass = nullptr;
    s->tiles = nullptr;
    s->count = s->cell_off = s->tile_total = s->tile_base = s->block_sums = nullptr;
    s->rect_set = false;
    s->sorted_valid = s->counts_pending = false;
    const int64_t n_pt = static_cast<int64_t>(g.n_tiles) + N_PSEUDO;     // tiles + the pseudo tiles
    const int64_t m = n_pt * TILE_CELLS;                                  // buckets
    const int64_t nb = (n_pt + SCAN_CHUNK - 1) / SCAN_CHUNK;
    // node storage: dense in the reference's layout, or (fluid_set_sparse) a pool of 8x8x4 blocks behind blk[tile]
    cudaFree(s->geo.sp.blk);
    cudaFree(s->geo.sp.free_list);
    cudaFree(s->geo.sp.scal);
    s->geo.sp = SparsePool{nullptr, nullptr, nullptr};
    g.sp = SparsePool{nullptr, nullptr, nullptr};
    s->pool_blocks = (D == 3 && s->tiled && s->sparse_blocks > 0) ? std::min<int64_t>(s->sparse_blocks, tiles + 1) : 0;
    if (s->pool_blocks > 0 && s->pool_blocks < 28) s->pool_blocks = 28;   // one tile's 3x3x3 neighbourhood + the overflow block
    s->node_alloc = s->pool_blocks ? s->pool_blocks * TILE_CELLS : nodes + 2 * g.guard;
    CU_TRY(cudaMalloc(&s->grid, s->node_alloc * sizeof(float4)));
    CU_TRY(cudaMalloc(&s->gmass, s->node_alloc * sizeof(float)));
    if (s->pool_blocks) {
        CU_TRY(cudaMalloc(&g.sp.blk, (tiles + 8) * sizeof(int)));
        CU_TRY(cudaMalloc(&g.sp.free_list, s->pool_blocks * sizeof(int)));
        CU_TRY(cudaMalloc(&g.sp.scal, 2 * sizeof(int)));
        if (!s->h_pool) CU_TRY(cudaMallocHost(&s->h_pool, 2 * sizeof(int)));
        s->geo.sp = g.sp;   // (owned by the handle from here on, also if a later allocation fails)
        const int64_t m_init = std::max<int64_t>(tiles, s->pool_blocks);
        k_pool_init<<<blocks_for(m_init, 256), 256, 0, s->stream>>>(g.sp, static_cast<int>(tiles), static_cast<int>(s->pool_blocks));
        CU_TRY(cudaMemsetAsync(s->gmass, 0, s->node_alloc * sizeof(float), s->stream));
        CU_TRY(cudaGetLastError());
    }
    s->tma = D == 3 && !s->pool_blocks && make_grid_map(g, s->grid + g.guard, &s->tm_grid);
    s->tma_mass = D == 3 && s->tma && make_mass_map(g, s->gmass, &s->tm_mass);
    if (!s->tma_mass) std::memset(&s->tm_mass, 0, sizeof(s->tm_mass));
    CU_TRY(cudaMalloc(&s->tiles, (static_cast<int64_t>(g.n_tiles) + N_PSEUDO) * sizeof(int4)));
    CU_TRY(cudaMalloc(&s->cand, (n_pt + 8) * sizeof(int)));
    CU_TRY(cudaMalloc(&s->imm_cnt, (n_pt + 8) * sizeof(int)));
    CU_TRY(cudaMemsetAsync(s->imm_cnt, 0, (n_pt + 8) * sizeof(int), s->stream));
    CU_TRY(cudaMalloc(&s->dirty_list, (n_pt + 8) * sizeof(int)));
    CU_TRY(cudaMalloc(&s->gz, (n_pt + 8) * sizeof(int)));
    CU_TRY(cudaMemsetAsync(s->gz, 0, (n_pt + 8) * sizeof(int), s->stream));
    CU_TRY(cudaMemsetAsync(s->d_epoch, 0, sizeof(int), s->stream));
    CU_TRY(cudaMalloc(&s->tile_info, (n_pt + 8) * sizeof(int2)));
    CU_TRY(cudaMalloc(&s->tab, (n_pt + 8) * TAB_BYTES));
    for (int b = 0; b < 2; ++b) {
        CU_TRY(cudaMalloc(&s->dirty[b], g.n_tiles + 8));
        CU_TRY(cudaMemsetAsync(s->dirty[b], 0, g.n_tiles + 8, s->stream));
    }
    CU_TRY(cudaMalloc(&s->count, (m + 8) * sizeof(int)));
    CU_TRY(cudaMalloc(&s->cell_off, (m + 8) * sizeof(int)));
    CU_TRY(cudaMalloc(&s->tile_total, (n_pt + 8) * sizeof(int)));
    CU_TRY(cudaMalloc(&s->tile_base, (n_pt + 8) * sizeof(int)));
    CU_TRY(cudaMalloc(&s->block_sums, nb * sizeof(int)));
    CU_TRY(cudaMemsetAsync(s->count, 0, (m + 8) * sizeof(int), s->stream));
    CU_TRY(cudaMemsetAsync(s->tile_total, 0, (n_pt + 8) * sizeof(int), s->stream));
    CU_TRY(cudaMemsetAsync(s->tile_base, 0, (n_pt + 8) * sizeof(int), s->stream));
    CU_TRY(cudaMemsetAsync(s->grid, 0, s->node_alloc * sizeof(float4), s->stream));
    s->grid_clean = s->pool_blocks > 0;   // block-sparse: arrays, flags and pool are consistent from the start
    s->grid_nodes = nodes;
    s->n_scan_blocks = nb;
    s->geo = g;
    s->rect_set = true;
    return FLUID_OK;
}

fluid_status fluid_get_rects(const fluid_sim* s, int32_t a_lo[3], int32_t a_hi[3], int32_t p_lo[3],
                             int32_t p_hi[3], int32_t grid_origin[3], int32_t grid_size[3]) {
    if (!s) return fail(FLUID_ERR_INVALID_ARG, "fluid_get_rects: null handle");
    if (!s->rect_set) return fail(FLUID_ERR_STATE, "fluid_get_rects: set_rect has not been called");
    for (int a = 0; a < 3; ++a) {
        if (a_lo) a_lo[a] = s->geo.a_lo[a];
        if (a_hi) a_hi[a] = s->geo.a_hi[a];
        if (p_lo) p_lo[a] = s->geo.p_lo[a];
        if (p_hi) p_hi[a] = s->geo.p_hi[a];
        if (grid_origin) grid_origin[a] = s->geo.org[a];
        if (grid_size) grid_size[a] = s->geo.size[a];
    }
    return FLUID_OK;
}

static fluid_status add_from_device(fluid_sim* s, const float* d_rec, const int* d_ids, int64_t n) {
    ST_TRY(ensure_capacity(s, s->n + n));
    const int base_id = s->next_id;
    if (s->dim == 3)
        k_unpack_records<3><<<blocks_for(n, 256), 256, 0, s->stream>>>(d_rec, d_ids, base_id, static_cast<int>(n),
                                                                      s->buf[s->cur], static_cast<int>(s->n));
    else
        k_unpack_records<2><<<blocks_for(n, 256), 256, 0, s->stream>>>(d_rec, d_ids, base_id, static_cast<int>(n),
                                                                      s->buf[s->cur], static_cast<int>(s->n));
    ++s->launches;
    CU_TRY(cudaGetLastError());
    s->n += n;
    s->next_id += static_cast<int32_t>(n);
    s->sorted_valid = s->counts_pending = false;
    return FLUID_OK;
}

fluid_status fluid_add_particles(fluid_sim* s, const float* records, const int32_t* ids, int64_t n) {
    if (!s || n < 0 || (n > 0 && !records)) return fail(FLUID_ERR_INVALID_ARG, "fluid_add_particles: bad argument");
    if (n == 0) return FLUID_OK;
    CU_TRY(cudaSetDevice(s->device));
    const int64_t floats = n * rec_floats(s->dim);
    ST_TRY(ensure_stage(s, floats, ids ? n : 0));
    CU_TRY(cudaMemcpyAsync(s->d_stage, records, floats * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    if (ids) CU_TRY(cudaMemcpyAsync(s->d_stage_ids, ids, n * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    ST_TRY(add_from_device(s, s->d_stage, ids ? s->d_stage_ids : nullptr, n));
    if (ids) {
        int32_t mx = *std::max_element(ids, ids + n);
        s->next_id = std::max(s->next_id, mx + 1);
    }
    return FLUID_OK;
}

fluid_status fluid_add_particles_device(fluid_sim* s, const float* d_records, const int32_t* d_ids, int64_t n) {
    if (!s || n < 0 || (n > 0 && !d_records)) return fail(FLUID_ERR_INVALID_ARG, "fluid_add_particles_device: bad argument");
    if (n == 0) return FLUID_OK;
    CU_TRY(cudaSetDevice(s->device));
    return add_from_device(s, d_records, d_ids, n);
}

fluid_status fluid_clear_particles(fluid_sim* s) {
    if (!s) return fail(FLUID_ERR_INVALID_ARG, "fluid_clear_particles: null handle");
    s->n = 0;
    s->next_id = 0;
    s->dropped_total = 0;
    s->sorted_valid = s->counts_pending = false;
    if (s->pool_blocks && s->rect_set) {
        // block-sparse: no particle, no block.  Otherwise the blocks of the old positions would stay allocated until
        // the next clear releases them, and re-seeding a scene would need a pool of old + new blocks for a moment.
        CU_TRY(cudaSetDevice(s->device));
        const int64_t m_init = std::max<int64_t>(s->geo.n_tiles, s->pool_blocks);
        k_pool_init<<<blocks_for(m_init, 256), 256, 0, s->stream>>>(s->geo.sp, s->geo.n_tiles, static_cast<int>(s->pool_blocks));
        ++s->launches;
        CU_TRY(cudaMemsetAsync(s->grid, 0, s->node_alloc * sizeof(float4), s->stream));
        CU_TRY(cudaMemsetAsync(s->gmass, 0, s->node_alloc * sizeof(float), s->stream));
        CU_TRY(cudaMemsetAsync(s->dirty[0], 0, s->geo.n_tiles, s->stream));
        CU_TRY(cudaMemsetAsync(s->dirty[1], 0, s->geo.n_tiles, s->stream));
        CU_TRY(cudaGetLastError());
        s->grid_clean = true;
    }
    return FLUID_OK;
}

fluid_status fluid_substeps(fluid_sim* s, int32_t n_substeps, const float* mouse_xy) {
    if (!s || n_substeps < 0) return fail(FLUID_ERR_INVALID_ARG, "fluid_substeps: bad argument");
    CU_TRY(cudaSetDevice(s->device));
    const float* d_mouse = nullptr;
    ST_TRY(upload_mouse(s, mouse_xy, &d_mouse));
    if (resident_eligible(s)) return substeps_resident(s, d_mouse, n_substeps);
    for (int32_t i = 0; i < n_substeps; ++i) {
        const bool timed = i == n_substeps - 1;       // the reference's phase timers show the last substep (3d:112)
        if (!timed && graph_eligible(s)) ST_TRY(substep_graphed(s, d_mouse));
        else ST_TRY(substep(s, d_mouse, timed, nullptr));
    }
    return FLUID_OK;
}

fluid_status fluid_set_sparse(fluid_sim* s, int64_t max_blocks) {
    if (!s || max_blocks < 0) return fail(FLUID_ERR_INVALID_ARG, "fluid_set_sparse: bad argument");
    if (s->dim != 3 && max_blocks > 0) return fail(FLUID_ERR_STATE, "fluid_set_sparse: the block-sparse grid belongs to the tiled 3D path");
    s->sparse_blocks = max_blocks;   // takes effect at the next fluid_set_rect
    return FLUID_OK;
}

fluid_status fluid_memory_stats(fluid_sim* s, int64_t out[6]) {
    if (!s || !out) return fail(FLUID_ERR_INVALID_ARG, "fluid_memory_stats: null argument");
    for (int k = 0; k < 6; ++k) out[k] = 0;
    if (!s->rect_set) return FLUID_OK;
    CU_TRY(cudaSetDevice(s->device));
    const int64_t per_node = sizeof(float4) + (s->dim == 3 ? sizeof(float) : 0);
    out[0] = s->node_alloc * per_node;                                   // node storage allocated
    out[1] = (s->grid_nodes + 2 * s->geo.guard) * per_node;              // what the dense layout takes
    out[2] = s->pool_blocks;
    out[4] = s->grid_nodes;
    if (s->pool_blocks) {
        CU_TRY(cudaMemcpyAsync(s->h_pool, s->geo.sp.scal, 2 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CU_TRY(cudaStreamSynchronize(s->stream));
        out[3] = s->pool_blocks - 1 - s->h_pool[0];                      // blocks in use
        out[5] = s->h_pool[1];                                           // pool ran out at some point
        if (s->h_pool[1]) return fail(FLUID_ERR_OUT_OF_MEMORY, "block-sparse node pool exhausted: results are invalid (raise fluid_set_sparse's block count)");
    }
    return FLUID_OK;
}

fluid_status fluid_set_deterministic(fluid_sim* s, int32_t on) {
    if (!s) return fail(FLUID_ERR_INVALID_ARG, "fluid_set_deterministic: null handle");
    if (s->geo.slab_on) return fail(FLUID_ERR_STATE, "fluid_set_deterministic: not available in slab runs");
    if (s->pool_blocks && on) return fail(FLUID_ERR_STATE, "fluid_set_deterministic: not available with block-sparse node storage");
    CU_TRY(cudaSetDevice(s->device));
    CU_TRY(cudaStreamSynchronize(s->stream));
    s->det = on != 0;
    // the fixed-point sums live in the particle-per-thread kernels; the tiled path keeps float reductions
    const char* force_generic = std::getenv("FLUID_B200_GENERIC");
    s->tiled = !s->det && !(force_generic && force_generic[0] == '1');
    s->sorted_valid = s->counts_pending = false;
    s->grid_clean = false;
    s->grid_is_fixed = false;
    return FLUID_OK;
}

fluid_status fluid_set_resident_max(fluid_sim* s, int64_t max_particles) {
    if (!s || max_particles < 0) return fail(FLUID_ERR_INVALID_ARG, "fluid_set_resident_max: bad argument");
    s->resident_max = max_particles;
    return FLUID_OK;
}

fluid_status fluid_step(fluid_sim* s, const float* mouse_xy) {
    if (!s) return fail(FLUID_ERR_INVALID_ARG, "fluid_step: null handle");
    return fluid_substeps(s, s->cfg.iterations, mouse_xy);   // 3d:111
}

fluid_status fluid_particle_counts(fluid_sim* s, int64_t counts[4]) {
    if (!s || !counts) return fail(FLUID_ERR_INVALID_ARG, "fluid_particle_counts: null argument");
    CU_TRY(cudaSetDevice(s->device));
    return refresh_counts(s, counts);
}

fluid_status fluid_particle_count(fluid_sim* s, int64_t* n_active) {
    if (!s || !n_active) return fail(FLUID_ERR_INVALID_ARG, "fluid_particle_count: null argument");
    int64_t c[4];
    ST_TRY(fluid_particle_counts(s, c));
    *n_active = c[0];
    return FLUID_OK;
}

fluid_status fluid_slot_count(const fluid_sim* s, int64_t* n_slots) {
    if (!s || !n_slots) return fail(FLUID_ERR_INVALID_ARG, "fluid_slot_count: null argument");
    *n_slots = s->n;
    return FLUID_OK;
}

fluid_status fluid_read_particles(fluid_sim* s, float* records, int32_t* ids, int64_t capacity, int64_t* n_written) {
    if (!s || capacity < 0 || (capacity > 0 && !records)) return fail(FLUID_ERR_INVALID_ARG, "fluid_read_particles: bad argument");
    CU_TRY(cudaSetDevice(s->device));
    if (n_written) *n_written = 0;
    if (s->n == 0 || !s->rect_set) return FLUID_OK;
    const int rf = rec_floats(s->dim);
    const int64_t cap = std::min<int64_t>(capacity, s->n);
    ST_TRY(ensure_stage(s, std::max<int64_t>(cap, 1) * rf, std::max<int64_t>(cap, 1)));
    CU_TRY(cudaMemsetAsync(s->d_counter, 0, sizeof(int), s->stream));
    const int n = static_cast<int>(s->n);
    if (s->dim == 3)
        k_pack_active<3><<<blocks_for(n, 256), 256, 0, s->stream>>>(s->geo, s->buf[s->cur], n, s->d_stage,
                                                                   s->d_stage_ids, s->d_counter, static_cast<int>(cap));
    else
        k_pack_active<2><<<blocks_for(n, 256), 256, 0, s->stream>>>(s->geo, s->buf[s->cur], n, s->d_stage,
                                                                   s->d_stage_ids, s->d_counter, static_cast<int>(cap));
    ++s->launches;
    CU_TRY(cudaGetLastError());
    int h = 0;
    CU_TRY(cudaMemcpyAsync(&h, s->d_counter, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(cudaStreamSynchronize(s->stream));
    if (n_written) *n_written = h;
    if (h > capacity) return fail(FLUID_ERR_TOO_SMALL, "fluid_read_particles: capacity smaller than the active particle count");
    if (h > 0) {
        CU_TRY(cudaMemcpyAsync(records, s->d_stage, static_cast<int64_t>(h) * rf * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
        if (ids) CU_TRY(cudaMemcpyAsync(ids, s->d_stage_ids, static_cast<int64_t>(h) * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CU_TRY(cudaStreamSynchronize(s->stream));
    }
    return FLUID_OK;
}

fluid_status fluid_get_dt(const fluid_sim* s, float* dt) {
    if (!s || !dt) return fail(FLUID_ERR_INVALID_ARG, "fluid_get_dt: null argument");
    *dt = s->cfg.dt;
    return FLUID_OK;
}

fluid_status fluid_get_phase_times(fluid_sim* s, double seconds[FLUID_NUM_PHASES], double* sort_seconds) {
    if (!s || !seconds) return fail(FLUID_ERR_INVALID_ARG, "fluid_get_phase_times: null argument");
    for (int i = 0; i < FLUID_NUM_PHASES; ++i) seconds[i] = 0.0;
    if (sort_seconds) *sort_seconds = 0.0;
    if (!s->timers_recorded) return FLUID_OK;
    CU_TRY(cudaSetDevice(s->device));
    if (s->resident_timed) {   // %globaltimer stamps (ns) of the resident kernel's last substep
        unsigned long long t[5];
        CU_TRY(cudaMemcpyAsync(t, s->d_stamps, sizeof(t), cudaMemcpyDeviceToHost, s->stream));
        CU_TRY(cudaStreamSynchronize(s->stream));
        seconds[0] = static_cast<double>(t[1] - t[0]) * 1e-9;   // clear
        seconds[1] = static_cast<double>(t[2] - t[1]) * 1e-9;   // p2g 1
        seconds[2] = static_cast<double>(t[3] - t[2]) * 1e-9;   // p2g 2
        seconds[3] = 0.0;                                       // update: folded into g2p's node read
        seconds[4] = static_cast<double>(t[4] - t[3]) * 1e-9;   // g2p
        return FLUID_OK;
    }
    const cudaEvent_t* ev = s->last_ev;
    CU_TRY(cudaEventSynchronize(ev[N_EVENTS - 1]));
    float ms[N_EVENTS - 1];
    for (int i = 0; i < N_EVENTS - 1; ++i) CU_TRY(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
    if (sort_seconds) *sort_seconds = ms[0] * 1e-3;
    seconds[0] = ms[1] * 1e-3;   // clear
    seconds[1] = ms[2] * 1e-3;   // p2g 1
    seconds[2] = ms[3] * 1e-3;   // p2g 2
    seconds[3] = 0.0;            // update: folded into g2p's node read
    seconds[4] = ms[4] * 1e-3;   // g2p
    return FLUID_OK;
}

fluid_status fluid_profile_enable(fluid_sim* s, int32_t on) {
    if (!s) return fail(FLUID_ERR_INVALID_ARG, "fluid_profile_enable: null handle");
    CU_TRY(cudaSetDevice(s->device));
    if (on) {
        if (s->pool.empty()) {
            s->pool.resize(static_cast<size_t>(PROFILE_POOL) * N_EVENTS, nullptr);
            for (cudaEvent_t& e : s->pool) CU_TRY(cudaEventCreate(&e));
        }
        CU_TRY(cudaStreamSynchronize(s->stream));
        s->pool_used = 0;
        for (double& v : s->prof_sum) v = 0.0;
        s->prof_substeps = 0;
        s->profiling = true;
    } else {
        ST_TRY(profile_drain(s));
        s->profiling = false;
    }
    return FLUID_OK;
}

fluid_status fluid_profile_read(fluid_sim* s, double seconds[6], int64_t* n_substeps) {
    if (!s || !seconds) return fail(FLUID_ERR_INVALID_ARG, "fluid_profile_read: null argument");
    CU_TRY(cudaSetDevice(s->device));
    ST_TRY(profile_drain(s));
    seconds[0] = s->prof_sum[0];   // sort
    seconds[1] = s->prof_sum[1];   // clear
    seconds[2] = s->prof_sum[2];   // p2g 1
    seconds[3] = s->prof_sum[3];   // p2g 2
    seconds[4] = 0.0;              // update (folded into g2p)
    seconds[5] = s->prof_sum[4];   // g2p
    if (n_substeps) *n_substeps = s->prof_substeps;
    return FLUID_OK;
}

fluid_status fluid_debug_substep(fluid_sim* s, const float* mouse_xy, int64_t capacity, int32_t* ids, int32_t* cell,
                                 int32_t* key, float* density, float* pressure, int64_t* n_written) {
    if (!s || capacity < 0) return fail(FLUID_ERR_INVALID_ARG, "fluid_debug_substep: bad argument");
    CU_TRY(cudaSetDevice(s->device));
    if (n_written) *n_written = 0;
    if (!s->rect_set || s->n == 0) return FLUID_OK;
    const int D = s->dim;
    const int64_t n = s->n;
    int *d_ids = nullptr, *d_cell = nullptr, *d_key = nullptr;
    float *d_den = nullptr, *d_prs = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_ids); cudaFree(d_cell); cudaFree(d_key); cudaFree(d_den); cudaFree(d_prs);
    };
    cudaError_t ce = cudaMalloc(&d_ids, n * sizeof(int));
    if (ce == cudaSuccess) ce = cudaMalloc(&d_cell, n * D * sizeof(int));
    if (ce == cudaSuccess) ce = cudaMalloc(&d_key, n * D * sizeof(int));
    if (ce == cudaSuccess) ce = cudaMalloc(&d_den, n * sizeof(float));
    if (ce == cudaSuccess) ce = cudaMalloc(&d_prs, n * sizeof(float));
    if (ce != cudaSuccess) {
        cleanup();
        return fail(FLUID_ERR_OUT_OF_MEMORY, "fluid_debug_substep: scratch allocation failed");
    }
    DebugTaps taps;
    taps.ids = d_ids; taps.cell = d_cell; taps.key = d_key; taps.density = d_den; taps.pressure = d_prs;
    const float* d_mouse = nullptr;
    fluid_status st = upload_mouse(s, mouse_xy, &d_mouse);
    if (st == FLUID_OK) st = substep(s, d_mouse, true, &taps);
    int h_dep = 0;
    if (st == FLUID_OK) {
        ce = cudaMemcpyAsync(&h_dep, s->tile_base + s->geo.n_tiles, sizeof(int), cudaMemcpyDeviceToHost, s->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(s->stream);
        if (ce != cudaSuccess) st = fail(FLUID_ERR_CUDA, cudaGetErrorString(ce));
    }
    if (st == FLUID_OK) {
        if (n_written) *n_written = h_dep;
        if (h_dep > capacity) st = fail(FLUID_ERR_TOO_SMALL, "fluid_debug_substep: capacity too small");
    }
    if (st == FLUID_OK && h_dep > 0) {
        if (ids) cudaMemcpy(ids, d_ids, h_dep * sizeof(int), cudaMemcpyDeviceToHost);
        if (cell) cudaMemcpy(cell, d_cell, static_cast<int64_t>(h_dep) * D * sizeof(int), cudaMemcpyDeviceToHost);
        if (key) cudaMemcpy(key, d_key, static_cast<int64_t>(h_dep) * D * sizeof(int), cudaMemcpyDeviceToHost);
        if (density) cudaMemcpy(density, d_den, h_dep * sizeof(float), cudaMemcpyDeviceToHost);
        if (pressure) cudaMemcpy(pressure, d_prs, h_dep * sizeof(float), cudaMemcpyDeviceToHost);
        ce = cudaGetLastError();
        if (ce != cudaSuccess) st = fail(FLUID_ERR_CUDA, cudaGetErrorString(ce));
    }
    cleanup();
    return st;
}

fluid_status fluid_debug_neighbour_table(fluid_sim* s, int64_t capacity, int32_t* ids, int32_t* cell_index,
                                         int64_t* n_written) {
    if (!s || capacity < 0) return fail(FLUID_ERR_INVALID_ARG, "fluid_debug_neighbour_table: bad argument");
    CU_TRY(cudaSetDevice(s->device));
    if (n_written) *n_written = 0;
    if (!s->rect_set || s->n == 0) return FLUID_OK;
    const int64_t n = s->n;
    ST_TRY(s->dim == 3 ? sort_cold<3>(s) : sort_cold<2>(s));
    int *d_ids = nullptr, *d_ref = nullptr;
    cudaError_t ce = cudaMalloc(&d_ids, n * sizeof(int));
    if (ce == cudaSuccess) ce = cudaMalloc(&d_ref, n * sizeof(int));
    if (ce != cudaSuccess) {
        cudaFree(d_ids); cudaFree(d_ref);
        return fail(FLUID_ERR_OUT_OF_MEMORY, "fluid_debug_neighbour_table: scratch allocation failed");
    }
    const int* n_dep = s->tile_base + s->geo.n_tiles;
    if (s->dim == 3)
        k_debug_keys<3><<<blocks_for(n, 256), 256, 0, s->stream>>>(s->geo, s->buf[s->cur], s->src, n_dep, d_ids, nullptr, nullptr, d_ref);
    else
        k_debug_keys<2><<<blocks_for(n, 256), 256, 0, s->stream>>>(s->geo, s->buf[s->cur], s->src, n_dep, d_ids, nullptr, nullptr, d_ref);
    ++s->launches;
    int h_dep = 0;
    fluid_status st = FLUID_OK;
    ce = cudaMemcpyAsync(&h_dep, n_dep, sizeof(int), cudaMemcpyDeviceToHost, s->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s->stream);
    if (ce != cudaSuccess) st = fail(FLUID_ERR_CUDA, cudaGetErrorString(ce));
    if (st == FLUID_OK) {
        if (n_written) *n_written = h_dep;
        if (h_dep > capacity) st = fail(FLUID_ERR_TOO_SMALL, "fluid_debug_neighbour_table: capacity too small");
    }
    if (st == FLUID_OK && h_dep > 0) {
        if (ids) cudaMemcpy(ids, d_ids, h_dep * sizeof(int), cudaMemcpyDeviceToHost);
        if (cell_index) cudaMemcpy(cell_index, d_ref, h_dep * sizeof(int), cudaMemcpyDeviceToHost);
    }
    cudaFree(d_ids);
    cudaFree(d_ref);
    return st;
}

char fluid_frame_char(int32_t n) {   // 3d:488-497
    static const char ramp[] = " .-=*%$#";
    return ramp[n < 1 ? 0 : (n > 7 ? 7 : n)];
}

fluid_status fluid_render_frame(fluid_sim* s, const float viewport_xy[2], int32_t cols, int32_t rows, int32_t* counts) {
    if (!s || !viewport_xy || !counts || cols <= 0 || rows <= 0 || static_cast<int64_t>(cols) * rows > (1 << 24))
        return fail(FLUID_ERR_INVALID_ARG, "fluid_render_frame: bad argument");
    CU_TRY(cudaSetDevice(s->device));
    const int64_t bins = static_cast<int64_t>(cols) * rows;
    std::memset(counts, 0, bins * sizeof(int32_t));
    if (s->n == 0 || !s->rect_set) return FLUID_OK;
    ST_TRY(ensure_stage(s, 0, bins));
    CU_TRY(cudaMemsetAsync(s->d_stage_ids, 0, bins * sizeof(int), s->stream));
    const int n = static_cast<int>(s->n);
    if (s->dim == 3)
        k_render_frame<3><<<blocks_for(n, 256), 256, 0, s->stream>>>(s->geo, s->buf[s->cur], n, viewport_xy[0], viewport_xy[1], cols, rows, s->d_stage_ids);
    else
        k_render_frame<2><<<blocks_for(n, 256), 256, 0, s->stream>>>(s->geo, s->buf[s->cur], n, viewport_xy[0], viewport_xy[1], cols, rows, s->d_stage_ids);
    ++s->launches;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(counts, s->d_stage_ids, bins * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(cudaStreamSynchronize(s->stream));
    return FLUID_OK;
}

fluid_status fluid_debug_tiles(fluid_sim* s, int64_t capacity_tiles, int32_t* tiles4, int64_t* n_tiles) {
    if (!s || capacity_tiles < 0 || !n_tiles) return fail(FLUID_ERR_INVALID_ARG, "fluid_debug_tiles: bad argument");
    CU_TRY(cudaSetDevice(s->device));
    *n_tiles = 0;
    if (!s->rect_set || !s->sorted_valid) return FLUID_OK;
    int h = 0;
    CU_TRY(cudaMemcpyAsync(&h, s->scal + SCAL_N_ACTIVE, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(cudaStreamSynchronize(s->stream));
    *n_tiles = h;
    if (!tiles4) return FLUID_OK;
    if (h > capacity_tiles) return fail(FLUID_ERR_TOO_SMALL, "fluid_debug_tiles: capacity too small");
    if (h > 0) CU_TRY(cudaMemcpy(tiles4, s->tiles, static_cast<size_t>(h) * sizeof(int4), cudaMemcpyDeviceToHost));
    for (int k = 0; k < h; ++k) tiles4[4 * k + 3] &= ORDER_Q_FLAG - 1;   // W without the order flag
    return FLUID_OK;
}

fluid_status fluid_debug_windows(fluid_sim* s, int64_t capacity, int32_t* window_lane, int64_t* n_written) {
    if (!s || capacity < 0 || !window_lane || !n_written) return fail(FLUID_ERR_INVALID_ARG, "fluid_debug_windows: bad argument");
    CU_TRY(cudaSetDevice(s->device));
    *n_written = 0;
    if (!s->rect_set || !s->sorted_valid || !s->tiled) return FLUID_OK;
    int h = 0;
    CU_TRY(cudaMemcpyAsync(&h, s->tile_base + s->geo.n_tiles, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(cudaStreamSynchronize(s->stream));
    if (h > capacity) return fail(FLUID_ERR_TOO_SMALL, "fluid_debug_windows: capacity too small");
    if (h == 0) return FLUID_OK;
    int* d_out = nullptr;
    CU_TRY(cudaMalloc(&d_out, static_cast<size_t>(h) * sizeof(int)));
    cudaError_t ce = cudaMemsetAsync(d_out, 0xff, static_cast<size_t>(h) * sizeof(int), s->stream);   // -1 = no window claimed the slot
    if (ce == cudaSuccess) {
        k_debug_windows<<<s->sm_count * 4, 128, 0, s->stream>>>(s->geo, s->tiles, s->scal + SCAL_N_ACTIVE, s->tab, d_out);
        ++s->launches;
        ce = cudaMemcpyAsync(window_lane, d_out, static_cast<size_t>(h) * sizeof(int), cudaMemcpyDeviceToHost, s->stream);
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s->stream);
    cudaFree(d_out);
    if (ce != cudaSuccess) return fail(FLUID_ERR_CUDA, cudaGetErrorString(ce));
    *n_written = h;
    return FLUID_OK;
}

fluid_status fluid_read_grid(fluid_sim* s, float* nodes, int64_t capacity_nodes, int64_t* n_nodes) {
    if (!s) return fail(FLUID_ERR_INVALID_ARG, "fluid_read_grid: null handle");
    if (!s->rect_set) return fail(FLUID_ERR_STATE, "fluid_read_grid: set_rect has not been called");
    CU_TRY(cudaSetDevice(s->device));
    if (n_nodes) *n_nodes = s->grid_nodes;
    if (!nodes) return FLUID_OK;
    if (capacity_nodes < s->grid_nodes) return fail(FLUID_ERR_TOO_SMALL, "fluid_read_grid: capacity too small");
    const int D = s->dim;
    if (s->grid_is_fixed && s->fx[0]) {   // deterministic mode: the node sums are fixed-point; give the float view
        const int64_t n_alloc = s->node_alloc;
        k_fixed_to_float<<<blocks_for(n_alloc, 256), 256, 0, s->stream>>>(s->fx[0], s->grid, n_alloc);
        ++s->launches;
    }
    float* d_out = nullptr;
    CU_TRY(cudaMalloc(&d_out, s->grid_nodes * (D + 1) * sizeof(float)));
    const int n = static_cast<int>(s->grid_nodes);
    if (s->pool_blocks) k_export_grid_any<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->geo, s->grid, n, d_out);
    else if (D == 3) k_export_grid<3><<<blocks_for(n, 256), 256, 0, s->stream>>>(s->geo, s->grid, n, d_out);
    else k_export_grid<2><<<blocks_for(n, 256), 256, 0, s->stream>>>(s->geo, s->grid, n, d_out);
    ++s->launches;
    cudaError_t ce = cudaMemcpyAsync(nodes, d_out, s->grid_nodes * (D + 1) * sizeof(float), cudaMemcpyDeviceToHost, s->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s->stream);
    cudaFree(d_out);
    if (ce != cudaSuccess) return fail(FLUID_ERR_CUDA, cudaGetErrorString(ce));
    return FLUID_OK;
}

fluid_status fluid_launch_count(const fluid_sim* s, int64_t* launches) {
    if (!s || !launches) return fail(FLUID_ERR_INVALID_ARG, "fluid_launch_count: null argument");
    *launches = s->launches;
    return FLUID_OK;
}

// ---- z-slab decomposition ----------------------------------------------

fluid_status fluid_reserve(fluid_sim* s, int64_t capacity) {
    if (!s || capacity < 0) return fail(FLUID_ERR_INVALID_ARG, "fluid_reserve: bad argument");
    CU_TRY(cudaSetDevice(s->device));
    return ensure_capacity(s, capacity);
}

fluid_status fluid_slab_set(fluid_sim* s, int32_t z_lo, int32_t z_hi, int32_t has_lower, int32_t has_upper) {
    if (!s) return fail(FLUID_ERR_INVALID_ARG, "fluid_slab_set: null handle");
    if (s->dim != 3 || (!s->tiled && !s->det))
        return fail(FLUID_ERR_STATE, "fluid_slab_set: needs the tiled 3D path (or the deterministic mode)");
    if (!s->rect_set) return fail(FLUID_ERR_STATE, "fluid_slab_set: call set_rect first");
    if (s->pool_blocks) return fail(FLUID_ERR_STATE, "fluid_slab_set: block-sparse node storage is single-GPU (the neighbours index each other's dense planes)");
    CU_TRY(cudaSetDevice(s->device));
    const int lo = z_lo - s->geo.org[2], hi = z_hi - s->geo.org[2];
    if (lo < 0 || hi > s->geo.size[2] || lo >= hi || (lo % T3::Z) != 0 || (has_upper && (hi % T3::Z) != 0))
        return fail(FLUID_ERR_INVALID_ARG, "fluid_slab_set: slab must lie in the grid with faces on tile boundaries (multiples of 4 cells from the grid origin)");
    if ((has_lower && lo < 2) || (has_upper && hi > s->geo.size[2] - 2))
        return fail(FLUID_ERR_INVALID_ARG, "fluid_slab_set: an interface needs a node plane on either side");
    CU_TRY(cudaStreamSynchronize(s->stream));
    graphs_drop(s);
    s->geo.slab_on = 1;
    s->geo.slab_lo = lo;
    s->geo.slab_hi = hi;
    s->has_nb[0] = has_lower != 0;
    s->has_nb[1] = has_upper != 0;
    const int64_t plane2 = 2LL * s->geo.size[0] * s->geo.size[1];
    if (s->mig_cap == 0) s->mig_cap = 1 << 18;
    if (!s->d_status) CU_TRY(cudaMalloc(&s->d_status, 16 * sizeof(int)));
    if (!s->h_status) CU_TRY(cudaMallocHost(&s->h_status, 16 * sizeof(int)));
    if (!s->flags) CU_TRY(cudaMalloc(&s->flags, 8 * sizeof(int)));
    CU_TRY(cudaMemsetAsync(s->flags, 0, 8 * sizeof(int), s->stream));
    s->barrier_epoch = 0;
    for (int sd = 0; sd < 2; ++sd) {
        cudaFree(s->mig_rec[sd]);
        cudaFree(s->halo_mass_recv[sd]);
        cudaFree(s->halo_node_recv[sd]);
        s->mig_rec[sd] = nullptr;
        s->halo_mass_recv[sd] = nullptr;
        s->halo_node_recv[sd] = nullptr;
        if (!s->has_nb[sd]) continue;
        CU_TRY(cudaMalloc(&s->mig_rec[sd], (static_cast<int64_t>(s->mig_cap) + 1) * MIG_WORDS * sizeof(float)));
        cudaFree(s->mig_recv[sd]);
        s->mig_recv[sd] = nullptr;
        CU_TRY(cudaMalloc(&s->mig_recv[sd], (static_cast<int64_t>(s->mig_cap) + 1) * MIG_WORDS * sizeof(float)));
        CU_TRY(cudaMemsetAsync(s->mig_recv[sd], 0, MIG_WORDS * sizeof(float), s->stream));
        CU_TRY(cudaMalloc(&s->halo_mass_recv[sd], plane2 * sizeof(float)));
        CU_TRY(cudaMalloc(&s->halo_node_recv[sd], plane2 * sizeof(float4)));
        cudaFree(s->halo_fx_recv[sd]);
        s->halo_fx_recv[sd] = nullptr;
        if (s->det) CU_TRY(cudaMalloc(&s->halo_fx_recv[sd], plane2 * 4 * sizeof(long long)));
    }
    s->sorted_valid = s->counts_pending = false;
    return FLUID_OK;
}

fluid_status fluid_slab_planes(fluid_sim* s, int32_t side, int32_t kind, void** d_own, void** d_recv, int64_t* n_elems) {
    if (!s || side < 0 || side > 1 || kind < 0 || kind > 1 || !d_own || !d_recv || !n_elems)
        return fail(FLUID_ERR_INVALID_ARG, "fluid_slab_planes: bad argument");
    if (!s->geo.slab_on || !s->has_nb[side]) return fail(FLUID_ERR_STATE, "fluid_slab_planes: no neighbour on that side");
    const int zb = side == 0 ? s->geo.slab_lo : s->geo.slab_hi;      // interface, relative cell
    const int64_t plane = static_cast<int64_t>(s->geo.size[0]) * s->geo.size[1];
    const int64_t first = s->geo.guard + (zb - 1) * plane;          // node planes zb-1 and zb
    *n_elems = 2 * plane;
    if (s->det) {
        // deterministic mode: kind 0 = the fixed-point sums after "p2g 1" (mass + momentum), kind 1 = the force
        // deposits of "p2g 2" (their own buffer); 32 bytes per node either way, so n_elems is scaled to what the
        // caller multiplies it with (4 bytes for kind 0, 16 for kind 1)
        ST_TRY(ensure_fixed(s, 2));
        *d_own = s->fx[kind] + first * 4;
        *d_recv = s->halo_fx_recv[side];
        *n_elems = kind == 0 ? 2 * plane * 8 : 2 * plane * 2;
        return FLUID_OK;
    }
    if (kind == 0) {
        *d_own = s->gmass + first;
        *d_recv = s->halo_mass_recv[side];
    } else {
        *d_own = s->grid + first;
        *d_recv = s->halo_node_recv[side];
    }
    return FLUID_OK;
}

fluid_status fluid_slab_phase(fluid_sim* s, int32_t phase, const float* mouse_xy) {
    if (!s || phase < 0 || phase > 2) return fail(FLUID_ERR_INVALID_ARG, "fluid_slab_phase: bad argument");
    if (!s->geo.slab_on) return fail(FLUID_ERR_STATE, "fluid_slab_phase: fluid_slab_set has not been called");
    CU_TRY(cudaSetDevice(s->device));
    const float* d_mouse = nullptr;
    if (phase == 2) ST_TRY(upload_mouse(s, mouse_xy, &d_mouse));
    if (s->n == 0 && !s->tiled) {   // deterministic mode, empty rank: its planes must still be zero for the exchanges
        if (phase == 0) {
            ST_TRY(ensure_fixed(s, 2));
            CU_TRY(cudaMemsetAsync(s->fx[0], 0, s->node_alloc * 4 * sizeof(long long), s->stream));
            CU_TRY(cudaMemsetAsync(s->fx[1], 0, s->node_alloc * 4 * sizeof(long long), s->stream));
            CU_TRY(cudaMemsetAsync(s->scal, 0, SCAL_COUNT * sizeof(int), s->stream));
        }
        return FLUID_OK;
    }
    if (s->n == 0) {
        // a rank may hold no particles yet still has to take part in the exchanges: its planes must be zero
        if (phase == 0 && !s->grid_clean) {
            const int64_t n_alloc = s->node_alloc;
            CU_TRY(cudaMemsetAsync(s->grid, 0, n_alloc * sizeof(float4), s->stream));
            CU_TRY(cudaMemsetAsync(s->gmass, 0, n_alloc * sizeof(float), s->stream));
            CU_TRY(cudaMemsetAsync(s->dirty[0], 0, s->geo.n_tiles, s->stream));
            CU_TRY(cudaMemsetAsync(s->dirty[1], 0, s->geo.n_tiles, s->stream));
            s->grid_clean = true;
        } else if (phase == 0) {
            CU_TRY(cudaMemsetAsync(s->scal, 0, SCAL_COUNT * sizeof(int), s->stream));
            k_dirty_list<<<blocks_for(s->geo.n_tiles, 256), 256, 0, s->stream>>>(
                s->geo, s->dirty[s->dirty_cur], s->dirty[s->dirty_cur ^ 1], s->dirty_list, s->scal + SCAL_N_DIRTY, true);
            k_clear_tiles<<<std::min<unsigned>(blocks_for(static_cast<int64_t>(s->geo.n_tiles) * 32, 128),
                                              static_cast<unsigned>(s->sm_count * 16)), 128, 0, s->stream>>>(
                s->geo, s->dirty_list, s->scal + SCAL_N_DIRTY, s->grid, s->gmass, s->tile_base, s->gz, s->d_epoch, false,
                s->p2p ? 1 : 3, nullptr);
            k_bump<<<1, 1, 0, s->stream>>>(s->d_epoch);   // no g2p ran: no stamp of an earlier substep may match the next clear
            s->dirty_cur ^= 1;
        } else if (phase == 2 && s->p2p) {
            ST_TRY(clear_mass_rim(s, false));   // the neighbours' "p2g 1" deposits of this substep
        }
        return FLUID_OK;
    }
    return substep_impl<3>(s, d_mouse, false, nullptr, 1 << phase);
}

fluid_status fluid_slab_accumulate(fluid_sim* s, int32_t side, int32_t kind) {
    if (!s || side < 0 || side > 1 || kind < 0 || kind > 1) return fail(FLUID_ERR_INVALID_ARG, "fluid_slab_accumulate: bad argument");
    if (!s->geo.slab_on || !s->has_nb[side]) return fail(FLUID_ERR_STATE, "fluid_slab_accumulate: no neighbour on that side");
    CU_TRY(cudaSetDevice(s->device));
    const int zb = side == 0 ? s->geo.slab_lo : s->geo.slab_hi;
    const int64_t plane = static_cast<int64_t>(s->geo.size[0]) * s->geo.size[1];
    const int64_t first = s->geo.guard + (zb - 1) * plane;
    const int64_t n = 2 * plane;
    if (s->det) {
        k_accumulate_fixed<<<blocks_for(n * 4, 256), 256, 0, s->stream>>>(s->fx[kind] + first * 4, s->halo_fx_recv[side], n * 4);
        ++s->launches;
        CU_TRY(cudaGetLastError());
        return FLUID_OK;
    }
    unsigned char* dirty = s->dirty[s->dirty_cur];
    if (kind == 0)
        k_accumulate_planes<float><<<blocks_for(n, 256), 256, 0, s->stream>>>(s->geo, s->gmass + first, s->halo_mass_recv[side], n, zb - 1, dirty);
    else
        k_accumulate_planes<float4><<<blocks_for(n, 256), 256, 0, s->stream>>>(s->geo, s->grid + first, s->halo_node_recv[side], n, zb - 1, dirty);
    ++s->launches;
    CU_TRY(cudaGetLastError());
    return FLUID_OK;
}

// ---- migration with one synchronisation per substep ---------------------------------------------

namespace {
// headers of the two send buffers (word 0 = number of records behind it) and this rank's counters in one block
__global__ void k_pack_status(const int* __restrict__ scal, const int* __restrict__ tile_tail, float* __restrict__ rec_lo,
                              float* __restrict__ rec_hi, int* __restrict__ status) {
    const int i = threadIdx.x;
    if (i < 8) status[i] = scal[i];
    if (i >= 8 && i < 12) status[i] = tile_tail[i - 8];
    if (i == 12 && rec_lo) reinterpret_cast<int*>(rec_lo)[0] = scal[SCAL_MIG_LO];
    if (i == 13 && rec_hi) reinterpret_cast<int*>(rec_hi)[0] = scal[SCAL_MIG_HI];
}
}  // namespace

fluid_status fluid_slab_migrants_begin(fluid_sim* s, void** d_send_lower, void** d_send_upper) {
    if (!s || !d_send_lower || !d_send_upper) return fail(FLUID_ERR_INVALID_ARG, "fluid_slab_migrants_begin: null argument");
    if (!s->geo.slab_on) return fail(FLUID_ERR_STATE, "fluid_slab_migrants_begin: fluid_slab_set has not been called");
    CU_TRY(cudaSetDevice(s->device));
    k_pack_status<<<1, 32, 0, s->stream>>>(s->scal, s->tile_base + s->geo.n_tiles, s->peer.mig[0] ? s->peer.mig[0] : s->mig_rec[0],
                                           s->peer.mig[1] ? s->peer.mig[1] : s->mig_rec[1], s->d_status);
    ++s->launches;
    CU_TRY(cudaGetLastError());
    *d_send_lower = s->mig_rec[0];
    *d_send_upper = s->mig_rec[1];
    return FLUID_OK;
}

fluid_status fluid_slab_migrants_end(fluid_sim* s, const void* d_recv_lower, const void* d_recv_upper, int64_t n_out[2],
                                     int64_t n_in[2]) {
    if (!s || !n_out || !n_in) return fail(FLUID_ERR_INVALID_ARG, "fluid_slab_migrants_end: null argument");
    if (!s->geo.slab_on) return fail(FLUID_ERR_STATE, "fluid_slab_migrants_end: fluid_slab_set has not been called");
    CU_TRY(cudaSetDevice(s->device));
    int* h = s->h_status;
    h[12] = h[13] = 0;
    CU_TRY(cudaMemcpyAsync(h, s->d_status, 12 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    // with the peer-memory migration the neighbours have written into this rank's own receive buffers
    if (!d_recv_lower && s->peer.mig[0]) d_recv_lower = s->mig_recv[0];
    if (!d_recv_upper && s->peer.mig[1]) d_recv_upper = s->mig_recv[1];
    if (d_recv_lower) CU_TRY(cudaMemcpyAsync(h + 12, d_recv_lower, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    if (d_recv_upper) CU_TRY(cudaMemcpyAsync(h + 13, d_recv_upper, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    h[14] = h[15] = 0;
    if (s->flags) CU_TRY(cudaMemcpyAsync(h + 14, s->flags + 2, 2 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(cudaStreamSynchronize(s->stream));   // the one synchronisation of the substep
    if (s->flags && h[14]) return fail(FLUID_ERR_STATE, "fluid_slab_migrants_end: a peer barrier timed out (a neighbour rank stopped)");
    if (s->flags && h[15])   // (counted by the append of the PREVIOUS substep: the records are gone by now)
        return fail(FLUID_ERR_STATE, "fluid_slab_migrants_end: " + std::to_string(h[15]) +
                    " received particle(s) lay outside this rank's slab and were lost (a particle crossed a whole slab in one "
                    "substep: use thicker slabs)");
    if (h[12] > s->mig_cap || h[13] > s->mig_cap) return fail(FLUID_ERR_TOO_SMALL, "fluid_slab_migrants_end: a neighbour handed over more particles than the migrant buffer holds");
    if (h[SCAL_MIG_OVERFLOW]) return fail(FLUID_ERR_TOO_SMALL, "fluid_slab_migrants_end: more particles left the slab in one substep than the migrant buffer holds");
    n_out[0] = h[SCAL_MIG_LO];
    n_out[1] = h[SCAL_MIG_HI];
    n_in[0] = h[12];
    n_in[1] = h[13];
    if (s->n > 0 && s->counts_pending) {
        // the buffer g2p + k_tail just wrote ends where this substep's sort put the dropped bucket
        s->dropped_total += h[8 + 2] - h[8 + 1];
        s->n = h[8 + 1];
    }
    if (!s->tiled && s->n > 0 && (n_out[0] + n_out[1]) > 0) {
        // particle-per-thread path (deterministic mode): the particles that left the slab are dead entries in
        // storage; sort by the new positions and keep what lies before the "dropped" bucket
        ST_TRY(sort_cold<3>(s));
        int tail[4];
        CU_TRY(cudaMemcpyAsync(tail, s->tile_base + s->geo.n_tiles, sizeof(tail), cudaMemcpyDeviceToHost, s->stream));
        CU_TRY(cudaStreamSynchronize(s->stream));
        const int keep = tail[1];
        if (keep > 0) {
            k_gather_range<3><<<blocks_for(keep, 256), 256, 0, s->stream>>>(s->buf[s->cur], s->buf[s->cur ^ 1], s->src, 0, keep);
            ++s->launches;
            CU_TRY(cudaGetLastError());
            s->cur ^= 1;
        }
        s->dropped_total += tail[2] - tail[1];
        s->n = keep;
        s->sorted_valid = false;
    }
    return FLUID_OK;
}

// ---- peer-memory halo (NVLink P2P through CUDA IPC) ---------------------------------------------

namespace {
// Signal both neighbours (a store into their memory over NVLink) and wait for their signals.  Everything
// this rank's earlier kernels wrote into the neighbours' arrays is visible to them once they see the flag.
__global__ void k_peer_barrier(int* __restrict__ my_flags, int* peer_lo, int* peer_hi, int epoch) {
    if (threadIdx.x != 0) return;
    __threadfence_system();
    if (peer_lo) *reinterpret_cast<volatile int*>(peer_lo) = epoch;
    if (peer_hi) *reinterpret_cast<volatile int*>(peer_hi) = epoch;
    __threadfence_system();
    const long long t0 = clock64();
    volatile int* f = my_flags;
    while ((peer_lo && f[0] < epoch) || (peer_hi && f[1] < epoch)) {
        if (clock64() - t0 > (1LL << 35)) {   // ~17 s: the neighbour is gone; report instead of hanging
            my_flags[2] = 1;
            break;
        }
        __nanosleep(64);
    }
    __threadfence_system();
}
}  // namespace

fluid_status fluid_slab_peer_barrier(fluid_sim* s) {
    if (!s) return fail(FLUID_ERR_INVALID_ARG, "fluid_slab_peer_barrier: null handle");
    if (!s->geo.slab_on || !s->p2p) return fail(FLUID_ERR_STATE, "fluid_slab_peer_barrier: no neighbour is mapped");
    CU_TRY(cudaSetDevice(s->device));
    ++s->barrier_epoch;
    k_peer_barrier<<<1, 32, 0, s->stream>>>(s->flags, s->peer.flag[0], s->peer.flag[1], s->barrier_epoch);
    ++s->launches;
    CU_TRY(cudaGetLastError());
    return FLUID_OK;
}

fluid_status fluid_slab_append_received(fluid_sim* s, int32_t side, int64_t n) {
    if (!s || side < 0 || side > 1 || n < 0) return fail(FLUID_ERR_INVALID_ARG, "fluid_slab_append_received: bad argument");
    if (!s->geo.slab_on || !s->mig_recv[side]) return fail(FLUID_ERR_STATE, "fluid_slab_append_received: no neighbour on that side");
    if (n > s->mig_cap) return fail(FLUID_ERR_TOO_SMALL, "fluid_slab_append_received: more records than the buffer holds");
    return fluid_slab_append(s, s->mig_recv[side] + MIG_WORDS, n);
}

fluid_status fluid_slab_ipc_export(fluid_sim* s, void* handles) {
    if (!s || !handles) return fail(FLUID_ERR_INVALID_ARG, "fluid_slab_ipc_export: null argument");
    if (!s->geo.slab_on) return fail(FLUID_ERR_STATE, "fluid_slab_ipc_export: fluid_slab_set has not been called");
    CU_TRY(cudaSetDevice(s->device));
    static_assert(FLUID_IPC_BYTES == 7 * sizeof(cudaIpcMemHandle_t), "FLUID_IPC_BYTES");
    cudaIpcMemHandle_t* h = static_cast<cudaIpcMemHandle_t*>(handles);
    std::memset(h, 0, FLUID_IPC_BYTES);
    CU_TRY(cudaIpcGetMemHandle(&h[0], s->grid));
    CU_TRY(cudaIpcGetMemHandle(&h[1], s->gmass));
    CU_TRY(cudaIpcGetMemHandle(&h[2], s->dirty[0]));
    CU_TRY(cudaIpcGetMemHandle(&h[3], s->dirty[1]));
    if (s->mig_recv[0]) CU_TRY(cudaIpcGetMemHandle(&h[4], s->mig_recv[0]));   // what the lower neighbour writes
    if (s->mig_recv[1]) CU_TRY(cudaIpcGetMemHandle(&h[5], s->mig_recv[1]));   // what the upper neighbour writes
    CU_TRY(cudaIpcGetMemHandle(&h[6], s->flags));
    CU_TRY(cudaMemsetAsync(s->flags, 0, 8 * sizeof(int), s->stream));
    s->barrier_epoch = 0;
    // The neighbours may deposit into this rank's planes before its own first substep gets to wipe the
    // arrays: wipe now (the caller puts a barrier between the imports and the first substep).
    const int64_t n_alloc = s->node_alloc;
    CU_TRY(cudaMemsetAsync(s->grid, 0, n_alloc * sizeof(float4), s->stream));
    CU_TRY(cudaMemsetAsync(s->gmass, 0, n_alloc * sizeof(float), s->stream));
    CU_TRY(cudaMemsetAsync(s->dirty[0], 0, s->geo.n_tiles, s->stream));
    CU_TRY(cudaMemsetAsync(s->dirty[1], 0, s->geo.n_tiles, s->stream));
    CU_TRY(cudaStreamSynchronize(s->stream));
    s->grid_clean = true;
    return FLUID_OK;
}

fluid_status fluid_slab_ipc_import(fluid_sim* s, int32_t side, const void* handles) {
    if (!s || side < 0 || side > 1) return fail(FLUID_ERR_INVALID_ARG, "fluid_slab_ipc_import: bad argument");
    if (!s->geo.slab_on) return fail(FLUID_ERR_STATE, "fluid_slab_ipc_import: fluid_slab_set has not been called");
    if (s->det && handles)
        return fail(FLUID_ERR_STATE, "fluid_slab_ipc_import: the deterministic mode exchanges its fixed-point planes (no peer-memory halo)");
    CU_TRY(cudaSetDevice(s->device));
    close_peer(s, side);
    if (!handles) {   // back to plane exchanges on this side
        s->p2p = s->peer.grid[0] || s->peer.grid[1];
        return FLUID_OK;
    }
    if (!s->has_nb[side]) return fail(FLUID_ERR_STATE, "fluid_slab_ipc_import: no neighbour on that side");
    const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(handles);
    // of the neighbour's two receive buffers this rank writes the one facing it: the lower neighbour's
    // "from above" (handle 5), the upper neighbour's "from below" (handle 4)
    const int which[6] = {0, 1, 2, 3, side == 0 ? 5 : 4, 6};
    for (int k = 0; k < 6; ++k) {
        cudaError_t e = cudaIpcOpenMemHandle(&s->peer_base[side][k], h[which[k]], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            close_peer(s, side);
            return fail(FLUID_ERR_CUDA, cudaGetErrorString(e));
        }
    }
    s->peer.grid[side] = static_cast<float4*>(s->peer_base[side][0]);
    s->peer.gmass[side] = static_cast<float*>(s->peer_base[side][1]);
    s->peer.dirty[side][0] = static_cast<unsigned char*>(s->peer_base[side][2]);
    s->peer.dirty[side][1] = static_cast<unsigned char*>(s->peer_base[side][3]);
    s->peer.mig[side] = static_cast<float*>(s->peer_base[side][4]);
    // this rank is the neighbour's upper (side 0) or lower (side 1) neighbour: its arrival flag there
    s->peer.flag[side] = static_cast<int*>(s->peer_base[side][5]) + (side == 0 ? 1 : 0);
    s->p2p = true;
    return FLUID_OK;
}

fluid_status fluid_slab_migrants(fluid_sim* s, void** d_lower, int64_t* n_lower, void** d_upper, int64_t* n_upper) {
    if (!s || !d_lower || !n_lower || !d_upper || !n_upper) return fail(FLUID_ERR_INVALID_ARG, "fluid_slab_migrants: null argument");
    if (!s->geo.slab_on) return fail(FLUID_ERR_STATE, "fluid_slab_migrants: fluid_slab_set has not been called");
    CU_TRY(cudaSetDevice(s->device));
    int h_scal[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int h_base[4] = {0, 0, 0, 0};   // tile_base[n_tiles .. n_tiles+3]: ignored | dropped | migrated | total
    CU_TRY(cudaMemcpyAsync(h_scal, s->scal, sizeof(h_scal), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(cudaMemcpyAsync(h_base, s->tile_base + s->geo.n_tiles, sizeof(h_base), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(cudaStreamSynchronize(s->stream));
    if (h_scal[SCAL_MIG_OVERFLOW]) return fail(FLUID_ERR_TOO_SMALL, "fluid_slab_migrants: more particles left the slab in one substep than the migrant buffer holds");
    *d_lower = s->mig_rec[0] ? s->mig_rec[0] + MIG_WORDS : nullptr;
    *d_upper = s->mig_rec[1] ? s->mig_rec[1] + MIG_WORDS : nullptr;
    *n_lower = h_scal[SCAL_MIG_LO];
    *n_upper = h_scal[SCAL_MIG_HI];
    if (s->n > 0 && s->counts_pending) {
        // the buffer g2p + k_tail just wrote ends where this substep's sort put the dropped bucket
        s->dropped_total += h_base[2] - h_base[1];
        s->n = h_base[1];
    }
    return FLUID_OK;
}

fluid_status fluid_slab_append(fluid_sim* s, const void* d_records, int64_t n) {
    if (!s || n < 0 || (n > 0 && !d_records)) return fail(FLUID_ERR_INVALID_ARG, "fluid_slab_append: bad argument");
    if (!s->geo.slab_on) return fail(FLUID_ERR_STATE, "fluid_slab_append: fluid_slab_set has not been called");
    if (n == 0) return FLUID_OK;
    CU_TRY(cudaSetDevice(s->device));
    if (s->n + n > s->cap) return fail(FLUID_ERR_OUT_OF_MEMORY, "fluid_slab_append: particle capacity exhausted (fluid_reserve more head room)");
    // With a live sort state the records join it (bucket + rank by global atomics, like k_tail);
    // without one (this rank was empty, or particles were just added) they are only stored and the
    // next substep sorts from scratch.
    const bool join = s->sorted_valid && s->counts_pending;
    k_append_migrants<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->geo, static_cast<const float*>(d_records),
                                                                static_cast<int>(n), s->buf[s->cur],
                                                                static_cast<int>(s->n), sort_tables(s), join,
                                                                s->flags ? s->flags + 3 : nullptr);
    if (!join) s->sorted_valid = s->counts_pending = false;
    ++s->launches;
    CU_TRY(cudaGetLastError());
    s->n += n;
    return FLUID_OK;
}

}  // extern "C"
