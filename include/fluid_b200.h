/*
 * fluid_b200.h — C ABI of the B200-native step engine (libfluid_b200.so).
 *
 * This is the drop-in boundary for the particle-state and step entry points of
 * GossiperLoturot/fluid-rs (`struct Simulation`, src/3d_multi.rs:50-134,383-387 and
 * src/2d_multi.rs:50-134,361-365).  The reference has no FFI of its own; the entry points
 * below are what a `extern "C"` block in the Rust binaries would bind (INTEGRATION.md shows
 * that block).  Plain pointers and sizes only — no torch / C++ types cross this boundary.
 *
 * Conventions
 *   - every call returns a fluid_status; nothing unwinds across the boundary (the reference
 *     panics via unwrap(), 3d:138,150,173,...; here invariant violations become codes);
 *   - a handle is single-owner and not thread-safe (the reference takes &mut self everywhere);
 *   - all device work of one handle is issued on one CUDA stream; calls are asynchronous
 *     unless they return data to host memory;
 *   - host particle records are packed f32, the field order of `struct Particle`
 *     (3d:35-41 / 2d:35-41):   3D: pos[3] vel[3] C[9] mass  (16 floats, C column-major,
 *     C[3*col+row] as glam Mat3), 2D: pos[2] vel[2] C[4] mass (9 floats).
 */
#ifndef FLUID_B200_H
#define FLUID_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLUID_B200_ABI_VERSION 1

typedef enum fluid_status {
    FLUID_OK = 0,
    FLUID_ERR_INVALID_ARG = 1,   /* null pointer, bad dim, negative count ...            */
    FLUID_ERR_CUDA = 2,          /* a CUDA runtime call failed; see fluid_last_error()   */
    FLUID_ERR_NO_DEVICE = 3,     /* no CUDA device: there is NO CPU fallback             */
    FLUID_ERR_OUT_OF_MEMORY = 4,
    FLUID_ERR_STATE = 5,         /* call not valid in this state (e.g. slab not set)     */
    FLUID_ERR_TOO_SMALL = 6      /* caller buffer smaller than the data to return        */
} fluid_status;

/* Mirrors `struct Config` (3d:3-15 / 2d:3-15).  Vectors are padded to 3 floats; 2D ignores
 * the z lane.  `pressure_clamp` is the literal lower bound of the Tait pressure, which the
 * reference hard-codes per binary (3d:218 = -0.1, 2d:212 = -0.0). */
typedef struct fluid_config {
    int32_t dim;                 /* 2 or 3                                                */
    float   dt;                  /* 3d:20 0.066 | 2d:20 0.032                             */
    int32_t iterations;          /* 3d:21 (1.0/0.032) as i32 = 31                         */
    int32_t grid_res;            /* 3d:22 16    | 2d:22 32   cells per block edge         */
    float   gravity[3];          /* 3d:23 (0,0.3,0)                                       */
    float   rest_density;        /* 3d:24 1.0   | 2d:24 4.0                               */
    float   dynamic_viscosity;   /* 3d:25 0.1                                             */
    float   eos_stiffness;       /* 3d:26 10.0                                            */
    float   eos_power;           /* 3d:27 4.0                                             */
    float   mouse_radius;        /* 3d:28 10.0                                            */
    float   clip_min[3];         /* 3d:29 boundary_clip.0                                 */
    float   clip_max[3];         /* 3d:29 boundary_clip.1                                 */
    float   boundary_damp_dist;  /* 3d:30 3.0                                             */
    float   pressure_clamp;      /* 3d:218 -0.1 | 2d:212 -0.0                             */
} fluid_config;

typedef struct fluid_sim fluid_sim;   /* opaque; owns all device memory */

/* Labels of the five phase timers, in the order of `debug_elapseds` (3d:112-132). */
#define FLUID_NUM_PHASES 5
/* "clear", "p2g 1", "p2g 2", "update", "g2p" */

/* ---- library ---------------------------------------------------------------------- */
int32_t      fluid_abi_version(void);
/* Message of the last failing call on this thread ("" if none). */
const char*  fluid_last_error(void);
/* `Config::default()` of the 2D / 3D binary (3d:17-33 / 2d:17-33). */
fluid_status fluid_config_default(int32_t dim, fluid_config* out);

/* ---- lifecycle: Simulation::new (3d:64-77) ------------------------------------------ */
fluid_status fluid_create(const fluid_config* cfg, int32_t device, fluid_sim** out);
fluid_status fluid_destroy(fluid_sim* sim);
/* Issue all later work of this handle on `cuda_stream` (a cudaStream_t / CUstream; NULL =
 * the handle's own stream).  Lets a host framework time or order the work with its own
 * events. */
fluid_status fluid_set_stream(fluid_sim* sim, void* cuda_stream);
fluid_status fluid_synchronize(fluid_sim* sim);

/* ---- Simulation::set_rect (3d:79-102) ----------------------------------------------- */
/* a_rect = [key(min), key(max)+1), p_rect = a_rect grown by one block; (re)allocates the
 * dense node grid over p_rect.  May be called before or after particles are added; does not
 * remove particles.  min/max hold `dim` floats. */
fluid_status fluid_set_rect(fluid_sim* sim, const float* min, const float* max);
/* Rect read-back: a_rect/p_rect as block keys (lo[dim], hi[dim]), grid origin cell and
 * grid size in cells (3d:94). */
fluid_status fluid_get_rects(const fluid_sim* sim, int32_t a_lo[3], int32_t a_hi[3],
                             int32_t p_lo[3], int32_t p_hi[3],
                             int32_t grid_origin[3], int32_t grid_size[3]);

/* ---- Simulation::add_particle (3d:104-108), bulk ------------------------------------- */
/* Append n packed host records.  `ids` may be NULL (ids continue from the running count).
 * The reference pushes one particle per call; a bulk array is the same operation n times. */
fluid_status fluid_add_particles(fluid_sim* sim, const float* records, const int32_t* ids,
                                 int64_t n);
/* Same, records already resident in device memory (same packed layout). */
fluid_status fluid_add_particles_device(fluid_sim* sim, const float* d_records,
                                        const int32_t* d_ids, int64_t n);
/* Remove all particles (no reference equivalent; needed to reuse a handle). */
fluid_status fluid_clear_particles(fluid_sim* sim);

/* ---- Simulation::step (3d:110-134) --------------------------------------------------- */
/* Runs config.iterations substeps of clear -> p2g 1 -> p2g 2 -> update -> g2p.
 * mouse_xy: NULL for `None`, else two floats (world x,y) for `Some(Vec2)` (3d:305-310). */
fluid_status fluid_step(fluid_sim* sim, const float* mouse_xy);
/* Same phases, an explicit number of substeps (bench / parity harness). */
fluid_status fluid_substeps(fluid_sim* sim, int32_t n_substeps, const float* mouse_xy);
/* The reference fixes the order of every node sum (grid_search, 3d:403-408); float reductions do not, so two
 * runs agree only to the last bits.  on != 0: deposits are rounded to 2^-34 and summed as 64-bit integers
 * (order-independent): results are bit-for-bit reproducible from run to run.  Uses the particle-per-thread
 * kernels (slower than the tiled path); also selectable with FLUID_B200_DETERMINISTIC=1. */
fluid_status fluid_set_deterministic(fluid_sim* sim, int32_t on);
/* Block-sparse node storage (the reference keeps a hash map of blocks plus a touched list so that cost follows
 * the fluid, not the domain: 3d:52-55, 89-96, 136-146).  max_blocks > 0: from the next fluid_set_rect on, the
 * node arrays are a pool of max_blocks blocks of 8x8x4 nodes (5 KB each) behind a per-tile indirection table;
 * blocks are taken when a tile's 3x3x3 neighbourhood first holds particles and given back when it no longer
 * does.  3D tiled path, single GPU; the TMA tile transfers need the dense layout, so this mode uses the
 * per-node loads and reductions instead.  0 = dense (default).  Also FLUID_B200_SPARSE_BLOCKS=N. */
fluid_status fluid_set_sparse(fluid_sim* sim, int64_t max_blocks);
/* out[0] node storage allocated (bytes), [1] what the dense layout would take, [2] pool blocks (0 = dense),
 * [3] pool blocks in use, [4] dense node count, [5] 1 if the pool ever ran out (then this call also returns
 * FLUID_ERR_OUT_OF_MEMORY: deposits were lost). */
fluid_status fluid_memory_stats(fluid_sim* sim, int64_t out[6]);
/* step()/substeps() of at most this many particles run as one cooperative launch with the particle state in
 * registers (the reference's 4,096-particle default scenes are launch-bound); 0 disables it.  Default 16384,
 * clamped to what the device can keep resident; FLUID_B200_RESIDENT_MAX overrides the default. */
fluid_status fluid_set_resident_max(fluid_sim* sim, int64_t max_particles);

/* ---- Simulation::iter_particle (3d:383-387) ------------------------------------------ */
/* Number of particles `iter_particle` would yield (those stored in a_rect blocks). */
fluid_status fluid_particle_count(fluid_sim* sim, int64_t* n_active);
/* Particle slots in use: an upper bound on what fluid_read_particles can return (every class, tombstones
 * not yet compacted included).  No device work and no side effects, unlike the exact counts above and
 * below, which re-run the neighbour search; size read-back buffers with this. */
fluid_status fluid_slot_count(const fluid_sim* sim, int64_t* n_slots);
/* Counts by class: [0] in a_rect blocks (advanced), [1] in the p_rect halo ring (deposit to
 * the grid but frozen, 3d:149 vs 3d:263), [2] outside p_rect (kept, ignored), [3] dropped so
 * far by migration out of p_rect (3d:356-366). */
fluid_status fluid_particle_counts(fluid_sim* sim, int64_t counts[4]);
/* Copy the a_rect particles to host: records (packed, capacity*floats_per_record) and ids
 * (may be NULL).  Order is unspecified, as in the reference (block order x shuffled Vec
 * order); use ids to match.  *n_written receives the count. */
fluid_status fluid_read_particles(fluid_sim* sim, float* records, int32_t* ids,
                                  int64_t capacity, int64_t* n_written);

/* ---- fields main/draw read: sim.config.dt (3d:541), sim.debug_elapseds (3d:502) ------- */
fluid_status fluid_get_dt(const fluid_sim* sim, float* dt);
/* Seconds spent in the five phases of the LAST substep of the last step call (the reference
 * clears the list every substep, 3d:112).  Blocks until that substep has finished.
 * Neighbour-search work (keys, sort, reorder) is reported separately in *sort_seconds
 * (may be NULL); it has no reference counterpart (the hash-of-Vecs is maintained in g2p). */
fluid_status fluid_get_phase_times(fluid_sim* sim, double seconds[FLUID_NUM_PHASES],
                                   double* sort_seconds);
const char*  fluid_phase_label(int32_t phase);
/* Profile mode: CUDA events around every phase of EVERY substep, summed until read.  on=1
 * (re)starts with zeroed sums, on=0 stops.  seconds[0..5] = sort, clear, p2g 1, p2g 2, update,
 * g2p.  The kernels launched are the same with or without it. */
fluid_status fluid_profile_enable(fluid_sim* sim, int32_t on);
fluid_status fluid_profile_read(fluid_sim* sim, double seconds[6], int64_t* n_substeps);

/* ---- headless frame: the binning of `draw` (3d:461-500) without a terminal ------------------- */
/* counts[row*cols + col] = number of a_rect particles whose
 *   console_xy = (pos.xy / viewport * (cols, rows)) as ivec2     (3d:473, truncating cast)
 * falls on that character cell; particles outside the console are skipped (3d:475-477).
 * The reference's main uses viewport (64,64) and an 80 x 40 console (3d:539-540). */
fluid_status fluid_render_frame(fluid_sim* sim, const float viewport_xy[2], int32_t cols, int32_t rows,
                                int32_t* counts);
/* The character `draw` prints for a bin count (3d:488-497): " .-=*%$#". */
char         fluid_frame_char(int32_t count);

/* ---- parity / debug (no reference equivalent; test harness only) ---------------------- */
/* Run ONE substep and return, per p_rect particle in the engine's sorted order:
 * id, cell (floor(pos), 3d:153), block key (3d:398-401), density and pressure (locals at
 * 3d:198,217).  Any pointer may be NULL.  cell/key hold dim ints per particle. */
fluid_status fluid_debug_substep(fluid_sim* sim, const float* mouse_xy, int64_t capacity,
                                 int32_t* ids, int32_t* cell, int32_t* key,
                                 float* density, float* pressure, int64_t* n_written);
/* Neighbour tables of the current state, built from the current positions: per p_rect
 * particle (sorted order) its id and linear cell index in the reference's layout
 * (x + y*sx + z*sx*sy over the p_rect grid, 3d:169-172).  Sorted order is non-decreasing in
 * the engine's tiled cell order; `cell_index` lets a checker rebuild cellStart/cellEnd. */
fluid_status fluid_debug_neighbour_table(fluid_sim* sim, int64_t capacity, int32_t* ids,
                                         int32_t* cell_index, int64_t* n_written);
/* Active tile list of the current sort (tiled paths): per tile {tile id, first sorted slot,
 * particles N, windows W}; a tile's slots are [first, first + N).  Call right after
 * fluid_debug_neighbour_table. */
fluid_status fluid_debug_tiles(fluid_sim* sim, int64_t capacity_tiles, int32_t* tiles4, int64_t* n_tiles);
/* Per sorted slot (same order as fluid_debug_neighbour_table): window * 32 + lane, i.e. the warp iteration
 * and the lane in which the tile kernels process that particle (computed on the device by the kernels' own
 * window walk).  With the cell indices of fluid_debug_neighbour_table a checker can verify the invariants
 * the shared-memory accumulation relies on: every slot is claimed exactly once, no two particles of one
 * window share an (x,y) cell column, and (3D, default order) no two particles of one quarter warp share a
 * 16-byte bank class (x + 2y) mod 8.  capacity in slots. */
fluid_status fluid_debug_windows(fluid_sim* sim, int64_t capacity, int32_t* window_lane, int64_t* n_written);
/* Node grid after the last substep, reference layout: vel_or_momentum[dim] then mass, per
 * node, x fastest (3d:169-172).  `stage`: 0 = as left by the last substep (velocities after
 * update where mass>0). capacity in nodes. */
fluid_status fluid_read_grid(fluid_sim* sim, float* nodes, int64_t capacity_nodes,
                             int64_t* n_nodes);
/* Number of kernels this handle has launched so far (bench `gpu_launches`). */
fluid_status fluid_launch_count(const fluid_sim* sim, int64_t* launches);

/* ---- multi-GPU: z-slab decomposition (no reference equivalent) --- */
/* One process per GPU.  Every rank creates its handle with the SAME config and set_rect (so all
 * ranks index the same dense node grid) and owns the particles whose cell floor(pos.z) lies in
 * [z_lo, z_hi) (world cells; z_lo - grid origin must be a multiple of 4, the tile depth).  Only
 * particles inside the slab are advanced; one that leaves it is handed to the neighbour.
 *
 * Per substep the caller drives (the host binding fluid-rs_b200/slab.py does exactly this):
 *   fluid_slab_phase(0)  sort + clear + p2g 1        -> exchange + accumulate MASS planes
 *   fluid_slab_phase(1)  p2g 2                       -> exchange + accumulate NODE planes
 *   fluid_slab_phase(2)  update + g2p, migrants packed
 *   fluid_slab_migrants  counts (synchronises)       -> exchange migrant records
 *   fluid_slab_append    received records join this rank's particles
 * The node planes shared with a neighbour across the interface at cell z_b are z_b - 1 and z_b
 * (stencil reach 1, 3d:157-158): two contiguous planes of the dense grid. */
fluid_status fluid_slab_set(fluid_sim* sim, int32_t z_lo, int32_t z_hi, int32_t has_lower, int32_t has_upper);
/* Make room for `capacity` particles now, so that appending migrants never reallocates. */
fluid_status fluid_reserve(fluid_sim* sim, int64_t capacity);
/* Shared planes of interface `side` (0 = at z_lo, 1 = at z_hi).  kind 0: node masses (float per
 * node), kind 1: node records (float4 {momentum.xyz, mass}).  *d_own points INTO the grid (this
 * rank's partial sums, send it as is), *d_recv is a staging buffer of the same size for the
 * neighbour's partial sums; *n_elems counts floats (kind 0) or float4s (kind 1). */
fluid_status fluid_slab_planes(fluid_sim* sim, int32_t side, int32_t kind, void** d_own, void** d_recv,
                               int64_t* n_elems);
fluid_status fluid_slab_phase(fluid_sim* sim, int32_t phase, const float* mouse_xy);
/* own planes += received planes (and mark the touched node blocks for the next clear). */
fluid_status fluid_slab_accumulate(fluid_sim* sim, int32_t side, int32_t kind);
/* After phase 2: device pointers to the packed records (17 words: 16 f32 + id) of the particles that
 * left through the lower / upper face and their counts.  Synchronises the stream. */
/* Migration with ONE synchronisation per substep.  Phase 2 leaves the records of the particles that left
 * through a face behind a one-record header whose first word is their number (int32).
 *   fluid_slab_migrants_begin  writes the headers; returns the two send buffers (header + records, NULL
 *                              where there is no neighbour).  The caller exchanges a fixed number of
 *                              records (+ the header) with each neighbour without knowing the counts.
 *   fluid_slab_migrants_end    the synchronisation: reads this rank's counters and the headers of the two
 *                              received buffers (NULL where there is none); n_out / n_in = records sent /
 *                              received through [lower, upper]; cuts the migrated tail off this rank.
 * If a count exceeds what was exchanged, both sides know it and exchange the rest; then
 * fluid_slab_append(received + one record, n_in). */
fluid_status fluid_slab_migrants_begin(fluid_sim* sim, void** d_send_lower, void** d_send_upper);
fluid_status fluid_slab_migrants_end(fluid_sim* sim, const void* d_recv_lower, const void* d_recv_upper,
                                     int64_t n_out[2], int64_t n_in[2]);
/* Peer-memory halo (NVLink P2P; one process per GPU, arrays mapped through CUDA IPC).  With the
 * neighbours' arrays imported, "p2g 1" and "p2g 2" add every deposit that falls into the two node planes
 * an interface shares into the neighbour's copy as well (red.global.add over NVLink, inside the tile
 * kernels' flush), so both ranks hold complete sums when the kernels end: fluid_slab_planes /
 * fluid_slab_accumulate are not used and the caller only puts a neighbour barrier after phase 0 and after
 * phase 1 (and keeps the migrant exchange after phase 2, which orders the next substep).
 *   fluid_slab_ipc_export  FLUID_IPC_BYTES of handles for this rank's arrays; wipes them (call before the
 *                          neighbours may deposit, i.e. before the barrier that ends the set-up)
 *   fluid_slab_ipc_import  map the arrays of the neighbour on `side` (0 = lower, 1 = upper); NULL unmaps.
 * fluid_set_rect reallocates the arrays: export / import again after it. */
#define FLUID_IPC_BYTES 448
fluid_status fluid_slab_ipc_export(fluid_sim* sim, void* handles);
/* With the neighbours mapped nothing in the substep loop goes through a communication library:
 *   fluid_slab_peer_barrier     one tiny kernel: store this rank's epoch into both neighbours' flags over
 *                               NVLink, wait for theirs (time-out after seconds -> error at migrants_end)
 *   phase 2                     writes the leavers' records straight into the neighbour's receive buffer
 *   fluid_slab_migrants_begin   writes the count headers there too; then a peer barrier
 *   fluid_slab_migrants_end     (NULL, NULL): reads the headers of this rank's own receive buffers
 *   fluid_slab_append_received  joins the n records the neighbour on `side` wrote */
fluid_status fluid_slab_peer_barrier(fluid_sim* sim);
fluid_status fluid_slab_append_received(fluid_sim* sim, int32_t side, int64_t n);
fluid_status fluid_slab_ipc_import(fluid_sim* sim, int32_t side, const void* handles);
fluid_status fluid_slab_migrants(fluid_sim* sim, void** d_lower, int64_t* n_lower, void** d_upper,
                                 int64_t* n_upper);
/* Append n packed 17-word records that are resident in device memory. */
fluid_status fluid_slab_append(fluid_sim* sim, const void* d_records, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* FLUID_B200_H */
