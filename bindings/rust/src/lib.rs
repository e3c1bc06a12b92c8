//! `Simulation` of fluid-rs (src/3d_multi.rs:50-408, src/2d_multi.rs:50-385) over the C ABI of
//! include/fluid_b200.h.  Same method names, argument meaning and panic behaviour as the reference,
//! so `main` / `draw` of the two binaries run unchanged apart from the `use`.
//!
//! NOT compiled in the build image of this repository (no Rust toolchain there); INTEGRATION.md
//! walks through the same code.  Citations `3d:` / `2d:` are the reference's source files.
use glam::{Mat2, Mat3, Vec2, Vec3};
use std::os::raw::{c_char, c_int};
use std::time::Duration;

/// `struct Config` (3d:3-15) plus the pressure clamp literal of 3d:218 / 2d:212.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct FluidConfig {
    pub dim: i32,
    pub dt: f32,
    pub iterations: i32,
    pub grid_res: i32,
    pub gravity: [f32; 3],
    pub rest_density: f32,
    pub dynamic_viscosity: f32,
    pub eos_stiffness: f32,
    pub eos_power: f32,
    pub mouse_radius: f32,
    pub clip_min: [f32; 3],
    pub clip_max: [f32; 3],
    pub boundary_damp_dist: f32,
    pub pressure_clamp: f32,
}

#[repr(C)]
pub struct FluidSim {
    _private: [u8; 0],
}

extern "C" {
    fn fluid_config_default(dim: i32, out: *mut FluidConfig) -> c_int; // Config::default, 3d:17-33 / 2d:17-33
    fn fluid_create(cfg: *const FluidConfig, device: i32, out: *mut *mut FluidSim) -> c_int; // Simulation::new, 3d:64
    fn fluid_destroy(sim: *mut FluidSim) -> c_int;
    fn fluid_set_rect(sim: *mut FluidSim, min: *const f32, max: *const f32) -> c_int; // set_rect, 3d:79
    fn fluid_add_particles(sim: *mut FluidSim, rec: *const f32, ids: *const i32, n: i64) -> c_int; // add_particle, 3d:104
    fn fluid_step(sim: *mut FluidSim, mouse_xy: *const f32) -> c_int; // step, 3d:110
    fn fluid_particle_count(sim: *mut FluidSim, n: *mut i64) -> c_int;
    fn fluid_slot_count(sim: *const FluidSim, n: *mut i64) -> c_int;
    fn fluid_set_deterministic(sim: *mut FluidSim, on: i32) -> c_int; // fixed-order node sums (grid_search, 3d:403-408)
    fn fluid_set_sparse(sim: *mut FluidSim, max_blocks: i64) -> c_int; // block-sparse node storage (3d:52-55, 89-96)
    fn fluid_memory_stats(sim: *mut FluidSim, out: *mut i64) -> c_int;
    fn fluid_read_particles(sim: *mut FluidSim, rec: *mut f32, ids: *mut i32, cap: i64, n: *mut i64) -> c_int; // iter_particle, 3d:383
    fn fluid_get_phase_times(sim: *mut FluidSim, sec: *mut f64, sort: *mut f64) -> c_int; // debug_elapseds, 3d:502
    fn fluid_render_frame(sim: *mut FluidSim, viewport_xy: *const f32, cols: i32, rows: i32, counts: *mut i32) -> c_int; // draw's binning, 3d:469-486
    fn fluid_frame_char(count: i32) -> c_char; // draw's character ramp, 3d:487-497
    fn fluid_last_error() -> *const c_char;
}

fn check(st: c_int) {
    if st != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(fluid_last_error()) }.to_string_lossy().into_owned();
        panic!("fluid_b200: status {st}: {msg}"); // the reference panics (unwrap) in the same places
    }
}

impl FluidConfig {
    pub fn default_for(dim: i32) -> Self {
        let mut c = std::mem::MaybeUninit::<FluidConfig>::uninit();
        check(unsafe { fluid_config_default(dim, c.as_mut_ptr()) });
        unsafe { c.assume_init() }
    }
}

const LABELS: [&str; 5] = ["clear", "p2g 1", "p2g 2", "update", "g2p"]; // 3d:116-132

/// State shared by the 2D and 3D front ends.
struct Handle {
    h: *mut FluidSim,
    rec_floats: usize,
    staging: Vec<f32>,
}

impl Handle {
    fn new(config: &FluidConfig) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { fluid_create(config, 0, &mut h) });
        let d = config.dim as usize;
        Self { h, rec_floats: 2 * d + d * d + 1, staging: Vec::new() }
    }
    fn step(&mut self, mouse: Option<[f32; 2]>) -> Vec<(&'static str, Duration)> {
        let ptr = mouse.as_ref().map_or(std::ptr::null(), |m| m.as_ptr());
        check(unsafe { fluid_step(self.h, ptr) });
        let (mut sec, mut sort) = ([0f64; 5], 0f64);
        check(unsafe { fluid_get_phase_times(self.h, sec.as_mut_ptr(), &mut sort) });
        LABELS.iter().zip(sec).map(|(l, s)| (*l, Duration::from_secs_f64(s))).collect()
    }
    fn read(&mut self) -> usize {
        let mut n = 0i64;
        check(unsafe { fluid_slot_count(self.h, &mut n) });   // upper bound, no device work
        self.staging.resize(n as usize * self.rec_floats, 0.0);
        check(unsafe { fluid_read_particles(self.h, self.staging.as_mut_ptr(), std::ptr::null_mut(), n, &mut n) });
        n as usize
    }
    /// The 80x40 frame `draw` prints (3d:469-500), binned on the device: no particle read-back.
    fn frame(&mut self, viewport: [f32; 2], cols: i32, rows: i32) -> String {
        let mut counts = vec![0i32; (cols * rows) as usize];
        check(unsafe { fluid_render_frame(self.h, viewport.as_ptr(), cols, rows, counts.as_mut_ptr()) });
        let mut out = String::with_capacity(((cols + 1) * rows) as usize);
        for r in 0..rows {
            for c in 0..cols {
                out.push(unsafe { fluid_frame_char(counts[(r * cols + c) as usize]) } as u8 as char);
            }
            out.push('\n');
        }
        out
    }
}

impl Drop for Handle {
    fn drop(&mut self) {
        unsafe { fluid_destroy(self.h) };
    }
}

pub mod d3 {
    use super::*;

    /// `struct Particle` (3d:35-41).
    #[derive(Clone, Copy, Default, Debug)]
    pub struct Particle {
        pub pos: Vec3,
        pub vel: Vec3,
        pub affine_momentum: Mat3,
        pub mass: f32,
    }

    pub struct Simulation {
        inner: Handle,
        pub config: FluidConfig,
        pub debug_elapseds: Vec<(&'static str, Duration)>,
        readback: Vec<Particle>,
    }

    impl Simulation {
        pub fn new(config: FluidConfig) -> Self {
            // 3d:64-77
            Self { inner: Handle::new(&config), config, debug_elapseds: Vec::new(), readback: Vec::new() }
        }
        pub fn set_rect(&mut self, min: Vec3, max: Vec3) {
            // 3d:79-102
            check(unsafe { fluid_set_rect(self.inner.h, min.to_array().as_ptr(), max.to_array().as_ptr()) });
        }
        pub fn add_particle(&mut self, p: Particle) {
            // 3d:104-108; record = pos vel C(column-major) mass
            let mut r = [0f32; 16];
            r[0..3].copy_from_slice(&p.pos.to_array());
            r[3..6].copy_from_slice(&p.vel.to_array());
            r[6..15].copy_from_slice(&p.affine_momentum.to_cols_array());
            r[15] = p.mass;
            check(unsafe { fluid_add_particles(self.inner.h, r.as_ptr(), std::ptr::null(), 1) });
        }
        /// Bulk form of `add_particle` (one host-to-device copy).
        pub fn add_particles(&mut self, records: &[[f32; 16]]) {
            check(unsafe { fluid_add_particles(self.inner.h, records.as_ptr() as *const f32, std::ptr::null(), records.len() as i64) });
        }
        pub fn step(&mut self, mouse_pos: &Option<Vec2>) {
            // 3d:110-134
            self.debug_elapseds = self.inner.step(mouse_pos.map(|m| m.to_array()));
        }
        pub fn iter_particle(&mut self) -> impl Iterator<Item = &Particle> + '_ {
            // 3d:383-387 (a device-to-host copy of every record: for dumps only)
            let n = self.inner.read();
            self.readback.clear();
            for r in self.inner.staging.chunks_exact(16).take(n) {
                self.readback.push(Particle {
                    pos: Vec3::from_slice(&r[0..3]),
                    vel: Vec3::from_slice(&r[3..6]),
                    affine_momentum: Mat3::from_cols_slice(&r[6..15]),
                    mass: r[15],
                });
            }
            self.readback.iter()
        }
        pub fn frame(&mut self, viewport: Vec2, cols: i32, rows: i32) -> String {
            self.inner.frame(viewport.to_array(), cols, rows)
        }
        /// Bit-reproducible node sums (64-bit fixed point), from run to run and across GPU counts.
        pub fn set_deterministic(&mut self, on: bool) {
            check(unsafe { fluid_set_deterministic(self.inner.h, on as i32) });
        }
        /// Block-sparse node storage with a pool of `max_blocks` 8x8x4 blocks (call before `set_rect`).
        pub fn set_sparse(&mut self, max_blocks: i64) {
            check(unsafe { fluid_set_sparse(self.inner.h, max_blocks) });
        }
        /// [bytes allocated, dense equivalent, pool blocks, blocks in use, dense nodes, pool ever exhausted]
        pub fn memory_stats(&mut self) -> [i64; 6] {
            let mut o = [0i64; 6];
            check(unsafe { fluid_memory_stats(self.inner.h, o.as_mut_ptr()) });
            o
        }
    }
}

pub mod d2 {
    use super::*;

    /// `struct Particle` (2d:35-41).  `Mat2` is 16-byte aligned in glam, so records are packed by hand.
    #[derive(Clone, Copy, Default, Debug)]
    pub struct Particle {
        pub pos: Vec2,
        pub vel: Vec2,
        pub affine_momentum: Mat2,
        pub mass: f32,
    }

    pub struct Simulation {
        inner: Handle,
        pub config: FluidConfig,
        pub debug_elapseds: Vec<(&'static str, Duration)>,
        readback: Vec<Particle>,
    }

    impl Simulation {
        pub fn new(config: FluidConfig) -> Self {
            // 2d:64-77
            Self { inner: Handle::new(&config), config, debug_elapseds: Vec::new(), readback: Vec::new() }
        }
        pub fn set_rect(&mut self, min: Vec2, max: Vec2) {
            // 2d:79-102
            check(unsafe { fluid_set_rect(self.inner.h, min.to_array().as_ptr(), max.to_array().as_ptr()) });
        }
        pub fn add_particle(&mut self, p: Particle) {
            // 2d:104-108; record = pos vel C(column-major) mass
            let mut r = [0f32; 9];
            r[0..2].copy_from_slice(&p.pos.to_array());
            r[2..4].copy_from_slice(&p.vel.to_array());
            r[4..8].copy_from_slice(&p.affine_momentum.to_cols_array());
            r[8] = p.mass;
            check(unsafe { fluid_add_particles(self.inner.h, r.as_ptr(), std::ptr::null(), 1) });
        }
        pub fn step(&mut self, mouse_pos: &Option<Vec2>) {
            // 2d:110-134
            self.debug_elapseds = self.inner.step(mouse_pos.map(|m| m.to_array()));
        }
        pub fn iter_particle(&mut self) -> impl Iterator<Item = &Particle> + '_ {
            // 2d:361-365
            let n = self.inner.read();
            self.readback.clear();
            for r in self.inner.staging.chunks_exact(9).take(n) {
                self.readback.push(Particle {
                    pos: Vec2::from_slice(&r[0..2]),
                    vel: Vec2::from_slice(&r[2..4]),
                    affine_momentum: Mat2::from_cols_slice(&r[4..8]),
                    mass: r[8],
                });
            }
            self.readback.iter()
        }
        pub fn frame(&mut self, viewport: Vec2, cols: i32, rows: i32) -> String {
            self.inner.frame(viewport.to_array(), cols, rows)
        }
    }
}
