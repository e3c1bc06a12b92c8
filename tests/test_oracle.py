"""Pins for the CPU oracle.  The reference has no tests or golden vectors,
so the oracle is pinned by what the reference's code makes derivable: closed forms,
conservation laws, the default-scene geometry constants, and agreement with a second,
independently written restatement (tests/ref_numpy.py)."""
import numpy as np
import pytest

from ref_numpy import RefSim


def make(orc, scene, n=None, records=None):
    sim = orc.OracleSim(scene.cfg)
    rec = scene.records(0, n) if records is None else records
    sim.add_particles(rec)              # particles first, then set_rect, as main does (3d:525-537)
    sim.set_rect(scene.rect_min, scene.rect_max)
    return sim, rec


# ---- helpers: integer rules ------------------------------------------------------------------

def test_quadratic_weights_closed_form(orc):
    # 3d:390-396 at the cell centre and at the two ends of c in [-0.5, 0.5)
    np.testing.assert_array_equal(orc.quadratic_weights(0.0), np.float32([0.125, 0.75, 0.125]))
    np.testing.assert_array_equal(orc.quadratic_weights(-0.5), np.float32([0.5, 0.5, 0.0]))
    w = orc.quadratic_weights(0.25)
    assert abs(float(w.sum()) - 1.0) < 1e-7


def test_weights_partition_of_unity(orc):
    for c in np.linspace(-0.5, 0.4999, 97, dtype=np.float32):
        assert abs(float(orc.quadratic_weights(float(c)).sum()) - 1.0) < 2e-7


def test_key_from_pos_div_euclid(orc):
    # 3d:398-401: floor-division semantics for negatives, exact at block faces
    pos = np.float32([[-0.5, 0.0, 15.999999], [16.0, -16.0, -16.000002], [31.999998, 63.9, 64.0],
                      [-1e-30, 1e9, -1e9]])
    key, cell = orc.key_from_pos(pos, 16)
    np.testing.assert_array_equal(key, [[-1, 0, 0], [1, -1, -2], [1, 3, 4], [-1, 62500000, -62500000]])
    np.testing.assert_array_equal(cell[:3], [[-1, 0, 15], [16, -16, -17], [31, 63, 64]])


@pytest.mark.parametrize("res", [3, 7, 10, 16, 31, 32, 33])
def test_key_equals_floor_div_of_cell(orc, res):
    # key_from_pos divides in f32 (3d:399); for integer grid_res the rounded quotient never
    # crosses a block face (ulp(res*k)/res > ulp(k)/2), so key == floor(floor(pos) / res) even
    # one ulp either side of every face.  The GPU sort relies on this.
    faces = np.arange(-40, 41, dtype=np.float32) * np.float32(res)
    pts = np.concatenate([faces, np.nextafter(faces, np.float32(-1e9)), np.nextafter(faces, np.float32(1e9)),
                          np.random.default_rng(3).uniform(-1300, 1300, 4000).astype(np.float32)])
    key, cell = orc.key_from_pos(pts[:, None], res)
    np.testing.assert_array_equal(key[:, 0], np.floor_divide(cell[:, 0], res))


def test_saturating_cast_and_nan(orc):
    pos = np.float32([[np.nan, np.inf, -np.inf]])
    key, cell = orc.key_from_pos(pos, 16)
    assert key[0, 0] == 0 and key[0, 1] == 2**31 - 1 and key[0, 2] == -2**31


# ---- geometry constants ------------------------------------------------

def test_default_geometry_3d(orc, scenes):
    sim, _ = make(orc, scenes.default_3d())
    r = sim.rects()
    np.testing.assert_array_equal(r["a_lo"], [0, 0, 0])
    np.testing.assert_array_equal(r["a_hi"], [5, 5, 5])
    np.testing.assert_array_equal(r["p_lo"], [-1, -1, -1])
    np.testing.assert_array_equal(r["p_hi"], [6, 6, 6])
    np.testing.assert_array_equal(r["origin"], [-16, -16, -16])
    np.testing.assert_array_equal(r["size"], [112, 112, 112])
    sim.phase(0)
    sim.phase(1)
    assert sim.touched_count() == 27 * 4096        # sparse_grid pushes per substep


def test_default_geometry_2d(orc, scenes):
    sim, _ = make(orc, scenes.default_2d())
    r = sim.rects()
    np.testing.assert_array_equal(r["a_hi"], [3, 3])
    np.testing.assert_array_equal(r["p_lo"], [-1, -1])
    np.testing.assert_array_equal(r["origin"], [-32, -32])
    np.testing.assert_array_equal(r["size"], [160, 160])
    sim.phase(0)
    sim.phase(1)
    assert sim.touched_count() == 9 * 4096


# ---- closed forms -----------------------------------------------------------------------------

@pytest.mark.parametrize("dim,expect", [(2, 0.352539), (3, 0.209320)])
def test_single_particle_density(orc, scenes, dim, expect):
    sc = scenes.default_2d() if dim == 2 else scenes.default_3d()
    rec = np.zeros((1, scenes.rec_floats(dim)), dtype=np.float32)
    rec[0, :dim] = 20.5
    rec[0, -1] = 1.0
    sim, _ = make(orc, sc, records=rec)
    for ph in (0, 1, 2):
        sim.phase(ph)
    d = sim.read(which=1, debug=True)
    assert abs(float(d["density"][0]) - 0.59375 ** dim) < 1e-6
    assert abs(float(d["density"][0]) - expect) < 1e-6
    # pressure sits on the clamp: (rho/rho0)^4 - 1 < 0
    assert np.float32(d["pressure"][0]) == np.float32(sc.cfg["pressure_clamp"])


@pytest.mark.parametrize("dim", [2, 3])
def test_mass_and_momentum_conservation(orc, scenes, dim):
    sc = scenes.default_2d() if dim == 2 else scenes.default_3d()
    rec = sc.records()
    rng = np.random.default_rng(7)
    rec[:, dim:2 * dim] = rng.normal(0, 0.3, (sc.n, dim)).astype(np.float32)
    rec[:, 2 * dim:2 * dim + dim * dim] = rng.normal(0, 0.05, (sc.n, dim * dim)).astype(np.float32)
    rec[:, -1] = rng.uniform(0.5, 1.5, sc.n).astype(np.float32)
    sim, _ = make(orc, sc, records=rec)
    sim.phase(0)
    sim.phase(1)
    g = sim.read_grid().astype(np.float64)
    m_tot = rec[:, -1].astype(np.float64).sum()
    p_tot = (rec[:, -1:].astype(np.float64) * rec[:, dim:2 * dim]).sum(axis=0)
    assert abs(g[:, dim].sum() - m_tot) / m_tot < 1e-6                  # sum of weights = 1
    np.testing.assert_allclose(g[:, :dim].sum(axis=0), p_tot, rtol=0, atol=2e-3)  # sum w*(x_i-x_p)=0
    sim.phase(2)
    g2 = sim.read_grid().astype(np.float64)
    np.testing.assert_allclose(g2[:, :dim].sum(axis=0), p_tot, rtol=0, atol=5e-3)  # internal forces cancel
    np.testing.assert_array_equal(g2[:, dim], g[:, dim])               # p2g_2 never touches mass


@pytest.mark.parametrize("dim", [2, 3])
def test_particle_count_and_clip_after_steps(orc, scenes, dim):
    sc = scenes.default_2d(512) if dim == 2 else scenes.default_3d(512)
    sim, rec = make(orc, sc)
    sim.substeps(200)
    out, ids = sim.read()
    assert out.shape[0] == 512 and sorted(ids.tolist()) == list(range(512))
    pos = out[:, :dim]
    assert (pos >= 0).all() and (pos <= 64).all()
    assert sim.count(2) == 0


def test_soft_wall_rule(orc, scenes):
    # one particle moving fast toward +y: after g2p, pos + vel == wall (61) up to rounding (3d:335)
    sc = scenes.default_3d()
    rec = np.zeros((1, 16), dtype=np.float32)
    rec[0, :3] = [32.3, 59.6, 32.7]
    rec[0, 3:6] = [0.0, 5.0, 0.0]
    rec[0, -1] = 1.0
    sim, _ = make(orc, sc, records=rec)
    sim.substeps(1)
    out, _ = sim.read()
    assert abs(float(out[0, 1] + out[0, 4]) - 61.0) < 1e-4


def test_migration_keeps_every_particle_once(orc, scenes):
    # a drifting cloud crosses block faces: ids stay unique, count constant (3d:345-380)
    sc = scenes.default_3d(2048)
    rec = sc.records()
    rec[:, 3] = 2.0       # +x drift, two cells per substep... clamped by the grid velocity field
    sim, _ = make(orc, sc, records=rec)
    for _ in range(40):
        sim.substeps(1)
        out, ids = sim.read()
        assert len(ids) == 2048 and len(set(ids.tolist())) == 2048


def test_halo_particles_deposit_but_do_not_move(orc, scenes):
    # a particle in a p_rect halo block (key -1) contributes mass but g2p never advances it and
    # iter_particle does not yield it (3d:149 vs 3d:263, 383-387)
    sc = scenes.default_3d()
    rec = np.zeros((2, 16), dtype=np.float32)
    rec[0, :3] = [-3.5, 20.5, 20.5]      # halo block
    rec[1, :3] = [20.5, 20.5, 20.5]
    rec[:, -1] = 1.0
    sim, _ = make(orc, sc, records=rec)
    sim.substeps(3)
    assert sim.count(0) == 1 and sim.count(1) == 2
    allp = sim.read(which=1, debug=True)
    frozen = allp["records"][allp["ids"] == 0][0]
    np.testing.assert_array_equal(frozen[:6], rec[0, :6])
    g = sim.read_grid()
    assert g[:, 3].sum() == pytest.approx(2.0, abs=1e-6)


def _wide_clip_cfg(sc):
    cfg = dict(sc.cfg)
    cfg["clip_min"] = [-200.0, -200.0, -200.0]
    cfg["clip_max"] = [200.0, 200.0, 200.0]
    cfg["gravity"] = [0.0, 0.0, 0.0]
    return cfg


def test_particle_walking_into_halo_block_freezes(orc, scenes):
    # clip box larger than the rect: a particle that walks from an a_rect block into the halo
    # ring is delivered there (3d:370-380) and is never advanced again (g2p walks a_rect only)
    sc = scenes.default_3d()
    sim = orc.OracleSim(_wide_clip_cfg(sc))
    rec = np.zeros((1, 16), dtype=np.float32)
    rec[0, :3] = [78.0, 20.5, 20.5]      # block 4 (last a_rect block); block 5 is halo
    rec[0, 3] = 4.0
    rec[0, -1] = 1.0
    sim.add_particles(rec)
    sim.set_rect([0, 0, 0], [64, 64, 64])
    sim.substeps(60)
    assert sim.count(0) == 0 and sim.count(1) == 1 and sim.count(2) == 0
    a = sim.read(which=1, debug=True)["records"][0].copy()
    assert 80.0 <= a[0] < 96.0
    sim.substeps(5)
    np.testing.assert_array_equal(sim.read(which=1, debug=True)["records"][0], a)


def test_particle_jumping_past_p_rect_is_dropped(orc, scenes):
    # more than one block per substep: the new key is outside p_rect -> dropped (3d:356-366)
    sc = scenes.default_3d()
    sim = orc.OracleSim(_wide_clip_cfg(sc))
    rec = np.zeros((2, 16), dtype=np.float32)
    rec[0, :3] = [78.0, 20.5, 20.5]
    rec[0, 3] = 300.0                    # 19.8 cells per substep: block 4 -> block 6
    rec[1, :3] = [20.5, 20.5, 20.5]
    rec[:, -1] = 1.0
    sim.add_particles(rec)
    sim.set_rect([0, 0, 0], [64, 64, 64])
    sim.substeps(1)
    assert sim.count(1) == 1 and sim.count(2) == 1
    _, ids = sim.read()
    assert ids.tolist() == [1]


# ---- agreement with the second restatement ------------------------------------------------------

@pytest.mark.parametrize("dim", [2, 3])
def test_oracle_matches_independent_numpy_restatement(orc, scenes, dim):
    sc = scenes.default_2d(48) if dim == 2 else scenes.default_3d(48)
    cfg = dict(sc.cfg)
    rng = np.random.default_rng(11)
    rec = np.zeros((48, scenes.rec_floats(dim)), dtype=np.float32)
    # a dense clump inside ONE block so that block order == input order and nothing migrates
    rec[:, :dim] = (20.0 + rng.uniform(0, 4.0, (48, dim))).astype(np.float32)
    rec[:, dim:2 * dim] = rng.normal(0, 0.2, (48, dim)).astype(np.float32)
    rec[:, 2 * dim:2 * dim + dim * dim] = rng.normal(0, 0.05, (48, dim * dim)).astype(np.float32)
    rec[:, -1] = rng.uniform(0.8, 1.2, 48).astype(np.float32)
    sim = orc.OracleSim(cfg)
    sim.add_particles(rec)
    sim.set_rect(sc.rect_min, sc.rect_max)
    r = sim.rects()
    ref = RefSim(cfg, r["origin"], r["size"])
    for row in rec:
        ref.add(row)
    for step in range(3):
        for ph in range(5):
            sim.phase(ph)
            if ph == 2:
                taps = sim.read(which=1, debug=True)
        ref.substep()
        o = np.argsort(taps["ids"])
        # same operation order in both -> agreement to the last bits; powf differs by <= 1 ulp
        np.testing.assert_allclose(taps["density"][o], np.float32(ref.density), rtol=1e-6)
        np.testing.assert_allclose(taps["pressure"][o], np.float32(ref.pressure), rtol=2e-6, atol=2e-6)
        out, ids = sim.read()
        o = np.argsort(ids)
        np.testing.assert_allclose(out[o], ref.records(), rtol=2e-5, atol=2e-6)


# ---- self-golden (regression guard; NOT reference-derived) ---------------------------------------

def test_oracle_self_golden(orc, scenes):
    from pathlib import Path
    f = Path(__file__).parent / "golden" / "oracle_3d_default_256_s31.npz"
    g = np.load(f)
    sc = scenes.default_3d(256)
    sim, _ = make(orc, sc)
    sim.step()
    out, ids = sim.read()
    o = np.argsort(ids)
    np.testing.assert_array_equal(out[o], g["records"])
