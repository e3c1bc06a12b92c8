"""CPU restatement of the tile order arithmetic of fluid-rs_b200/csrc/sort.cuh (ORDER_CLASS_RR, ORDER_CLASS_Q).

The kernels place a particle with a closed form (k_tile_tables builds per-tile tables, k_build_src turns
(cell, rank) into a slot).  This file restates both the DEFINITION of the order (deal the class-major
sequence round robin into W windows, merge the classes round robin inside a window) and the CLOSED FORM the
kernels use, and checks that they agree and that the invariants the shared-memory accumulation relies on
hold.  It does not run the CUDA code (tests/test_gpu_parity.py::test_windows_hold_distinct_columns does that
on the device); it pins the arithmetic, including the division-free floor the kernels use.
"""
import numpy as np
import pytest

CELLS, CLASSES, COLS_PER_CLASS, Z = 256, 8, 8, 4     # local_cell_3d: cell = z + 4 * (column + 8 * class)


def small_div(x, w):
    """floor(x / w) as sort.cuh::small_div computes it: (float(x) + 0.5f) * (1.0f / float(w)), truncated."""
    inv = np.float32(1.0) / np.float32(w)
    return int(np.float32(np.float32(x) + np.float32(0.5)) * inv)


def windows_of(counts):
    n = int(counts.sum())
    col = counts.reshape(CLASSES * COLS_PER_CLASS, Z).sum(axis=1)
    return max((n + 31) // 32, int(col.max()))


def order_by_definition(counts):
    """slot of every (cell, rank), from the definition of the order."""
    n, w_count = int(counts.sum()), windows_of(counts)
    per, extra = divmod(n, w_count)
    cells = np.repeat(np.arange(CELLS), counts)             # class-major cell order = the sequence q
    ranks = np.concatenate([np.arange(c) for c in counts]) if n else np.zeros(0, int)
    cls = cells >> 5
    slot = {}
    for w in range(w_count):
        members = [q for q in range(n) if q % w_count == w]
        # merge round robin over the classes: k-th member of every class that has one, classes ascending
        by_cls = {b: [q for q in members if cls[q] == b] for b in range(CLASSES)}
        lane = 0
        first = w * per + min(w, extra)
        k = 0
        while lane < len(members):
            for b in range(CLASSES):
                if k < len(by_cls[b]):
                    q = by_cls[b][k]
                    slot[(int(cells[q]), int(ranks[q]))] = first + lane
                    lane += 1
            k += 1
        assert lane == per + (1 if w < extra else 0)
    return slot, w_count


def order_by_closed_form(counts):
    """the same through the kernels' tables and formulas (k_tile_tables / k_build_src)."""
    n, w_count = int(counts.sum()), windows_of(counts)
    per, extra = divmod(n, w_count)
    cell_off = np.concatenate([[0], np.cumsum(counts)])[:-1]
    n_cls = counts.reshape(CLASSES, -1).sum(axis=1)
    s_cls = np.concatenate([[0], np.cumsum(n_cls)])[:-1]
    tab = np.zeros((w_count, CLASSES), dtype=int)           # members of class b in window w
    for b in range(CLASSES):
        sb_mod = s_cls[b] - small_div(s_cls[b], w_count) * w_count
        for w in range(w_count):
            off = w - sb_mod
            if off < 0:
                off += w_count
            tab[w, b] = small_div(n_cls[b] - off + w_count - 1, w_count) if off < n_cls[b] else 0
    assert tab.max(initial=0) <= 8                          # 8 columns per class and W >= the fullest column
    slot = {}
    for cell in range(CELLS):
        for rank in range(counts[cell]):
            q = cell_off[cell] + rank
            b = cell >> 5
            w = q - small_div(q, w_count) * w_count
            s_mod = s_cls[b] - small_div(s_cls[b], w_count) * w_count
            off = w - s_mod
            if off < 0:
                off += w_count
            k = small_div(q - (s_cls[b] + off), w_count)
            pos = int(np.minimum(tab[w], k).sum()) + int((tab[w, :b] > k).sum())
            slot[(cell, rank)] = w * per + min(w, extra) + pos
    return slot, w_count


def random_counts(rng, kind):
    if kind == "poisson1":
        return rng.poisson(0.92, CELLS)
    if kind == "sparse":
        c = np.zeros(CELLS, dtype=int)
        c[rng.integers(0, CELLS, 5)] += rng.integers(1, 4, 5)
        return c
    if kind == "one_column":                                 # W is set by the fullest column
        c = rng.poisson(0.3, CELLS)
        c[40:44] += rng.integers(3, 9, 4)
        return c
    if kind == "one_class":
        c = np.zeros(CELLS, dtype=int)
        c[64:96] = rng.poisson(3.0, 32)
        return c
    if kind == "dense":
        return rng.poisson(6.0, CELLS)
    raise ValueError(kind)


def test_small_div_is_exact():
    for w in range(1, 33):
        x = np.arange(0, 8192)
        got = np.array([small_div(int(v), w) for v in x[:: max(1, 8192 // 997)]])
        assert np.array_equal(got, x[:: max(1, 8192 // 997)] // w)
        for v in (0, 1, w - 1, w, w + 1, 8191 - (8191 % w), 8191):
            assert small_div(v, w) == v // w


def test_small_div_tolerates_an_approximate_reciprocal():
    """k_build_src takes 1/W from MUFU.RCP (1 ulp off at most): the floor stays exact with a reciprocal that is
    off by up to 2 ulp either way, for every W <= 32 and every x < 8192."""
    for w in range(1, 33):
        inv = np.float32(1.0) / np.float32(w)
        for d in (-2, -1, 1, 2):
            iv = inv
            for _ in range(abs(d)):
                iv = np.nextafter(iv, np.float32(np.inf if d > 0 else -np.inf))
            x = np.arange(8192, dtype=np.float32)
            got = np.trunc((x + np.float32(0.5)) * iv).astype(np.int64)
            assert np.array_equal(got, np.arange(8192) // w), (w, d)


@pytest.mark.parametrize("kind", ["poisson1", "sparse", "one_column", "one_class", "dense"])
def test_closed_form_matches_definition(kind):
    rng = np.random.default_rng(hash(kind) % 1000)
    for _ in range(6):
        counts = random_counts(rng, kind).astype(int)
        if counts.sum() == 0 or counts.sum() >= 8192:
            continue
        want, w1 = order_by_definition(counts)
        got, w2 = order_by_closed_form(counts)
        assert w1 == w2 and got == want
        n = int(counts.sum())
        assert sorted(got.values()) == list(range(n))       # a permutation of the tile's slots
        # no two particles of one window share an (x,y) column; a window holds at most 32
        per, extra = divmod(n, w1)
        window_of_slot = np.empty(n, dtype=int)
        first = 0
        for w in range(w1):
            ln = per + (1 if w < extra else 0)
            assert ln <= 32
            window_of_slot[first:first + ln] = w
            first += ln
        seen = set()
        for (cell, _), s in got.items():
            key = (int(window_of_slot[s]), cell >> 2)        # column = cell / 4
            assert key not in seen
            seen.add(key)


def test_quarter_warps_see_distinct_bank_classes_when_classes_are_balanced():
    """With every class holding 3 or 4 particles of a window, the 8 lanes of a quarter warp fall into 8
    different classes (16-byte bank groups of the float4 node tile)."""
    counts = np.zeros(CELLS, dtype=int)
    counts[::1] = 1                                          # one particle per cell: 32 per class, 4 per column
    slot, w_count = order_by_closed_form(counts)
    assert w_count == 8
    cls_of_slot = np.empty(256, dtype=int)
    for (cell, _), s in slot.items():
        cls_of_slot[s] = cell >> 5
    for w in range(8):
        lanes = cls_of_slot[32 * w:32 * w + 32]
        for qd in range(4):
            assert len(set(lanes[8 * qd:8 * qd + 8])) == 8


# ---- ORDER_CLASS_Q: rounds on quarter-warp boundaries (the 3D tiled path's default) ---------------------------

def q_windows_of(counts):
    """W of ORDER_CLASS_Q, or None where k_tile_tables falls back to the plain round robin (W > 32)."""
    col = counts.reshape(CLASSES * COLS_PER_CLASS, Z).sum(axis=1)
    n_cls = counts.reshape(CLASSES, -1).sum(axis=1)
    w = max(int(col.max()), (int(n_cls.max()) + 3) // 4)
    return w if w <= 32 else None


def q_order_by_definition(counts):
    """{(cell, rank): (slot, window, lane)} from the definition: class b's (column, z) sequence q_b is dealt from
    window 0 (window = q_b mod W, round = q_b div W); round k of a window is quarter warp k, classes ascending;
    slots run window after window, lane after lane."""
    w_count = q_windows_of(counts)
    n_cls = counts.reshape(CLASSES, -1).sum(axis=1)
    place = {}
    for b in range(CLASSES):
        qb = 0
        for cell in range(32 * b, 32 * b + 32):
            for rank in range(counts[cell]):
                place[(cell, rank)] = (qb % w_count, qb // w_count, b)
                qb += 1
    out, slot = {}, 0
    for w in range(w_count):
        for k in range(4):
            present = [b for b in range(CLASSES) if n_cls[b] > w + k * w_count]
            assert len(present) <= 8
            for j, b in enumerate(present):
                key = [kk for kk, v in place.items() if v == (w, k, b)]
                assert len(key) == 1
                out[key[0]] = (slot, w, 8 * k + j)
                slot += 1
    assert slot == int(counts.sum())
    return out, w_count


def q_order_by_closed_form(counts):
    """the same slots through the table row of k_tile_tables<ORDER_CLASS_Q> and the formulas of build_src_slot."""
    w_count = q_windows_of(counts)
    cell_off = np.concatenate([[0], np.cumsum(counts)])[:-1]
    n_cls = counts.reshape(CLASSES, -1).sum(axis=1)
    assert n_cls.max() <= 128                                # one byte each
    d_cls = np.array([small_div(int(v), w_count) for v in n_cls])
    m_cls = n_cls - d_cls * w_count
    assert np.array_equal(d_cls, n_cls // w_count)
    slot = {}
    for cell in range(CELLS):
        for rank in range(counts[cell]):
            b = cell >> 5
            qb = cell_off[cell] + rank - cell_off[32 * b]
            k = small_div(int(qb), w_count)
            w = qb - k * w_count
            pos = w * int(d_cls.sum()) + int(np.minimum(m_cls, w).sum())
            for kk in range(3):
                if kk < k:
                    pos += int((n_cls > w + kk * w_count).sum())
            pos += int((n_cls[:b] > qb).sum())
            slot[(cell, rank)] = pos
    return slot, w_count


def q_window_lane(n_cls, w_count, lane, w, first):
    """phases_tiled.cuh::window_lane for ORDER_CLASS_Q: (active, slot, quarter_on, first of the next window)."""
    m = [int((n_cls > w + k * w_count).sum()) for k in range(4)]
    k, j = lane >> 3, lane & 7
    return j < m[k], first + sum(m[:k]) + j, m[k] > 0, first + sum(m)


@pytest.mark.parametrize("kind", ["poisson1", "sparse", "one_column", "one_class", "dense"])
def test_quarter_order_closed_form_matches_definition(kind):
    rng = np.random.default_rng(7 + hash(kind) % 1000)
    done = 0
    for _ in range(8):
        counts = random_counts(rng, kind).astype(int)
        if kind == "dense":
            counts = np.minimum(counts, 3)                  # keep every class below 4 * 32
        if counts.sum() == 0 or q_windows_of(counts) is None:
            continue
        done += 1
        want, w1 = q_order_by_definition(counts)
        got, w2 = q_order_by_closed_form(counts)
        assert w1 == w2 and got == {k: v[0] for k, v in want.items()}
        n = int(counts.sum())
        assert sorted(got.values()) == list(range(n))       # a permutation of the tile's slots
        n_cls = counts.reshape(CLASSES, -1).sum(axis=1)
        # the kernels' window walk claims every slot exactly once, with the lane the definition gives
        lane_of = {v[0]: (v[1], v[2]) for v in want.values()}
        first, claimed = 0, {}
        for w in range(w1):
            nxt = first
            for lane in range(32):
                active, s, quarter_on, nxt = q_window_lane(n_cls, w1, lane, w, first)
                if active:
                    assert s not in claimed
                    claimed[s] = (w, lane)
                    assert quarter_on
                if lane == 0:
                    assert active                            # idle quarters mirror lane 0: it always holds a particle
                if (lane & 7) == 0:
                    assert active == quarter_on              # ... and a quarter's first lane whenever the quarter does
            first = nxt
        assert claimed == lane_of
        # invariants: no two particles of a window share a column; no two of a quarter warp share a bank class;
        # the quarter warps in use are exactly the fullest class (the lower bound of any order)
        seen_col, seen_cls, quarters = set(), set(), set()
        for (cell, _), (s, w, lane) in want.items():
            assert (w, cell >> 2) not in seen_col
            seen_col.add((w, cell >> 2))
            assert (w, lane >> 3, cell >> 5) not in seen_cls
            seen_cls.add((w, lane >> 3, cell >> 5))
            quarters.add((w, lane >> 3))
        assert len(quarters) == int(n_cls.max())
    assert done >= 3


def test_quarter_order_falls_back_when_a_tile_is_too_crowded():
    counts = np.zeros(CELLS, dtype=int)
    counts[0:4] = 9                                          # one column of 36: W would be 36
    assert q_windows_of(counts) is None
    counts = np.full(CELLS, 4)                               # 128 per class: W = 32, the largest that fits
    assert q_windows_of(counts) == 32
    got, _ = q_order_by_closed_form(counts)
    assert sorted(got.values()) == list(range(1024))
