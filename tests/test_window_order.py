"""CPU restatement of the tile order arithmetic of fluid-rs_b200/csrc/sort.cuh (ORDER_CLASS_RR).

The kernels place a particle with a closed form (k_tile_tables builds per-tile tables, k_build_src turns
(cell, rank) into a slot).  This file restates both the DEFINITION of the order (cap every (x,y) column at
W = ceil(N/32) particles and move the rest to the tile's overflow segment; deal the remaining class-major
sequence round robin into W windows; merge the classes round robin inside a window) and the CLOSED FORM the
kernels use, and checks that they agree and that the invariants the shared-memory accumulation relies on
hold.  It does not run the CUDA code (tests/test_gpu_parity.py::test_windows_hold_distinct_columns does that
on the device); it pins the arithmetic, including the division-free floor the kernels use.
"""
import numpy as np
import pytest

CELLS, CLASSES, COLS_PER_CLASS, Z = 256, 8, 8, 4     # local_cell_3d: cell = z + 4 * (column + 8 * class)
COLS = CLASSES * COLS_PER_CLASS
MAX_MERGE_W, MAX_MERGE_N = 32, 8192                   # sort.cuh: PERM_MAX_W, and the 8192-particle bound of small_div


def small_div(x, w):
    """floor(x / w) as sort.cuh::small_div computes it: (float(x) + 0.5f) * (1.0f / float(w)), truncated."""
    inv = np.float32(1.0) / np.float32(w)
    return int(np.float32(np.float32(x) + np.float32(0.5)) * inv)


def plan(counts):
    """W and the per-column overflow of a tile (k_tile_tables): with W0 = ceil(N/32) <= 32 windows every column
    keeps its first W0 particles (cell order, then rank) and the rest go to the overflow segment; a tile too
    crowded for the merge table keeps the old rule W = max(W0, fullest column) and has no overflow."""
    n = int(counts.sum())
    col = counts.reshape(COLS, Z).sum(axis=1)
    w0 = (n + 31) // 32
    if w0 <= MAX_MERGE_W and n < MAX_MERGE_N:
        return w0, np.maximum(col - w0, 0)
    return max(w0, int(col.max())), np.zeros(COLS, dtype=int)


def order_by_definition(counts):
    """slot of every (cell, rank), from the definition of the order."""
    n = int(counts.sum())
    w_count, ovf_col = plan(counts)
    cells = np.repeat(np.arange(CELLS), counts)             # class-major cell order = the sequence q
    ranks = np.concatenate([np.arange(c) for c in counts]) if n else np.zeros(0, int)
    col_rank = np.zeros(n, dtype=int)
    seen_col = np.zeros(COLS, dtype=int)
    for q in range(n):
        c = cells[q] >> 2
        col_rank[q] = seen_col[c]
        seen_col[c] += 1
    main = [q for q in range(n) if col_rank[q] < w_count or ovf_col[cells[q] >> 2] == 0]
    over = [q for q in range(n) if q not in set(main)]
    n_main = len(main)
    per, extra = divmod(n_main, w_count)
    cls = cells >> 5
    slot = {}
    for w in range(w_count):
        members = [q for j, q in enumerate(main) if j % w_count == w]
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
    for j, q in enumerate(over):                            # overflow segment: cell order
        slot[(int(cells[q]), int(ranks[q]))] = n_main + j
    return slot, w_count, n_main


def order_by_closed_form(counts):
    """the same through the kernels' tables and formulas (k_tile_tables / k_build_src)."""
    n = int(counts.sum())
    w_count, ovf_col = plan(counts)
    col = counts.reshape(COLS, Z).sum(axis=1)
    ovf_before = np.concatenate([[0], np.cumsum(ovf_col)])  # per column (+ total at [COLS])
    n_main = n - int(ovf_before[COLS])
    per, extra = divmod(n_main, w_count)
    cell_off = np.concatenate([[0], np.cumsum(counts)])[:-1]     # uncapped cellStart (relative to the tile)
    main_col = col - ovf_col
    n_cls = main_col.reshape(CLASSES, COLS_PER_CLASS).sum(axis=1)
    s_cls = np.concatenate([[0], np.cumsum(n_cls)])[:-1]
    tab = np.zeros((w_count, CLASSES), dtype=int)           # members of class b in window w
    for b in range(CLASSES):
        sb_mod = s_cls[b] - small_div(s_cls[b], w_count) * w_count
        for w in range(w_count):
            off = w - sb_mod
            if off < 0:
                off += w_count
            tab[w, b] = small_div(n_cls[b] - off + w_count - 1, w_count) if off < n_cls[b] else 0
    assert tab.max(initial=0) <= 8                          # 8 columns per class, each at most once per window
    slot = {}
    for cell in range(CELLS):
        c = cell >> 2
        for rank in range(counts[cell]):
            col_rank = cell_off[cell] - cell_off[cell & ~3] + rank
            if ovf_col[c] > 0 and col_rank >= w_count:      # overflow segment
                slot[(cell, rank)] = n_main + ovf_before[c] + (col_rank - w_count)
                continue
            q = cell_off[cell] + rank - ovf_before[c]       # position in the main sequence
            b = cell >> 5
            s_b = cell_off[b << 5] - ovf_before[b * COLS_PER_CLASS]
            assert s_b == s_cls[b]
            w = q - small_div(q, w_count) * w_count
            s_mod = s_b - small_div(s_b, w_count) * w_count
            off = w - s_mod
            if off < 0:
                off += w_count
            k = small_div(q - (s_b + off), w_count)
            pos = int(np.minimum(tab[w], k).sum()) + int((tab[w, :b] > k).sum())
            slot[(cell, rank)] = w * per + min(w, extra) + pos
    return slot, w_count, n_main


def random_counts(rng, kind):
    if kind == "poisson1":
        return rng.poisson(0.92, CELLS)
    if kind == "sparse":
        c = np.zeros(CELLS, dtype=int)
        c[rng.integers(0, CELLS, 5)] += rng.integers(1, 4, 5)
        return c
    if kind == "one_column":                                 # one column far above W: it overflows
        c = rng.poisson(0.3, CELLS)
        c[40:44] += rng.integers(3, 9, 4)
        return c
    if kind == "one_class":
        c = np.zeros(CELLS, dtype=int)
        c[64:96] = rng.poisson(3.0, 32)
        return c
    if kind == "dense":
        return rng.poisson(3.5, CELLS)
    if kind == "crowded":                                    # too many windows for the merge table: old rule, no overflow
        return rng.poisson(6.0, CELLS)
    if kind == "compressed":                                 # what the bottom of the dam looks like: 1.1 per cell, clumpy
        return rng.poisson(rng.uniform(0.6, 1.6, CELLS))
    raise ValueError(kind)


def test_small_div_is_exact():
    for w in range(1, 33):
        x = np.arange(0, 8192)
        got = np.array([small_div(int(v), w) for v in x[:: max(1, 8192 // 997)]])
        assert np.array_equal(got, x[:: max(1, 8192 // 997)] // w)
        for v in (0, 1, w - 1, w, w + 1, 8191 - (8191 % w), 8191):
            assert small_div(v, w) == v // w


@pytest.mark.parametrize("kind", ["poisson1", "sparse", "one_column", "one_class", "dense", "crowded", "compressed"])
def test_closed_form_matches_definition(kind):
    rng = np.random.default_rng(sum(map(ord, kind)))
    overflowed = 0
    for _ in range(6):
        counts = random_counts(rng, kind).astype(int)
        if counts.sum() == 0 or counts.sum() >= 8192:
            continue
        want, w1, m1 = order_by_definition(counts)
        got, w2, m2 = order_by_closed_form(counts)
        assert w1 == w2 and m1 == m2 and got == want
        n = int(counts.sum())
        overflowed += n - m1
        assert sorted(got.values()) == list(range(n))       # a permutation of the tile's slots
        # no two particles of one window share an (x,y) column; a window holds at most 32
        per, extra = divmod(m1, w1)
        window_of_slot = np.full(n, -1, dtype=int)
        first = 0
        for w in range(w1):
            ln = per + (1 if w < extra else 0)
            assert ln <= 32
            window_of_slot[first:first + ln] = w
            first += ln
        assert first == m1
        seen = set()
        for (cell, _), s in got.items():
            if s >= m1:
                continue                                     # overflow segment: deposited with global reductions
            key = (int(window_of_slot[s]), cell >> 2)        # column = cell / 4
            assert key not in seen
            seen.add(key)
    if kind in ("one_column", "compressed"):
        assert overflowed > 0                                # the case is exercised


def test_quarter_warps_see_distinct_bank_classes_when_classes_are_balanced():
    """With every class holding 3 or 4 particles of a window, the 8 lanes of a quarter warp fall into 8
    different classes (16-byte bank groups of the float4 node tile)."""
    counts = np.zeros(CELLS, dtype=int)
    counts[::1] = 1                                          # one particle per cell: 32 per class, 4 per column
    slot, w_count, n_main = order_by_closed_form(counts)
    assert w_count == 8 and n_main == 256
    cls_of_slot = np.empty(256, dtype=int)
    for (cell, _), s in slot.items():
        cls_of_slot[s] = cell >> 5
    for w in range(8):
        lanes = cls_of_slot[32 * w:32 * w + 32]
        for qd in range(4):
            assert len(set(lanes[8 * qd:8 * qd + 8])) == 8
