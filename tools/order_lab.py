"""Window-order laboratory (CPU, numpy) — a design tool, not product code.

Question: how many shared-memory wavefronts per 128-bit access does a tile cost under a given window order?
Rule (tools/smem_probe.cu): a 128-bit access is served per quarter warp; a quarter costs one wavefront per
distinct address in its fullest 16-byte bank group; the group of a node column is its class (x + 2y) mod 8 and
every stencil offset shifts all lanes alike, so the cost of a window is the same for each of its 63 accesses:
    cost(window) = sum over quarters of max(1 if any lane is active, max_b members of class b in the quarter).

Orders compared on the tiles of an oracle dam break (positions.npy) or a Poisson cloud:
  rr     the default: class-major cell order dealt round robin into W windows, classes merged round robin inside
         the window, lanes compact (sort.cuh ORDER_CLASS_RR)
  quart  ORDER_CLASS_Q (FLUID_B200_ORDER=q): every class is dealt from window 0 (window = q_b mod W, round = q_b div W
         with q_b the particle's place in its class's (column, z) order), a round is one quarter warp, classes ascending
         inside it; W = max(fullest column, ceil(fullest class / 4)).  It meets the lower bound with zero conflicts,
         but needs 10-13 % more windows: measured on B200 it trades 3.4 % fewer wavefronts for 25 % more instructions
         (profiles/r02_order_q_vs_rr.md) and is not the default.

usage: python tools/order_lab.py [positions.npy ox oy oz | poisson <mean per cell>]
"""
import sys
import numpy as np

TX, TY, TZ = 8, 8, 4


def tiles_of(pos, org):
    cell = np.floor(pos).astype(np.int64) - np.asarray(org)
    t3 = cell // np.array([TX, TY, TZ])
    loc = cell - t3 * np.array([TX, TY, TZ])
    tid = (t3[:, 2] * 4096 + t3[:, 1]) * 4096 + t3[:, 0]
    order = np.argsort(tid, kind="stable")
    tid_s = tid[order]
    starts = np.flatnonzero(np.r_[True, tid_s[1:] != tid_s[:-1]])
    ends = np.r_[starts[1:], len(tid_s)]
    for a, b in zip(starts, ends):
        sel = order[a:b]
        yield loc[sel, 0], loc[sel, 1], loc[sel, 2]


def cost_of(windows):
    """windows: list of arrays [32] of class per lane (-1 = idle) -> wavefronts per access."""
    tot = 0
    for w in windows:
        for qd in range(4):
            c = w[8 * qd:8 * qd + 8]
            c = c[c >= 0]
            if len(c):
                tot += np.bincount(c, minlength=8).max()
    return tot


def order_rr(lx, ly, lz):
    n = len(lx)
    cls = (lx + 2 * ly) & 7
    col = ly * TX + lx
    key = (cls * 64 + col) * 4 + lz
    qo = np.argsort(key, kind="stable")
    colmax = np.bincount(col, minlength=64).max()
    W = max((n + 31) // 32, colmax)
    c_sorted = cls[qo]
    w = np.arange(n) % W
    wins = []
    for ww in range(W):
        mem = c_sorted[w == ww]                      # in q order: class ascending
        nb = np.bincount(mem, minlength=8)
        lanes = []
        for k in range(nb.max() if len(mem) else 0):
            lanes += [b for b in range(8) if nb[b] > k]
        arr = -np.ones(32, dtype=np.int64)
        arr[:len(lanes)] = lanes
        wins.append(arr)
    return wins, W


def order_quart(lx, ly, lz, align0=True):
    n = len(lx)
    cls = (lx + 2 * ly) & 7
    col = ly * TX + lx
    colmax = np.bincount(col, minlength=64).max()
    nb = np.bincount(cls, minlength=8)
    W = max((n + 31) // 32, colmax, (nb.max() + 3) // 4)
    wins = [-np.ones(32, dtype=np.int64) for _ in range(W)]
    for ww in range(W):
        for k in range(4):
            present = [b for b in range(8) if nb[b] > ww + k * W]
            wins[ww][8 * k:8 * k + len(present)] = present
    return wins, W


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "poisson":
        mean = float(sys.argv[2])
        rng = np.random.default_rng(3)
        n = int(mean * 64 * 64 * 64)
        pos = rng.random((n, 3)) * 64
        org = (0, 0, 0)
    else:
        pos = np.load(sys.argv[1])
        org = tuple(int(a) for a in sys.argv[2:5])
    res = {"rr": [0, 0, 0], "quart": [0, 0, 0]}
    n_tot = tiles = 0
    lb = 0
    for lx, ly, lz in tiles_of(pos, org):
        n = len(lx)
        if n < 16:
            continue
        tiles += 1
        n_tot += n
        nb = np.bincount((lx + 2 * ly) & 7, minlength=8)
        lb += max((n + 7) // 8, nb.max())
        for name, fn in (("rr", order_rr), ("quart", order_quart)):
            wins, W = fn(lx, ly, lz)
            res[name][0] += cost_of(wins)
            res[name][1] += W
            res[name][2] += sum(int((w >= 0).sum()) for w in wins)
    print(f"tiles {tiles}  particles {n_tot}  per tile {n_tot / tiles:.1f}")
    print(f"lower bound max(ceil(N/8), fullest class): {lb / n_tot * 8:.3f} x N/8")
    for name, (c, w, lanes) in res.items():
        assert lanes == n_tot
        print(f"{name:6s} wavefronts per access = {c / n_tot * 8:.3f} x N/8   windows per tile {w / tiles:.2f}   "
              f"lane fill {n_tot / (32 * w):.3f}   63 accesses -> {63 * c / n_tot:.2f} wavefronts per particle")




def g2p_compact_cost(wins):
    """g2p walks the tile's slots 32 at a time whatever the windows are: cost of that walk."""
    seq = np.concatenate([w[w >= 0] for w in wins])
    tot = 0
    for s0 in range(0, len(seq), 8):
        tot += np.bincount(seq[s0:s0 + 8], minlength=8).max()
    return tot


def main_g2p():
    pos = np.load(sys.argv[2])
    org = tuple(int(a) for a in sys.argv[3:6])
    res = {"rr": 0, "quart": 0}
    n_tot = 0
    for lx, ly, lz in tiles_of(pos, org):
        if len(lx) < 16:
            continue
        n_tot += len(lx)
        res["rr"] += g2p_compact_cost(order_rr(lx, ly, lz)[0])
        res["quart"] += g2p_compact_cost(order_quart(lx, ly, lz)[0])
    for k, v in res.items():
        print(f"g2p compact walk under {k}: {v / n_tot * 8:.3f} x N/8")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "g2p":
        main_g2p()
    else:
        main()
