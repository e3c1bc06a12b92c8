"""Shared-memory wavefront model of the tiled kernels (CPU, numpy) — a design tool, not product code.

Replays the sort's window order (sort.cuh, ORDER_CLASS_RR) on a particle cloud and counts the
shared-memory wavefronts of the inner loops of k_mass_tiled / k_p2g_tiled / k_g2p_tiled under the
hardware rules measured with tools/smem_probe.cu on B200: a 128-bit access is served per quarter
warp (8 lanes), one wavefront per distinct 16-byte address per 16-byte bank group; a 32-bit access
per warp, one wavefront per distinct word per bank.

usage: python tools/bank_model.py positions.npy origin_x origin_y origin_z [variant]
"""
import sys
import numpy as np

TX, TY, TZ = 8, 8, 4


def tile_order(lx, ly, lz, cls_of, col_rule=True):
    """Slots of one tile's particles in ORDER_CLASS_RR.  Returns (order, W, per, extra): order[s] = particle at slot s."""
    n = len(lx)
    cls = cls_of(lx, ly)
    col = ly * TX + lx
    cell = lz + 4 * (ly + 8 * cls)                 # class-major numbering needs (cls, ly) unique per column
    key = (cls * 64 + col) * 4 + lz                # class-major, then column, then z
    q_order = np.argsort(key, kind="stable")
    need = (n + 31) // 32
    colmax = np.bincount(col, minlength=64).max() if col_rule else 0
    W = max(need, colmax)
    per, extra = n // W, n % W
    q = np.arange(n)
    w = q % W
    c_sorted = cls[q_order]
    # k = index among my class's members in my window; n_b[w] = members of class b in window w
    slot = np.empty(n, dtype=np.int64)
    nb = np.zeros((W, 8), dtype=np.int64)
    k = np.empty(n, dtype=np.int64)
    for b in range(8):
        idx = np.nonzero(c_sorted == b)[0]
        if len(idx) == 0:
            continue
        wb = w[idx]
        # members of class b appear in q order; index within (class, window)
        cnt = np.zeros(W, dtype=np.int64)
        for j, ww in zip(idx, wb):
            k[j] = cnt[ww]
            cnt[ww] += 1
        nb[:, b] = cnt
    for j in range(n):
        ww, b, kk = w[j], c_sorted[j], k[j]
        pos = np.minimum(nb[ww], kk).sum() + np.count_nonzero(nb[ww, :b] > kk)
        slot[j] = ww * per + min(ww, extra) + pos
    order = np.empty(n, dtype=np.int64)
    order[slot] = q_order
    return order, W, per, extra


def wf128(idx, part):
    """idx [K,32] float4 indices, part [32] participating lanes -> wavefronts per access [K]."""
    K = idx.shape[0]
    tot = np.zeros(K, dtype=np.int64)
    for qd in range(4):
        lanes = [l for l in range(8 * qd, 8 * qd + 8) if part[l]]
        if not lanes:
            continue
        sub = idx[:, lanes]
        worst = np.zeros(K, dtype=np.int64)
        for gidx in range(8):
            m = (sub % 8) == gidx
            # distinct addresses in this group
            cnt = np.array([len(set(sub[r][m[r]])) for r in range(K)])
            worst = np.maximum(worst, cnt)
        tot += np.maximum(worst, 1)
    return tot


def wf32(idx, part):
    K = idx.shape[0]
    lanes = [l for l in range(32) if part[l]]
    if not lanes:
        return np.zeros(K, dtype=np.int64)
    sub = idx[:, lanes]
    out = np.zeros(K, dtype=np.int64)
    for r in range(K):
        best = 1
        banks = sub[r] % 32
        for b in np.unique(banks):
            best = max(best, len(set(sub[r][banks == b])))
        out[r] = best
    return out


def main():
    pos = np.load(sys.argv[1])
    org = np.array([int(a) for a in sys.argv[2:5]])
    variant = sys.argv[5] if len(sys.argv) > 5 else "base"
    max_tiles = int(sys.argv[6]) if len(sys.argv) > 6 else 150
    cell = np.floor(pos).astype(np.int64) - org
    t3 = cell // np.array([TX, TY, TZ])
    loc = cell - t3 * np.array([TX, TY, TZ])
    tid = (t3[:, 2] * 1000 + t3[:, 1]) * 1000 + t3[:, 0]
    uniq, inv = np.unique(tid, return_inverse=True)
    rng = np.random.default_rng(1)
    pick = rng.permutation(len(uniq))
    cls_of = lambda lx, ly: (lx + 2 * ly) & 7
    # strides
    NX4, PL4 = 10, 104
    if variant == "base":
        SR, SP = 12, 128
    elif variant == "s10":
        SR, SP = 10, 104
    else:
        SR, SP = 12, 128
    offs = [(ox, oy, oz) for oy in range(3) for ox in range(3) for oz in range(3)]
    O4 = np.array([ox + NX4 * oy + PL4 * oz for ox, oy, oz in offs])
    OS = np.array([ox + SR * oy + SP * oz for ox, oy, oz in offs])
    tot = dict(n=0, lanes=0, rmw_ld=0, rmw_st=0, s_ld=0, s_st=0, g2p=0, wins=0, tiles=0)
    done = 0
    for ti in pick:
        sel = np.nonzero(inv == ti)[0]
        if len(sel) < 64:
            continue
        lx, ly, lz = loc[sel, 0], loc[sel, 1], loc[sel, 2]
        order, W, per, extra = tile_order(lx, ly, lz, cls_of)
        n = len(sel)
        n4 = lx + NX4 * ly + PL4 * lz
        ns = lx + SR * ly + SP * lz
        prev4 = np.zeros(32, dtype=np.int64)
        prevs = np.zeros(32, dtype=np.int64)
        for w in range(W):
            first = w * per + min(w, extra)
            ln = per + (1 if w < extra else 0)
            ids = order[first:first + ln]
            cur4, curs = prev4.copy(), prevs.copy()
            cur4[:ln] = n4[ids]
            curs[:ln] = ns[ids]
            part = np.zeros(32, dtype=bool)
            part[:ln] = True
            allp = np.ones(32, dtype=bool)
            i4 = cur4[None, :] + O4[:, None]
            isx = curs[None, :] + OS[:, None]
            # (since the idle-lane skip of round 2 the loads are executed by the active lanes only, like the stores; before it
            # the idle lanes executed them at a mirrored address and an LDS.128 cost 4 wavefronts whatever its lanes did)
            tot["rmw_ld"] += wf128(i4, part).sum()
            tot["rmw_st"] += wf128(i4, part).sum()
            tot["s_ld"] += wf32(isx, allp).sum()
            tot["s_st"] += wf32(isx, part).sum()
            prev4, prevs = cur4, curs
            tot["wins"] += 1
            tot["lanes"] += ln
        # g2p: 32 consecutive slots at a time, LDS.128 gathers
        for s0 in range(0, n, 32):
            ids = order[s0:s0 + 32]
            cur = np.zeros(32, dtype=np.int64)
            cur[:len(ids)] = n4[ids]
            part = np.zeros(32, dtype=bool)
            part[:len(ids)] = True
            tot["g2p"] += wf128(cur[None, :] + O4[:, None], part).sum()
        tot["n"] += n
        tot["tiles"] += 1
        done += 1
        if done >= max_tiles:
            break
    n = tot["n"]
    print(f"variant {variant}: tiles {tot['tiles']} particles {n} windows {tot['wins']} fill {tot['lanes'] / (32 * tot['wins']):.3f}")
    print(f"  128-bit RMW   : LDS {tot['rmw_ld'] / n:.2f} + STS {tot['rmw_st'] / n:.2f} wavefronts/particle "
          f"(ideal at this fill {2 * 27 * 4 * tot['wins'] / n:.2f}, at fill 1: {2 * 27 * 4 / 32:.2f});  per access {tot['rmw_ld'] / (27 * tot['wins']):.2f} / {tot['rmw_st'] / (27 * tot['wins']):.2f}")
    print(f"  32-bit  gather: {tot['s_ld'] / n:.2f} wavefronts/particle; RMW {(tot['s_ld'] + tot['s_st']) / n:.2f}  (ideal {27 * tot['wins'] / n:.2f} / {54 * tot['wins'] / n:.2f}); per access {tot['s_ld'] / (27 * tot['wins']):.2f}")
    print(f"  g2p 128 gather: {tot['g2p'] / n:.2f} wavefronts/particle (ideal {27 * 4 / 32:.2f})")


if __name__ == "__main__":
    main()
