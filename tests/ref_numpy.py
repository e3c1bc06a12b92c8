"""Second, independent restatement of the reference's five phases in plain Python loops over
numpy float32 scalars — tiny cases only.  Written from src/3d_multi.rs:136-381 (and the 2D
file) separately from oracle/oracle.cpp so that a transcription slip in either one (row/column
order of the affine matrix, sign of x_n - x_p, weight index) shows up as a disagreement.
Dense grid, no block structure: particle order is the input order, which for one block and no
migration is the reference's order too."""
import math

import numpy as np

F = np.float32


def _weights(c):
    return [F(0.5) * (F(0.5) - c) * (F(0.5) - c), F(0.75) - c * c, F(0.5) * (F(0.5) + c) * (F(0.5) + c)]


class RefSim:
    def __init__(self, cfg: dict, origin, size):
        self.cfg = cfg
        self.d = cfg["dim"]
        self.origin = list(origin)
        self.size = list(size)
        n = int(np.prod(size))
        self.gvel = np.zeros((n, self.d), dtype=np.float32)
        self.gmass = np.zeros(n, dtype=np.float32)
        self.P = []   # dicts: pos, vel, C (C[col][row]), mass

    def add(self, rec):
        d = self.d
        rec = np.asarray(rec, dtype=np.float32)
        C = rec[2 * d:2 * d + d * d].reshape(d, d).copy()   # C[col][row]
        self.P.append(dict(pos=rec[:d].copy(), vel=rec[d:2 * d].copy(), C=C, mass=F(rec[-1])))

    def _stencil(self, p):
        d = self.d
        cell = [int(math.floor(float(p["pos"][a]))) for a in range(d)]
        cd = [p["pos"][a] - (F(cell[a]) + F(0.5)) for a in range(d)]
        W = [_weights(cd[a]) for a in range(d)]
        out = []
        rng = [(x, y) for y in range(3) for x in range(3)] if d == 2 else \
              [(x, y, z) for z in range(3) for y in range(3) for x in range(3)]
        for n in rng:
            cn = [cell[a] + n[a] - 1 for a in range(d)]
            dn = [p["pos"][a] - (F(cn[a]) + F(0.5)) for a in range(d)]
            w = W[0][n[0]] * W[1][n[1]]
            if d == 3:
                w = w * W[2][n[2]]
            rel = [cn[a] - self.origin[a] for a in range(d)]
            if any(r < 0 or r >= self.size[a] for a, r in enumerate(rel)):
                idx = -1
            else:
                idx = rel[0] + rel[1] * self.size[0] + (rel[2] * self.size[0] * self.size[1] if d == 3 else 0)
            out.append((idx, w, dn))
        return out

    def _matvec(self, M, v):
        d = self.d
        res = []
        for r in range(d):
            acc = M[0][r] * v[0]
            for c in range(1, d):
                acc = acc + M[c][r] * v[c]
            res.append(acc)
        return res

    def substep(self):
        cfg, d = self.cfg, self.d
        self.gvel[:] = 0
        self.gmass[:] = 0
        touched = []
        for p in self.P:                                       # p2g_1
            for idx, w, dn in self._stencil(p):
                q = self._matvec(p["C"], [-x for x in dn])
                mc = w * p["mass"]
                if idx >= 0:
                    self.gmass[idx] += mc
                    for a in range(d):
                        self.gvel[idx, a] += mc * (p["vel"][a] + q[a])
                    touched.append(idx)
        self.density = []
        self.pressure = []
        for p in self.P:                                       # p2g_2
            st = self._stencil(p)
            rho = F(0.0)
            for idx, w, dn in st:
                if idx >= 0:
                    rho = rho + self.gmass[idx] * w
            vol = p["mass"] / rho
            eos = F(cfg["eos_stiffness"]) * (F(math.pow(float(rho / F(cfg["rest_density"])), cfg["eos_power"])) - F(1.0))
            pr = max(F(cfg["pressure_clamp"]), eos)
            self.density.append(rho)
            self.pressure.append(pr)
            T = [[None] * d for _ in range(d)]
            for c in range(d):
                for r in range(d):
                    strain = p["C"][c][r] + p["C"][r][c]
                    visc = F(cfg["dynamic_viscosity"]) * strain
                    stress = (-pr) * F(1.0 if c == r else 0.0) + visc
                    T[c][r] = (F(-4.0) * vol) * stress * F(cfg["dt"])
            for idx, w, dn in st:
                if idx >= 0:
                    M = [[w * T[c][r] for r in range(d)] for c in range(d)]
                    f = self._matvec(M, [-x for x in dn])
                    for a in range(d):
                        self.gvel[idx, a] += f[a]
        done = set()
        for idx in touched:                                    # update_grid
            if idx not in done and self.gmass[idx] > 0:
                self.gvel[idx] = self.gvel[idx] / self.gmass[idx]
                for a in range(d):
                    self.gvel[idx, a] += F(cfg["dt"]) * F(cfg["gravity"][a])
                done.add(idx)
        for p in self.P:                                       # g2p
            st = self._stencil(p)
            vel = [F(0.0)] * d
            B = [[F(0.0)] * d for _ in range(d)]
            for idx, w, dn in st:
                if idx >= 0:
                    wv = [self.gvel[idx, a] * w for a in range(d)]
                    for c in range(d):
                        for r in range(d):
                            B[c][r] = B[c][r] + wv[r] * (-dn[c])
                    vel = [vel[a] + wv[a] for a in range(d)]
            p["C"] = np.array([[F(4.0) * B[c][r] for r in range(d)] for c in range(d)], dtype=np.float32)
            p["vel"] = np.array(vel, dtype=np.float32)
            p["pos"] = np.array([p["pos"][a] + vel[a] * F(cfg["dt"]) for a in range(d)], dtype=np.float32)
            for a in range(d):
                x = p["pos"][a]
                x = x if x > F(cfg["clip_min"][a]) else F(cfg["clip_min"][a])
                x = x if x < F(cfg["clip_max"][a]) else F(cfg["clip_max"][a])
                p["pos"][a] = x
                nxt = x + p["vel"][a]
                wmin = F(cfg["clip_min"][a]) + F(cfg["boundary_damp_dist"])
                wmax = F(cfg["clip_max"][a]) - F(cfg["boundary_damp_dist"])
                if nxt < wmin:
                    p["vel"][a] += wmin - nxt
                if nxt > wmax:
                    p["vel"][a] += wmax - nxt

    def records(self):
        d = self.d
        out = np.zeros((len(self.P), 2 * d + d * d + 1), dtype=np.float32)
        for i, p in enumerate(self.P):
            out[i, :d] = p["pos"]
            out[i, d:2 * d] = p["vel"]
            out[i, 2 * d:2 * d + d * d] = p["C"].reshape(-1)
            out[i, -1] = p["mass"]
        return out
