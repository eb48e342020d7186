#!/usr/bin/env python
"""Thread-by-thread emulation (NumPy, CPU) of the index arithmetic of csrc/conv_small.cu.

The build container has no GPU, so a wrong offset in a new kernel costs a round trip to the GPU box.
This script replays what every thread of small_fprop_kernel / small_bwd_kernel reads and writes --
same item decoding, same shared-memory layouts, same guards, stale data between stages modelled as
huge finite numbers -- on tiny geometries and compares with the oracle.  It checks indexing, not speed.

    python tools/emu_conv_small.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import theanet_oracle as O   # noqa: E402

KST, F = 256, 3
STALE = np.float32(3e30)


def cdiv(a, b):
    return (a + b - 1) // b


def guard_x(S):
    return (S + 8 + 3) // 4 * 4


def act_fwd(z, nn):
    return z if z > 0 else np.float32(np.float32(z * np.float32(nn)) / np.float32(100))


def act_bwd(a, nn):
    s = np.float32(nn) / np.float32(100)
    return np.float32(1) if a > 0 else (s if a < 0 else np.float32(1) + s)


def fprop(x, W, bias, nn, P, NB, grid):
    B, C, S, _ = x.shape
    M = W.shape[0]
    Oo = S - F + 1
    Pc, G = (Oo + 1) // 2, (M + 3) // 4
    coP, SS, CSS = 4 * G, S * S, C * S * S
    a = np.full((B, M, Oo, Oo), np.nan, np.float32)
    pooled = np.full((B, M, P, P), np.nan, np.float32)
    ws = np.zeros(C * F * F * coP, np.float32)
    for t in range(ws.size):
        co, r = t % coP, t // coP
        v, r = r % F, r // F
        u, c = r % F, r // F
        ws[t] = W[co, c, F - 1 - u, F - 1 - v] if co < M else 0
    xf = x.reshape(-1)
    for blk in range(grid):
        xs = np.full(NB * CSS + guard_x(S), STALE, np.float32)
        xs[NB * CSS:] = 0
        grp = blk
        while grp * NB < B:
            b0 = grp * NB
            nb = min(NB, B - b0)
            xs[:nb * CSS] = xf[b0 * CSS:(b0 + nb) * CSS]
            for it in range(G * NB * Pc * Pc):
                q1, cell = divmod(it, Pc * Pc)
                pi, pj = divmod(cell, Pc)
                g, b = divmod(q1, NB)
                if b >= nb:
                    continue
                acc = np.zeros((4, 4), np.float32)
                xb = b * CSS + 2 * pi * S + 2 * pj
                for c in range(C):
                    p = np.array([[xs[xb + r * S + e] for e in range(F + 1)] for r in range(F + 1)])
                    for u in range(F):
                        for v in range(F):
                            w = ws[(((c * F + u) * F + v) * G + g) * 4:][:4]
                            for dy in range(2):
                                for dx in range(2):
                                    acc[dy * 2 + dx] += p[dy + u, dx + v] * w
                    xb += SS
                i0, j0 = 2 * pi, 2 * pj
                r1, c1 = i0 + 1 < Oo, j0 + 1 < Oo
                for q in range(4):
                    m = 4 * g + q
                    if m >= M:
                        break
                    vals = [act_fwd(np.float32(acc[pp, q] + bias[m]), nn) for pp in range(4)]
                    mx = vals[0]
                    a[b0 + b, m, i0, j0] = vals[0]
                    if c1:
                        a[b0 + b, m, i0, j0 + 1] = vals[1]
                        mx = max(mx, vals[1])
                    if r1:
                        a[b0 + b, m, i0 + 1, j0] = vals[2]
                        mx = max(mx, vals[2])
                        if c1:
                            a[b0 + b, m, i0 + 1, j0 + 1] = vals[3]
                            mx = max(mx, vals[3])
                    if pi < P and pj < P:
                        pooled[b0 + b, m, pi, pj] = mx
            grp += grid
    return a, pooled


def bwd(x, a, pooled, dtop, W, nn, NB, grid, L, need_dx):
    B, C, S, _ = x.shape
    M = W.shape[0]
    Oo, P = a.shape[-1], pooled.shape[-1]
    G, CG = (M + 3) // 4, (C + 3) // 4
    cP, mP, pd, FF = 4 * CG, 4 * G, F - 1, F * F
    Hp = Oo + 2 * pd
    ps = 4 * G
    if (ps // 4) % 2 == 0:
        ps += 4
    gg, gx = 4 * ps, guard_x(S)
    SS, CSS, OO, PP = S * S, C * S * S, Oo * Oo, P * P
    gzimg, grow = Hp * Hp * ps, Hp * ps
    T = G * C
    nsl = max(1, KST // T)
    nseg = cdiv(Oo, L)
    upi = ((Oo + 1) // 2) * nseg
    strips = cdiv(S, 4)
    npl = max(1, KST // G)
    nW = T * 4 * FF
    nout = nW + mP
    partial = np.zeros((grid, nout), np.float64)
    dx = np.full(x.shape, np.nan, np.float32)
    xf, af, pf, df = x.reshape(-1), a.reshape(-1), pooled.reshape(-1), dtop.reshape(-1)
    wd = np.zeros(mP * FF * cP, np.float32)
    for t in range(wd.size):
        co, r = t % cP, t // cP
        v, r = r % F, r // F
        u, m = r % F, r // F
        wd[t] = W[m, co, u, v] if (co < C and m < M) else 0
    for blk in range(grid):
        gz = np.zeros(NB * gzimg + gg, np.float32)
        xs = np.full(NB * CSS + gx, STALE, np.float32)
        xs[NB * CSS:] = 0
        acc = np.zeros((KST, 4, FF), np.float64)
        dba = np.zeros((KST, 4), np.float64)
        grp = blk
        while grp * NB < B:
            b0 = grp * NB
            nb = min(NB, B - b0)
            xs[:nb * CSS] = xf[b0 * CSS:(b0 + nb) * CSS]
            for t in range(nb * M * PP):
                po, d = pf[b0 * M * PP + t], df[b0 * M * PP + t]
                bm, p = divmod(t, PP)
                pi, pj = divmod(p, P)
                b, m = divmod(bm, M)
                ggv = np.float32(d * act_bwd(po, nn))
                ar = (b0 * M + bm) * OO + 2 * pi * Oo + 2 * pj
                gr = b * gzimg + ((2 * pi + pd) * Hp + 2 * pj + pd) * ps + m
                r1, c1 = 2 * pi + 1 < Oo, 2 * pj + 1 < Oo
                gz[gr] = ggv if af[ar] == po else 0
                if c1:
                    gz[gr + ps] = ggv if af[ar + 1] == po else 0
                if r1:
                    gz[gr + grow] = ggv if af[ar + Oo] == po else 0
                    if c1:
                        gz[gr + grow + ps] = ggv if af[ar + Oo + 1] == po else 0
            for tid in range(KST):
                if tid < T * nsl:
                    sl, combo = divmod(tid, T)
                    mg, c = divmod(combo, C)
                    un = sl
                    while un < nb * upi:
                        b, r = divmod(un, upi)
                        rp, seg = divmod(r, nseg)
                        i0, j0 = 2 * rp, seg * L
                        ln = min(L, Oo - j0)
                        xr = (b * C + c) * SS + i0 * S + j0
                        g0 = b * gzimg + ((i0 + pd) * Hp + pd + j0) * ps + 4 * mg
                        for j in range(ln):
                            ga, gb = gz[g0 + j * ps:][:4], gz[g0 + grow + j * ps:][:4]
                            for u in range(F):
                                for v in range(F):
                                    x0 = xs[xr + u * S + j + v]
                                    x1 = xs[xr + (u + 1) * S + j + v]
                                    acc[tid, :, u * F + v] += ga.astype(np.float64) * x0 + gb.astype(np.float64) * x1
                        un += nsl
                if tid < G * npl:
                    pl, mgd = divmod(tid, G)
                    for px in range(pl, nb * Hp * Hp, npl):
                        dba[tid] += gz[px * ps + 4 * mgd:][:4]
            if need_dx:
                for it in range(CG * NB * S * strips):
                    r, s = divmod(it, strips)
                    r2, y = divmod(r, S)
                    cg, b = divmod(r2, NB)
                    if b >= nb:
                        continue
                    x0 = 4 * s
                    ac = np.zeros((4, 4), np.float64)
                    gb0 = b * gzimg + (y * Hp + x0) * ps
                    for m4 in range(G):
                        for u in range(F):
                            gv = [gz[gb0 + (u * Hp + e) * ps + 4 * m4:][:4] for e in range(4 + F - 1)]
                            for v in range(F):
                                base = ((((4 * m4) * F + u) * F + v) * CG + cg) * 4
                                w = [wd[base + q * FF * CG * 4:][:4] for q in range(4)]
                                for l in range(4):
                                    for q in range(4):
                                        ac[l] += float(gv[l + v][q]) * w[q].astype(np.float64)
                    for q in range(4):
                        ch = 4 * cg + q
                        if ch >= C:
                            break
                        for l in range(4):
                            if x0 + l < S:
                                dx[b0 + b, ch, y, x0 + l] = ac[l, q]
            # stale data of this stage stays behind for the next one (finite, as on the device)
            grp += grid
        red = acc.reshape(-1)[:T * nsl * 4 * FF].reshape(nsl, nW) if T * nsl * 4 * FF <= acc.size else None
        full = np.zeros((nsl, nW))
        for tid in range(T * nsl):
            sl, combo = divmod(tid, T)
            full[sl, combo * 4 * FF:(combo + 1) * 4 * FF] = acc[tid].reshape(-1)
        partial[blk, :nW] = full.sum(0)
        dbr = np.zeros((npl, mP))
        for tid in range(G * npl):
            pl, mgd = divmod(tid, G)
            dbr[pl, 4 * mgd:4 * mgd + 4] = dba[tid]
        partial[blk, nW:] = dbr.sum(0)
    tot = partial.sum(0)
    dW = np.zeros(W.shape, np.float64)
    db = np.zeros(M, np.float64)
    for o in range(nout):
        if o < nW:
            e, r = o % FF, o // FF
            q, r = r & 3, r >> 2
            cc, g = r % C, r // C
            m, u, v = 4 * g + q, e // F, e % F
            if m < M:
                dW[m, cc, F - 1 - u, F - 1 - v] = tot[o]
        elif o < nW + M:
            db[o - nW] = tot[o]
    return dW, db, dx


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - b)) / max(np.max(np.abs(b)), 1e-30))


def main():
    cases = [  # B, C, S, M, ignore_border, NB, grid, L
        (5, 2, 7, 5, False, 2, 2, 5),
        (3, 1, 8, 4, False, 2, 1, 2),
        (4, 3, 9, 6, True, 3, 2, 3),
        (3, 5, 6, 9, False, 1, 2, 4),
    ]
    for case, (B, C, S, M, ib, NB, grid, L) in enumerate(cases):
        rng = np.random.default_rng(case)
        x = (rng.integers(-3, 4, (B, C, S, S)) / 4).astype(np.float32)
        W = (rng.integers(-2, 3, (M, C, F, F)) / 4).astype(np.float32)
        b = ((2 * rng.integers(-2, 3, M) + 1) / 32).astype(np.float32)
        actn, nn = 'relu05', 5
        z, cache = O.conv_forward(x, W, 'valid')
        a = O.act_forward(actn, z + b[None, :, None, None])
        pooled, pcache = O.pool_forward(a, 2, ib)
        P = pooled.shape[-1]
        ea, ep = fprop(x, W, b, nn, P, NB, grid)
        assert not np.isnan(ea).any() and not np.isnan(ep).any(), case
        assert rel(ea, a) < 1e-6, (case, rel(ea, a))
        assert np.array_equal(ep, O.pool_forward(ea, 2, ib)[0]), case
        dtop = rng.standard_normal(pooled.shape).astype(np.float32)
        da = O.pool_backward(dtop, pcache)
        gz = O.act_backward(actn, z + b[None, :, None, None], a, da)
        dW, db, dx = O.conv_backward(gz, W, cache)
        for need_dx in (True, False):
            eW, eb, ex = bwd(x, a, pooled, dtop, W, nn, NB, grid, L, need_dx)
            assert rel(eW, dW) < 1e-5, (case, 'dW', rel(eW, dW))
            assert rel(eb, db) < 1e-5, (case, 'db', rel(eb, db))
            if need_dx:
                assert not np.isnan(ex).any(), case
                assert rel(ex, dx) < 1e-5, (case, 'dx', rel(ex, dx))
        print('case', case, 'ok')


if __name__ == '__main__':
    main()
