#!/usr/bin/env python
"""CPU emulation of the lane-level data flow of solve_kernel2 (csrc/solve_kernel_v2.cuh): DMMA m8n8k4 fragment
ownership, the Chebyshev-product c_k tiles with their shuffle fix-up, and the stride-4 in-register cos/sin
chains of the gradient.  Checks the index algebra against the plain formulas before any GPU time is spent."""
import numpy as np

rng = np.random.default_rng(1)
LANES = np.arange(32)
G, Q = LANES >> 2, LANES & 3


def dmma(acc, a, b):
    """acc: (32, 2) per-lane {D[g][2q], D[g][2q+1]}; a: (32,) A[g][q]; b: (32,) B[q][g]"""
    A = np.zeros((8, 4)); B = np.zeros((4, 8))
    A[G, Q] = a
    B[Q, G] = b
    D = A @ B
    acc[:, 0] += D[G, 2 * Q]
    acc[:, 1] += D[G, 2 * Q + 1]


def shfl(v, src):
    return v[src]


def ck_emulation(nb, T):
    tiles = (nb + 7) // 8
    thy, thx = rng.uniform(0, np.pi, T), rng.uniform(0, np.pi, T)
    cy = np.cos(np.outer(np.arange(8 * tiles + 1), thy))  # [order][state]
    cx = np.cos(np.outer(np.arange(8 * tiles + 1), thx))
    acc = np.zeros((tiles, tiles, 32, 2))
    Tp = (T + 3) // 4 * 4
    cyp = np.zeros((cy.shape[0], Tp)); cyp[:, :T] = cy
    cxp = np.zeros((cx.shape[0], Tp)); cxp[:, :T] = cx
    for s in range(Tp // 4):
        st = 4 * s + Q  # the state of this lane's fragment element
        A0, B0 = cyp[G, st], cxp[G, st]
        Aa = [A0] + [A0 * cyp[8 * a, st] for a in range(1, tiles)]
        Bb = [B0] + [B0 * cxp[8 * b, st] for b in range(1, tiles)]
        for a in range(tiles):
            for b in range(tiles):
                dmma(acc[a, b], Aa[a], Bb[b])
    # row pass
    for a in range(1, tiles):
        for b in range(tiles):
            for e in range(2):
                other = shfl(acc[a - 1, b, :, e], 4 * ((8 - G) & 7) + Q)
                acc[a, b, :, e] = np.where(G > 0, 2 * acc[a, b, :, e] - other, acc[a, b, :, e])
    # column pass
    for b in range(1, tiles):
        for a in range(tiles):
            o0 = shfl(acc[a, b - 1, :, 0], 4 * G + ((4 - Q) & 3))
            o1 = shfl(acc[a, b - 1, :, 1], 4 * G + (3 - Q))
            acc[a, b, :, 0] = np.where(Q > 0, 2 * acc[a, b, :, 0] - o0, acc[a, b, :, 0])
            acc[a, b, :, 1] = 2 * acc[a, b, :, 1] - o1
    C = np.zeros((8 * tiles, 8 * tiles))
    for a in range(tiles):
        for b in range(tiles):
            for e in range(2):
                C[8 * a + G, 8 * b + 2 * Q + e] = acc[a, b, :, e]
    want = cy[:8 * tiles] @ cx[:8 * tiles].T
    err = np.abs(C[:nb, :nb] - want[:nb, :nb]).max() / np.abs(want).max()
    return err


def grad_emulation(nb, ntile_steps=8):
    """one 8-step tile: R1 = SA . SX, R2 = SB . CX with the x tables generated in registers (stride-4 chains),
    y tables [ky][slot]; returns max error of e_x, e_y against the direct double sums"""
    tiles, gks = (nb + 7) // 8, (nb + 3) // 4
    S = rng.normal(size=(nb, nb))
    ax, by = 0.37, 0.23
    thx, thy = rng.uniform(0, np.pi, 8), rng.uniform(0, np.pi, 8)  # alpha*x_t, beta*y_t of the 8 steps
    # A fragments SA[mi][ks] (lane (g,q) holds S[8mi+g][4ks+q] * a_kx), SB likewise with b_ky
    Sp = np.zeros((8 * tiles, 4 * gks)); Sp[:nb, :nb] = S
    SA = [[Sp[8 * mi + G, 4 * ks + Q] * ((4 * ks + Q) * ax) for ks in range(gks)] for mi in range(tiles)]
    SB = [[Sp[8 * mi + G, 4 * ks + Q] * ((8 * mi + G) * by) for ks in range(gks)] for mi in range(tiles)]
    # in-register stride-4 chains of lane (g, q) for time step g
    c, s = np.cos(thx[G]), np.sin(thx[G])
    c2, s2 = 2 * c * c - 1, 2 * s * c
    c3, s3 = 2 * c * c2 - c, 2 * c * s2 - s
    c4, s4 = 2 * c2 * c2 - 1, 2 * s2 * c2
    X0 = np.choose(Q, [np.ones(32), c, c2, c3]); Xm = np.choose(Q, [c4, c3, c2, c])
    Z0 = np.choose(Q, [np.zeros(32), s, s2, s3]); Zm = -np.choose(Q, [s4, s3, s2, s])
    m4 = 2 * c4
    R1 = np.zeros((tiles, 32, 2)); R2 = np.zeros((tiles, 32, 2))
    X, Z = X0, Z0
    for ks in range(gks):
        for mi in range(tiles):
            dmma(R1[mi], SA[mi][ks], Z)
            dmma(R2[mi], SB[mi][ks], X)
        X, Xm = m4 * X - Xm, X
        Z, Zm = m4 * Z - Zm, Z
    # epilogue: lane (g, q) holds R[ky = 8mi+g][t = 2q+e]; y tables [ky][slot]
    ex = np.zeros(8); ey = np.zeros(8)
    for mi in range(tiles):
        for e in range(2):
            ky = 8 * mi + G
            ok = ky < nb
            t = 2 * Q + e
            cyv = np.cos(ky * thy[t]); syv = np.sin(ky * thy[t])
            np.add.at(ex, t, np.where(ok, cyv * R1[mi][:, e], 0.0))
            np.add.at(ey, t, np.where(ok, syv * R2[mi][:, e], 0.0))
    kx, ky = np.arange(nb), np.arange(nb)
    wx = np.array([np.cos(ky * thy[t]) @ (S * (kx * ax)[None, :]) @ np.sin(kx * thx[t]) for t in range(8)])
    wy = np.array([(np.sin(ky * thy[t]) * (ky * by)) @ S @ np.cos(kx * thx[t]) for t in range(8)])
    return max(np.abs(ex - wx).max(), np.abs(ey - wy).max()) / max(np.abs(wx).max(), np.abs(wy).max())


if __name__ == "__main__":
    for nb in (8, 10, 12, 16, 20, 24, 32):
        print(f"nb {nb:2d}: c_k fix-up max rel err {ck_emulation(nb, 150):.2e}   gradient tile max rel err {grad_emulation(nb):.2e}")
