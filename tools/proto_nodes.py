"""Numerical prototype of the tile-node TPS evaluation (fp64: isolates the approximation error from rounding).
far field of a 32x8 tile = tensor-product polynomial interpolation from NX x NY Chebyshev nodes; control points within the
near box of the tile are evaluated exactly per pixel."""
import sys
import numpy as np
sys.path.insert(0, '.')
from oracle import dvsg_oracle as O


def cheb(n, lo, hi):
    k = np.arange(n)
    return 0.5 * (lo + hi) - 0.5 * (hi - lo) * np.cos((2 * k + 1) * np.pi / (2 * n))


def lagrange(nodes, x):
    """L[a, i] = weight of node a at point x[i]"""
    L = np.ones((len(nodes), len(x)))
    for a in range(len(nodes)):
        for b in range(len(nodes)):
            if a != b:
                L[a] *= (x - nodes[b]) / (nodes[a] - nodes[b])
    return L


def phi(d2):
    return d2 * np.log(d2 + 1e-6)


def run(H, W, m, amp, NX, NY, near_x_px, near_y_px, seed=0, tiles=400, TW=32, TH=8, pad=0.5):
    rng = np.random.default_rng(seed)
    coord = O.regular_mesh(m, m).astype(np.float64)[None]
    vec = rng.uniform(-amp, amp, coord.shape)
    T = O.tps_solve(coord, coord + vec, dtype=np.float64)[0]
    cx, cy = T[0, 3:], T[1, 3:]
    px, py = coord[0, :, 0], coord[0, :, 1]
    sx, sy = 2.0 / (W - 1), 2.0 / (H - 1)
    xn = cheb(NX, -pad, TW - 1 + pad)
    yn = cheb(NY, -pad, TH - 1 + pad)
    Lx = lagrange(xn, np.arange(TW, dtype=np.float64))      # [NX, TW]
    Ly = lagrange(yn, np.arange(TH, dtype=np.float64))      # [NY, TH]
    worst, near_counts = 0.0, []
    ntx, nty = W // TW, H // TH
    # tiles: random + the ones closest to control points
    cand = [(int(rng.integers(0, ntx)), int(rng.integers(0, nty))) for _ in range(tiles)]
    for k in range(m * m):
        ccol, crow = (px[k] + 1) / sx, (py[k] + 1) / sy
        for dx in (-2, -1, 0, 1, 2):
            for dy in (-5, -4, -3, -2, -1, 0, 1, 2, 3, 4, 5):
                cand.append((min(max(int(ccol // TW) + dx, 0), ntx - 1), min(max(int(crow // TH) + dy, 0), nty - 1)))
    for tx, ty in set(cand):
        col0, row0 = tx * TW, ty * TH
        # near set: control point within the tile box grown by near_x_px / near_y_px
        ccol, crow = (px + 1) / sx, (py + 1) / sy
        near = (ccol > col0 - near_x_px) & (ccol < col0 + TW - 1 + near_x_px) & (crow > row0 - near_y_px) & (crow < row0 + TH - 1 + near_y_px)
        near_counts.append(int(near.sum()))
        far = ~near
        # exact
        xs = -1 + sx * (col0 + np.arange(TW))
        ys = -1 + sy * (row0 + np.arange(TH))
        X, Y = np.meshgrid(xs, ys)
        d2 = (X[..., None] - px) ** 2 + (Y[..., None] - py) ** 2
        P = phi(d2)
        exact_x, exact_y = P @ cx, P @ cy
        # nodes
        xn_n = -1 + sx * (col0 + xn)
        yn_n = -1 + sy * (row0 + yn)
        XN, YN = np.meshgrid(xn_n, yn_n)
        d2n = (XN[..., None] - px[far]) ** 2 + (YN[..., None] - py[far]) ** 2
        Pn = phi(d2n)
        Fx, Fy = Pn @ cx[far], Pn @ cy[far]                 # [NY, NX]
        ax = Ly.T @ Fx @ Lx + P[..., near] @ cx[near]
        ay = Ly.T @ Fy @ Lx + P[..., near] @ cy[near]
        worst = max(worst, np.abs(ax - exact_x).max(), np.abs(ay - exact_y).max())
    return worst, np.mean(near_counts), np.max(near_counts), np.abs(cx).sum()


if __name__ == '__main__':
    for (H, W, m, amp) in [(288, 512, 4, 0.1), (288, 512, 5, 0.1), (720, 1280, 4, 0.1), (1080, 1920, 4, 0.1), (1080, 1920, 5, 0.1), (2160, 3840, 16, 0.02), (288, 512, 16, 0.02), (135, 240, 16, 0.02), (90, 160, 8, 0.03)]:
        for NX, NY in [(6, 5)]:
            for nx_px, ny_px in [(48, 24), (48, 32), (64, 24)]:
                w, mean_near, max_near, mass = run(H, W, m, amp, NX, NY, nx_px, ny_px)
                print('%4dx%4d m=%2d nodes %dx%d near +-(%3d,%3d)px: max approx err %.2e  near cps/tile mean %.2f max %d  sum|c| %.2f' % (H, W, m, NX, NY, nx_px, ny_px, w, mean_near, max_near, mass))
