"""fp64 prototype of the TWO-LEVEL node evaluation for large separable meshes (16 x 16): the far-far field of a super-tile
(16 tiles = 512 x 8 px) on NX2 x 5 super-nodes, interpolated in x to the 6 x 5 nodes of each tile; control points within
the grown super-tile box evaluated at the tile nodes; tile-near ones per pixel.  Reports the worst coordinate error."""
import sys
import numpy as np
sys.path.insert(0, '.')
from oracle import dvsg_oracle as O
from tools.proto_nodes import cheb, lagrange, phi


def run(H, W, m, amp, NX2=16, ST=16, snear_x=128.0, snear_y=128.0, near_x=48.0, near_y=24.0, seed=0, strips=24):
    rng = np.random.default_rng(seed)
    coord = O.regular_mesh(m, m).astype(np.float64)[None]
    vec = rng.uniform(-amp, amp, coord.shape)
    T = O.tps_solve(coord, coord + vec, dtype=np.float64)[0]
    cx, cy = T[0, 3:], T[1, 3:]
    px, py = coord[0, :, 0], coord[0, :, 1]
    sx, sy = 2.0 / (W - 1), 2.0 / (H - 1)
    ccol, crow = (px + 1) / sx, (py + 1) / sy
    TW, TH, NX, NY = 32, 8, 6, 5
    xoff = cheb(NX, 0.0, TW - 1.0)
    yoff = cheb(NY, 0.0, TH - 1.0)
    Lx = lagrange(xoff, np.arange(TW, dtype=np.float64))
    Ly = lagrange(yoff, np.arange(TH, dtype=np.float64))
    sw = ST * TW
    sxoff = cheb(NX2, 0.0, sw - 1.0)                                    # super-nodes of the 512-px interval
    # weights super-node -> tile node x, per tile of the super-tile: [ST][NX][NX2]
    W2 = np.stack([lagrange(sxoff, t * TW + xoff).T for t in range(ST)])
    print('Lebesgue constant of the super-interpolation at the tile nodes: %.2f' % np.abs(W2).sum(-1).max())
    worst = 0.0
    nty, nsx = H // TH, (W + sw - 1) // sw
    rows = sorted(set([int(rng.integers(0, nty)) for _ in range(strips)] + [min(max(int(r // TH) + d, 0), nty - 1) for r in crow for d in (-3, -1, 0, 1, 3)]))
    for ty in rows:
        row0 = ty * TH
        yn = -1 + sy * (row0 + yoff)
        row_near = (crow > row0 - near_y) & (crow < row0 + TH - 1 + near_y)
        for s in range(nsx):
            c0 = s * sw
            sn = (crow > row0 - snear_y) & (crow < row0 + TH - 1 + snear_y) & (ccol > c0 - snear_x) & (ccol < c0 + sw - 1 + snear_x)
            ff = ~sn
            xs2 = -1 + sx * (c0 + sxoff)
            d2 = (xs2[None, :, None] - px[ff][None, None, :]) ** 2 + (yn[:, None, None] - py[ff][None, None, :]) ** 2     # [NY, NX2, k]
            F2x, F2y = (phi(d2) * cx[ff]).sum(-1), (phi(d2) * cy[ff]).sum(-1)
            for t in range(ST):
                col0 = c0 + t * TW
                if col0 >= W:
                    break
                near = row_near & (ccol > col0 - near_x) & (ccol < col0 + TW - 1 + near_x)
                mid = sn & ~near
                xn = -1 + sx * (col0 + xoff)
                Fx, Fy = F2x @ W2[t].T, F2y @ W2[t].T                                                                    # [NY, NX]
                if mid.any():
                    d2m = (xn[None, :, None] - px[mid][None, None, :]) ** 2 + (yn[:, None, None] - py[mid][None, None, :]) ** 2
                    Fx, Fy = Fx + (phi(d2m) * cx[mid]).sum(-1), Fy + (phi(d2m) * cy[mid]).sum(-1)
                xs = -1 + sx * (col0 + np.arange(TW))
                ys = -1 + sy * (row0 + np.arange(TH))
                far = ~near
                d2e = (xs[None, :, None] - px[far][None, None, :]) ** 2 + (ys[:, None, None] - py[far][None, None, :]) ** 2
                ex, ey = (phi(d2e) * cx[far]).sum(-1), (phi(d2e) * cy[far]).sum(-1)
                ax, ay = Ly.T @ Fx @ Lx, Ly.T @ Fy @ Lx
                worst = max(worst, np.abs(ax - ex).max(), np.abs(ay - ey).max())
    return worst


if __name__ == '__main__':
    for (H, W, m, amp) in ((2160, 3840, 16, 0.1), (1080, 1920, 16, 0.1), (2160, 3840, 16, 0.3)):
        for nx2, snx, sny in ((16, 128.0, 160.0), (16, 96.0, 128.0)):
            print(H, W, m, amp, 'NX2', nx2, 'super-near', snx, sny, 'worst |two-level - exact far field| = %.2e' % run(H, W, m, amp, NX2=nx2, snear_x=snx, snear_y=sny), flush=True)
