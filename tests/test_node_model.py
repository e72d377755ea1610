"""CPU: the APPROXIMATION error of the tile-node TPS evaluation, isolated from fp32 rounding -- an fp64 model of exactly what
the tile kernels do (csrc/tile_common.cuh tile_node_coords) using the very tables the kernels are compiled with
(csrc/node_tables.cuh): far field at the 6 x 5 Chebyshev nodes of a 32 x 8 tile + tensor-product Lagrange interpolation,
control points within the tile box grown by 48 x 24 px evaluated per pixel.  Bar: <= 2e-7 normalised units on every HD shape
for which the kernels select the node evaluation (dense meshes on small frames: stated separately) -- an order of magnitude below the fp32 noise of the reference's own sum."""
import os
import re

import numpy as np
import pytest

from oracle import dvsg_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TC, TR, NEAR_X, NEAR_Y = 32, 8, 48.0, 24.0


def load_tables():
    src = open(os.path.join(ROOT, 'coupe', 'dvsg_b200', 'csrc', 'node_tables.cuh')).read()

    def arr(name):
        body = src[src.index(name):]
        body = body[body.index('=') + 1:body.index(';')]
        return np.array([float(v.rstrip('f')) for v in re.findall(r'-?\d+\.\d+e[+-]\d+f|0\.0f', body)])
    xoff, yoff = arr('NODE_XOFF[NNX]'), arr('NODE_YOFF[NNY]')
    lx = arr('NODE_LX[32][8]').reshape(32, 8)[:, :len(xoff)]          # [column][node]
    my = arr('NODE_MY[NNY][8]').reshape(len(yoff), 8)                 # [node][row]
    return xoff, yoff, lx, my


def model_error(H, W, m, amp, seed=0, n_random=300):
    xoff, yoff, lx, my = load_tables()
    rng = np.random.default_rng(seed)
    coord = O.regular_mesh(m, m).astype(np.float64)[None]
    T = O.tps_solve(coord, coord + rng.uniform(-amp, amp, coord.shape), dtype=np.float64)[0]
    cx, cy, px, py = T[0, 3:], T[1, 3:], coord[0, :, 0], coord[0, :, 1]
    sx, sy = 2.0 / (W - 1), 2.0 / (H - 1)
    ccol, crow = (px + 1) / sx, (py + 1) / sy
    phi = lambda d2: d2 * np.log(d2 + 1e-6)
    ntx, nty = W // TC, H // TR
    tiles = {(int(rng.integers(0, ntx)), int(rng.integers(0, nty))) for _ in range(n_random)}
    for k in range(m * m):          # and every tile around every control point
        for dx in (-2, -1, 0, 1, 2):
            for dy in range(-5, 6):
                tiles.add((min(max(int(ccol[k] // TC) + dx, 0), ntx - 1), min(max(int(crow[k] // TR) + dy, 0), nty - 1)))
    worst = 0.0
    for tx, ty in tiles:
        col0, row0 = tx * TC, ty * TR
        near = (ccol > col0 - NEAR_X) & (ccol < col0 + TC - 1 + NEAR_X) & (crow > row0 - NEAR_Y) & (crow < row0 + TR - 1 + NEAR_Y)
        X, Y = np.meshgrid(-1 + sx * (col0 + np.arange(TC)), -1 + sy * (row0 + np.arange(TR)))
        P = phi((X[..., None] - px) ** 2 + (Y[..., None] - py) ** 2)
        XN, YN = np.meshgrid(-1 + sx * (col0 + xoff), -1 + sy * (row0 + yoff))
        d2n = (XN[..., None] - px[~near]) ** 2 + (YN[..., None] - py[~near]) ** 2
        Pn = d2n * np.log(d2n) + 1e-6          # the far field folds the epsilon into a constant (tps_affine0)
        for c in (cx, cy):
            approx = my.T @ (Pn @ c[~near]) @ lx.T + P[..., near] @ c[near]
            worst = max(worst, np.abs(approx - P @ c).max())
    return worst


@pytest.mark.parametrize('case', [(540, 960, 4, 0.1), (540, 960, 5, 0.3), (720, 1280, 4, 0.1), (720, 1280, 5, 0.3), (1080, 1920, 4, 0.1),
                                  (1080, 1920, 5, 0.1), (1080, 1920, 8, 0.05), (2160, 3840, 16, 0.02)],
                         ids=lambda c: '%dx%d_m%d_a%g' % c)
def test_node_evaluation_approximation_error(case):
    err = model_error(*case)
    print('%s: approximation error %.2e' % (case, err))
    assert err <= 2e-7


def test_node_evaluation_on_the_densest_supported_case():
    """16 x 16 mesh on 288 x 512 (12 near control points per tile, coefficient mass sum|c| = 107): the error scales with
    the coefficient mass; bar 1e-6 = a tenth of the fp32 noise of that sum (1.1e-5, test_gpu_forward node gate)."""
    assert model_error(288, 512, 16, 0.02, n_random=150) <= 1e-6
    # 8 x 8 mesh on the smallest frame the node evaluation is selected for (sum|c| = 12): 2.3e-7
    assert model_error(200, 400, 8, 0.05, n_random=150) <= 5e-7


def test_lagrange_tables_are_a_partition_of_unity():
    xoff, yoff, lx, my = load_tables()
    assert lx.shape == (32, 6) and my.shape == (5, 8)
    assert np.abs(lx.sum(1) - 1).max() <= 1e-6 and np.abs(my.sum(0) - 1).max() <= 1e-6
    # reproduces polynomials up to the node count's degree: x^5 along the columns, y^4 along the rows
    c = np.arange(32.0)
    assert np.abs(lx @ xoff ** 5 - c ** 5).max() <= 1e-6 * 31 ** 5
    r = np.arange(8.0)
    assert np.abs(my.T @ yoff ** 4 - r ** 4).max() <= 1e-6 * 7 ** 4


# ---- second level (16 x 16 meshes): csrc/tile_common.cuh, G == TKS branch of tile_node_coords / tile_node_tables ----
SN_TILES, SN_NX, SN_NEAR_X, SN_NEAR_Y = 16, 16, 128.0, 160.0


def load_super_tables():
    src = open(os.path.join(ROOT, 'coupe', 'dvsg_b200', 'csrc', 'node_tables.cuh')).read()

    def arr(name):
        body = src[src.index(name):]
        body = body[body.index('=') + 1:body.index(';')]
        return np.array([float(v.rstrip('f')) for v in re.findall(r'-?\d+\.\d+e[+-]\d+f|0\.0f', body)])
    sxoff = arr('SNODE_XOFF[SNODE_NX]')
    w = arr('SNODE_W[SNODE_TILES][NNX][SNODE_NX]').reshape(SN_TILES, 6, SN_NX)      # [tile][x node][super node]
    return sxoff, w


def two_level_error(H, W, m, amp, seed=0, n_strips=10):
    """fp64 model of the two-level evaluation: far-far field (everything outside the super-tile box grown by 128 x 160 px) on
    16 x 5 super-nodes, interpolated in x to the tile nodes with SNODE_W; super-near control points that are not tile-near at
    the tile nodes; tile-near ones per pixel.  Returns the worst distance to the exact sum."""
    xoff, yoff, lx, my = load_tables()
    sxoff, w2 = load_super_tables()
    rng = np.random.default_rng(seed)
    coord = O.regular_mesh(m, m).astype(np.float64)[None]
    T = O.tps_solve(coord, coord + rng.uniform(-amp, amp, coord.shape), dtype=np.float64)[0]
    cx, cy, px, py = T[0, 3:], T[1, 3:], coord[0, :, 0], coord[0, :, 1]
    sx, sy = 2.0 / (W - 1), 2.0 / (H - 1)
    ccol, crow = (px + 1) / sx, (py + 1) / sy
    phi = lambda d2: d2 * np.log(d2 + 1e-6)
    far = lambda d2: d2 * np.log(d2) + 1e-6
    nty, sw = H // TR, SN_TILES * TC
    strips = sorted({int(rng.integers(0, nty)) for _ in range(n_strips)} | {min(max(int(r // TR) + d, 0), nty - 1) for r in crow[::m] for d in (-2, 0, 3)})
    worst, max_sn = 0.0, 0
    for ty in strips:
        row0 = ty * TR
        yn = -1 + sy * (row0 + yoff)
        rn = (crow > row0 - NEAR_Y) & (crow < row0 + TR - 1 + NEAR_Y)
        for c0 in range(0, W, sw):
            sn = (crow > row0 - SN_NEAR_Y) & (crow < row0 + TR - 1 + SN_NEAR_Y) & (ccol > c0 - SN_NEAR_X) & (ccol < c0 + sw - 1 + SN_NEAR_X)
            max_sn = max(max_sn, int(sn.sum()))
            xs2 = -1 + sx * (c0 + sxoff)
            d2 = (xs2[None, :, None] - px[~sn]) ** 2 + (yn[:, None, None] - py[~sn]) ** 2      # [y node, super node, k]
            f2 = [far(d2) @ c[~sn] for c in (cx, cy)]
            for t in range(SN_TILES):
                col0 = c0 + t * TC
                if col0 + TC > W:
                    break
                near = rn & (ccol > col0 - NEAR_X) & (ccol < col0 + TC - 1 + NEAR_X)
                mid = sn & ~near
                X, Y = np.meshgrid(-1 + sx * (col0 + np.arange(TC)), -1 + sy * (row0 + np.arange(TR)))
                P = phi((X[..., None] - px) ** 2 + (Y[..., None] - py) ** 2)
                XN, YN = np.meshgrid(-1 + sx * (col0 + xoff), yn)
                Pm = far((XN[..., None] - px[mid]) ** 2 + (YN[..., None] - py[mid]) ** 2)
                for c, f in zip((cx, cy), f2):
                    fn = f @ w2[t].T + Pm @ c[mid]                                             # [y node, x node]
                    approx = my.T @ fn @ lx.T + P[..., near] @ c[near]
                    worst = max(worst, np.abs(approx - P @ c).max())
    return worst, max_sn


@pytest.mark.parametrize('case', [(2160, 3840, 16, 0.1), (1080, 1920, 16, 0.05)], ids=lambda c: '%dx%d_m%d_a%g' % c)
def test_two_level_node_evaluation_approximation_error(case):
    """Bar: <= 2e-7 like the single level (measured 1e-8 at 4K, 4e-8 at 1080p), with at most 32 super-near control points
    per super-tile at these shapes (the kernels fall back to the single level beyond that)."""
    err, max_sn = two_level_error(*case)
    print('%s: two-level approximation error %.2e, at most %d super-near control points' % (case, err, max_sn))
    assert err <= 2e-7 and max_sn <= 32


def test_super_tile_weights_are_a_partition_of_unity_and_reproduce_polynomials():
    xoff, _, _, _ = load_tables()
    sxoff, w2 = load_super_tables()
    assert sxoff.shape == (16,) and w2.shape == (16, 6, 16)
    assert np.abs(w2.sum(-1) - 1).max() <= 1e-6
    for t in range(16):
        x = t * TC + xoff
        assert np.abs(w2[t] @ sxoff ** 7 - x ** 7).max() <= 1e-6 * 511.0 ** 7
