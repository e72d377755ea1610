"""Shared synthetic-input builders for the tests (NumPy only; no reference access)."""
import numpy as np


def smooth_image(rng, b, h, w, c, period=32.0):
    """Band-limited frames in [0,1]: sums of sinusoids with periods >= `period` px plus
    1% uniform noise (the 1e-4 pixel tolerance is only meaningful on
    frames whose gradient is <~ 0.1 per pixel)."""
    yy, xx = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing='ij')
    im = np.zeros((b, h, w, c), np.float64)
    for bi in range(b):
        for ci in range(c):
            for _ in range(4):
                ang = rng.uniform(0, 2 * np.pi)
                f = rng.uniform(0.2, 1.0) / period
                im[bi, :, :, ci] += np.sin(2 * np.pi * f * (np.cos(ang) * xx + np.sin(ang) * yy) + rng.uniform(0, 6.28))
    im = (im - im.min()) / (im.max() - im.min())
    return (0.99 * im + 0.01 * rng.random(im.shape)).astype(np.float32)


def tiled_mesh(n_rows, n_cols, b):
    from oracle import dvsg_oracle as O
    return np.ascontiguousarray(np.tile(O.regular_mesh(n_rows, n_cols)[None], (b, 1, 1)))


def smooth_flow(rng, b, h, w, amp=8.0, jitter=0.5):
    """Smooth flow field: bilinear upsampling of a 9x16 lattice of U(-amp, amp) px plus
    U(-jitter, jitter) px per-pixel jitter."""
    lat = rng.uniform(-amp, amp, (b, 9, 16, 2))
    ys = np.linspace(0, 8, h)
    xs = np.linspace(0, 15, w)
    y0 = np.minimum(np.floor(ys).astype(int), 7)
    x0 = np.minimum(np.floor(xs).astype(int), 14)
    fy = (ys - y0)[None, :, None, None]
    fx = (xs - x0)[None, None, :, None]
    a = lat[:, y0][:, :, x0]
    b_ = lat[:, y0][:, :, x0 + 1]
    c = lat[:, y0 + 1][:, :, x0]
    d = lat[:, y0 + 1][:, :, x0 + 1]
    f = (1 - fy) * ((1 - fx) * a + fx * b_) + fy * ((1 - fx) * c + fx * d)
    return (f + rng.uniform(-jitter, jitter, f.shape)).astype(np.float32)


# ---- gradient comparisons ----------------------------------------------------------------------------------------
def rel_max(a, b):
    """max-norm relative error max|a-b| / max|b|."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def close_elementwise(a, b, rtol=1e-4, atol_frac=2e-5):
    """Element-wise |a-b| <= rtol*|b| + atol_frac*max|b| (the absolute floor covers entries that are themselves sums of
    cancelling terms).  Returns (ok, worst ratio of error to allowance, index of the worst element)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    allow = rtol * np.abs(b) + atol_frac * max(np.abs(b).max(), 1e-30)
    ratio = np.abs(a - b) / allow
    i = int(np.argmax(ratio))
    return bool(ratio.reshape(-1)[i] <= 1.0), float(ratio.reshape(-1)[i]), i


def tps_flip_correction(O, u, coord, target, out_size, x_a, y_a, x_b, y_b, g_out):
    """Sampling corners are floor() of the coordinates, so two valid fp32 evaluations of the same grid (a, b) disagree
    on the integer corner of the few pixels whose coordinate sits within rounding noise of an integer.
    The warped pixel and grad_image are CONTINUOUS across such a flip (bilinear interpolation is C0), but d out / d x
    is not: it jumps to the neighbouring cell's slope.  This returns (n_flipped, correction) where correction is the
    exact effect of those pixels' slope jumps on grad_target, i.e.  grad_target(b) - grad_target(a) restricted to the
    flipped pixels, chained through the basis and W^-T in float64 -- so that end-to-end gradients can be compared on
    ALL elements instead of only when no corner flipped."""
    u = np.asarray(u)
    oh, ow = int(out_size[0]), int(out_size[1])
    H, W = u.shape[1:3]
    _, _, ax0, _, ay0, _ = O.tps_sample_indices(x_a, y_a, H, W)
    _, _, bx0, _, by0, _ = O.tps_sample_indices(x_b, y_b, H, W)
    flipped = (ax0 != bx0) | (ay0 != by0)
    n = int(flipped.sum())
    pn = coord.shape[1]
    if n == 0:
        return 0, np.zeros((u.shape[0], pn, 2), np.float64)
    _, gxa, gya = O.tps_interpolate_bwd(u, x_a, y_a, oh, ow, g_out)
    _, gxb, gyb = O.tps_interpolate_bwd(u, x_b, y_b, oh, ow, g_out)
    dgx = np.where(flipped, gxb.astype(np.float64) - gxa, 0.0)
    dgy = np.where(flipped, gyb.astype(np.float64) - gya, 0.0)
    g_t = O.tps_grid_bwd(coord.astype(np.float64), oh, ow, dgx, dgy, dtype=np.float64)
    _, w_inv = O.tps_solve(coord.astype(np.float64), np.asarray(target, np.float64), dtype=np.float64, return_inverse=True)
    return n, O.tps_solve_bwd(w_inv, g_t)
