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
