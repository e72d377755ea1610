"""CPU: analytic known-answer tests and independent cross-checks of the oracle
(SURVEY.md section 4).  None of this needs the reference or a GPU."""
import numpy as np
import pytest
import torch

from oracle import dvsg_oracle as O
from helpers import smooth_image, tiled_mesh


def test_tf_linspace_is_not_np_linspace():
    a = O.tf_linspace(-1.0, 1.0, 1920)
    assert a.dtype == np.float32 and a[0] == -1.0
    assert a[-1] != np.float32(1.0)              # TF 1.x formula does not pin the last element
    assert np.abs(a - np.linspace(-1, 1, 1920)).max() < 2.5e-7


@pytest.mark.parametrize('n', [3, 4, 5])
def test_tps_identity_and_affine_reproduction(n):
    coord = tiled_mesh(n, n, 1)
    t = O.tps_solve(coord, coord)
    expect = np.zeros_like(t)
    expect[0, 0, 1] = 1.0
    expect[0, 1, 2] = 1.0
    assert np.abs(t - expect).max() < 5e-6
    a = np.array([[1.1, 0.2], [-0.15, 0.9]], np.float32)
    off = np.array([0.05, -0.07], np.float32)
    t = O.tps_solve(coord, coord @ a.T + off)
    assert np.abs(t[0, :, 3:]).max() < 5e-6      # rbf weights vanish for an affine map
    assert np.abs(t[0, :, 0] - off).max() < 5e-6
    assert np.abs(t[0, :, 1:3] - a).max() < 5e-6


def test_tps_sampler_integer_coordinates_return_source_pixels():
    rng = np.random.default_rng(1)
    h, w = 6, 8
    im = rng.random((1, h, w, 3), dtype=np.float32)
    jj, ii = np.meshgrid(np.arange(w - 1), np.arange(h - 1))
    # invert x_pix = (x+1)*W/2 for integer pixel targets that are exactly representable
    x = (2.0 * jj.reshape(-1) / w - 1.0).astype(np.float32)
    y = (2.0 * ii.reshape(-1) / h - 1.0).astype(np.float32)
    xp, yp, *_ = O.tps_sample_indices(x, y, h, w)
    keep = (xp == np.round(xp)) & (yp == np.round(yp))
    out = np.zeros((h * w, 3), np.float32)
    res = O.tps_interpolate(np.tile(im, (1, 1, 1, 1)), np.resize(x, h * w), np.resize(y, h * w), h, w)
    src = im[0, np.resize(ii.reshape(-1), h * w), np.resize(jj.reshape(-1), h * w)]
    k = np.resize(keep, h * w)
    assert k.sum() > 10
    np.testing.assert_array_equal(res[k], src[k])
    del out


def test_tps_identity_warp_zeroes_last_row_and_column():
    h, w = 10, 12
    ones = np.ones((1, h, w, 1), np.float32)
    coord = tiled_mesh(4, 4, 1)
    out, x, y = O.thin_plate_spline(ones, coord, np.zeros_like(coord), (h, w))
    # row/col 0 sit exactly on the x_pix = 0 boundary: fp noise in T decides their side
    assert np.abs(out[0, 1:-1, 1:-1, 0] - 1.0).max() < 1e-5
    assert np.abs(out[0, -1]).max() < 1e-5 and np.abs(out[0, :, -1]).max() < 1e-5


def test_tps_sampler_outside_frame_is_about_zero():
    """Clamp-then-weight: outside [0,W-1)x[0,H-1) the paired weights cancel (H5)."""
    rng = np.random.default_rng(2)
    im = rng.random((1, 9, 11, 3), dtype=np.float32)
    x = np.array([-1.5, 1.4, 0.0, 0.0, -3.0], np.float32)
    y = np.array([0.0, 0.0, -1.7, 1.2, 3.0], np.float32)
    out = O.tps_interpolate(im, np.resize(x, 99), np.resize(y, 99), 9, 11)
    assert np.abs(out[:5]).max() < 1e-5


@pytest.mark.parametrize('c', [1, 3, 18])
def test_bilinear_interp_matches_torch_grid_sample(c):
    """Independent implementation of the zero-padded (W-1)/2 sampler: grid_sample with
    align_corners=True, padding_mode='zeros'."""
    rng = np.random.default_rng(3)
    b, h, w, oh, ow = 2, 13, 17, 9, 11
    im = rng.random((b, h, w, c), dtype=np.float32)
    x = rng.uniform(-1.4, 1.4, b * oh * ow).astype(np.float32)
    y = rng.uniform(-1.4, 1.4, b * oh * ow).astype(np.float32)
    out = O.bilinear_interp(im, x, y, (oh, ow)).reshape(b, oh, ow, c)
    grid = torch.from_numpy(np.stack([x, y], -1).reshape(b, oh, ow, 2))
    ref = torch.nn.functional.grid_sample(torch.from_numpy(im).permute(0, 3, 1, 2), grid, mode='bilinear',
                                          padding_mode='zeros', align_corners=True).permute(0, 2, 3, 1).numpy()
    assert np.abs(out - ref).max() < 2e-5


def test_tf_warp_zero_flow_is_identity_and_integer_shift_is_exact():
    rng = np.random.default_rng(4)
    im = rng.random((2, 7, 9, 3), dtype=np.float32)
    np.testing.assert_array_equal(O.tf_warp(im, np.zeros((2, 7, 9, 2), np.float32), 7, 9), im)
    flow = np.zeros((2, 7, 9, 2), np.float32)
    flow[..., 0] = 2.0
    flow[..., 1] = -1.0
    out = O.tf_warp(im, flow, 7, 9)
    np.testing.assert_array_equal(out[:, 1:, :-2], im[:, :-1, 2:])
    assert np.abs(out[:, 0]).max() == 0 and np.abs(out[:, :, -2:]).max() == 0


def _fd(fn, v, eps):
    g = np.zeros(v.shape)
    for i in np.ndindex(*v.shape):          # v may be a non-contiguous view: index in place
        old = v[i]
        v[i] = old + eps
        hi = fn()
        v[i] = old - eps
        lo = fn()
        v[i] = old
        g[i] = (hi - lo) / (2 * eps)
    return g


def test_tps_backward_matches_fp64_finite_differences():
    rng = np.random.default_rng(5)
    b, h, w, c = 1, 10, 12, 2
    u = smooth_image(rng, b, h, w, c).astype(np.float64)
    coord = tiled_mesh(3, 3, b).astype(np.float64)
    vec = rng.uniform(-0.08, 0.08, coord.shape)
    g_out = rng.standard_normal((b, h, w, c))
    g_x = rng.standard_normal(b * h * w) * 0.1
    g_y = rng.standard_normal(b * h * w) * 0.1

    def loss():
        out, x, y = O.thin_plate_spline(u, coord, vec, (h, w), dtype=np.float64)
        return (out * g_out).sum() + (x * g_x).sum() + (y * g_y).sum()

    g_u, g_v = O.thin_plate_spline_bwd(u, coord, vec, (h, w), g_out, g_x, g_y, dtype=np.float64)
    fd_v = _fd(loss, vec, 1e-6)
    assert np.abs(fd_v - g_v).max() <= 1e-4 * np.abs(fd_v).max()
    sub = u[:, 2:5, 3:6]                                  # a view: perturbs u in place
    fd_u = _fd(loss, sub, 1e-6)
    assert np.abs(fd_u - g_u[:, 2:5, 3:6]).max() <= 1e-6 * max(1.0, np.abs(fd_u).max())


def test_flow_warp_backward_matches_fp64_finite_differences():
    rng = np.random.default_rng(6)
    b, h, w, c = 1, 6, 7, 2
    im = rng.random((b, h, w, c))
    flow = rng.uniform(-2.5, 2.5, (b, h, w, 2))
    flow = np.where(np.abs(flow - np.round(flow)) < 1e-3, flow + 0.01, flow)
    g_out = rng.standard_normal((b, h, w, c))

    def loss():
        return (O.tf_warp(im, flow, h, w, dtype=np.float64) * g_out).sum()

    g_im, g_flow = O.tf_warp_bwd(im, flow, h, w, g_out, dtype=np.float64)
    assert np.abs(_fd(loss, im, 1e-6) - g_im).max() < 1e-6
    assert np.abs(_fd(loss, flow, 1e-7) - g_flow).max() < 1e-5


# ---- N4: frame ingest / egress -------------------------------------------------------------------
def test_frame_io_oracle_matches_cv2_and_numpy_casts():
    """eval.py:76-81,112-113 are plain cv2 / NumPy calls: the oracle's restatement is compared with those very calls
    (cv2.cvtColor for the channel order, np.uint8() for the cast), and the device formula for u/255 -- q = u*c,
    r = fma(-255, q, u), q + r*c -- is checked for all 256 inputs in exact arithmetic."""
    import cv2
    rng = np.random.default_rng(0)
    f8 = rng.integers(0, 256, (2, 9, 11, 3), dtype=np.uint8)
    for i in range(2):
        ref = (cv2.cvtColor(f8[i], cv2.COLOR_BGR2RGB) / 255.).astype(np.float32)
        np.testing.assert_array_equal(O.frames_u8_to_f32(f8[i], True), ref)
    x = rng.uniform(0, 1, (9, 11, 3)).astype(np.float32)
    ref8 = cv2.cvtColor(np.uint8(x * 255.), cv2.COLOR_RGB2BGR)
    np.testing.assert_array_equal(O.frames_f32_to_u8(x, True), ref8)
    np.testing.assert_array_equal(O.frames_f32_to_u8(O.frames_u8_to_f32(f8, True), True), f8)   # round trip is the identity
    from fractions import Fraction

    def rn(fr):     # exact rational -> nearest float32
        return np.float32(float(fr)) if fr == 0 else np.float32(np.float64(fr.numerator) / np.float64(fr.denominator))
    c = np.float32(1.0 / 255.0)
    for u in range(256):
        q = np.float32(np.float32(u) * c)
        r = rn(Fraction(u) - 255 * Fraction(float(q)))
        q2 = rn(Fraction(float(q)) + Fraction(float(r)) * Fraction(float(c)))
        assert q2 == np.float32(u / 255.), u


# ---- N3: masked MSE and surf loss ------------------------------------------------------------------
def test_masked_mse_and_surf_loss_known_answers():
    pred = np.full((2, 4, 4, 3), 0.75, np.float32)
    gt = np.full((2, 4, 4, 3), 0.25, np.float32)
    mask = np.zeros((2, 4, 4, 3), np.float32)
    mask[0, :2] = 1.0                                    # 24 active elements in frame 0, none in frame 1
    loss, sq, ms = O.masked_mse(pred, gt, mask)
    assert sq[0] == 24 * 0.25 and ms[0] == 24 and sq[1] == 0 and ms[1] == 0
    assert loss == (0.25 + 0.0) / 2                      # div_no_nan: the empty frame contributes 0
    # identity spline: the grid holds the normalised pixel coordinates, so features that did not move give zero loss
    h, w, b = 6, 8, 1
    xs = np.tile(O.tf_linspace(-1, 1, w)[None], (h, 1)).reshape(-1)
    ys = np.tile(O.tf_linspace(-1, 1, h)[:, None], (1, w)).reshape(-1)
    surf = np.zeros((1, 2, 3, 2), np.int32)
    surf[0, 0] = surf[0, 1] = [[1, 2], [7, 5], [0, 0]]
    loss, idx = O.surf_loss(surf, xs, ys, np.array([3.0]), b, w, h)
    assert list(idx[0]) == [1 + 2 * w, 7 + 5 * w, 0] and loss < 1e-12
    surf[0, 0, 0] = [3, 2]                               # unstable feature two pixels to the right
    loss, _ = O.surf_loss(surf, xs, ys, np.array([3.0]), b, w, h)
    assert abs(loss - (2 * 2.0 / (w - 1)) ** 2 / 3.0) < 1e-6


def test_resize_restatement_matches_cv2():
    """eval.py:80 calls cv2.resize on the float64 frame: the oracle's restatement of OpenCV's INTER_LINEAR agrees with
    cv2 itself to 1e-12 in float64 (the last double bits depend on OpenCV's own summation / FMA use), hence to at most
    one fp32 ulp -- in practice bit for bit -- after the fp32 feed cast, for integer and non-integer ratios, up- and
    down-scaling, and its impulse responses (the interpolation weights) are cv2's."""
    import cv2
    rng = np.random.default_rng(1)
    for hs, ws, h, w in [(720, 1280, 288, 512), (480, 640, 288, 512), (100, 150, 288, 512), (288, 512, 288, 512), (37, 53, 61, 19),
                         (719, 1279, 288, 512), (2, 2, 5, 7), (3, 5, 3, 5)]:
        f8 = rng.integers(0, 256, (hs, ws, 3), dtype=np.uint8)
        ref = cv2.resize(cv2.cvtColor(f8, cv2.COLOR_BGR2RGB) / 255., (w, h))
        got = O.resize_linear_f64(f8[..., ::-1] / 255., w, h)
        assert np.abs(ref - got).max() <= 1e-12
        mine, ref32 = O.read_frame(f8, w, h), ref.astype(np.float32)
        assert np.abs(mine - ref32).max() <= 1.2e-7 and (mine != ref32).mean() <= 1e-3
    imp = np.zeros((2, 9, 1))
    imp[:, 4, 0] = 1.0
    np.testing.assert_allclose(O.resize_linear_f64(imp, 7, 2)[..., 0], cv2.resize(imp, (7, 2)), atol=1e-15)


@pytest.mark.parametrize('sizes', [(72, 128, 288, 512), (288, 512, 288, 512), (180, 320, 288, 512), (360, 640, 288, 512), (100, 150, 61, 19), (1, 1, 5, 7)])
def test_flow_ingest_restatement_matches_cv2(sizes):
    """data_loader.py:239 is plain cv2 + NumPy: the oracle's restatement of cv2.resize on a float32 field is compared with
    cv2 itself (bit-identical with cv2 4.13), then `* [w, h]` and the float32 feed."""
    cv2 = pytest.importorskip('cv2')
    hs, ws, h, w = sizes
    rng = np.random.default_rng(hs + w)
    flow = rng.uniform(-0.05, 0.05, (hs, ws, 2)).astype(np.float32)
    ref = (cv2.resize(flow, (w, h)).reshape(h, w, 2) * [w, h]).astype(np.float32)
    np.testing.assert_array_equal(O.flow_ingest(flow, w, h), ref)
