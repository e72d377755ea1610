"""CPU: the NumPy oracle against golden vectors produced by the UNMODIFIED reference sources
(executed over oracle/tf1_shim.py -- see tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import dvsg_oracle as O
from conftest import load_golden

TPS_CASES = ['tps_4x4', 'tps_5x5', 'tps_4x4_big', 'tps_resize', 'tps2_4x4', 'tps_8x8']
# fp32 noise of the coefficient solve grows with the system's condition number
COORD_TOL = {'tps_8x8': 1e-4}


@pytest.mark.parametrize('name', TPS_CASES)
def test_tps_forward(name):
    g = load_golden(name)
    fn = O.thin_plate_spline if int(g['variant']) == 1 else O.thin_plate_spline2
    out, x, y = fn(g['u'], g['coord'], g['second'], g['out_size'])
    tol = COORD_TOL.get(name, 2e-5)
    assert out.shape == g['out'].shape and out.dtype == np.float32
    assert np.abs(x - g['x']).max() <= tol
    assert np.abs(y - g['y']).max() <= tol
    # frames are tiny (<= 52 px wide), so coordinate noise maps to << 1e-4 in pixel values
    assert np.abs(out - g['out']).max() <= 1e-4
    # both fp32 evaluations sit equally close to the fp64 run of the same reference graph
    assert np.abs(x - g['x64']).max() <= 2 * max(np.abs(g['x'] - g['x64']).max(), 1e-6)


@pytest.mark.parametrize('name', TPS_CASES)
def test_tps_sampler_stage_is_bit_exact_on_identical_coordinates(name):
    """A4 given the golden x, y: integer corners, weights and blend are pure IEEE ops."""
    g = load_golden(name)
    oh, ow = (int(v) for v in g['out_size'])
    out = O.tps_interpolate(g['u'], g['x'], g['y'], oh, ow).reshape(g['out'].shape)
    np.testing.assert_array_equal(out, g['out'])


@pytest.mark.parametrize('name', ['tps_4x4', 'tps_5x5', 'tps_4x4_big', 'tps_resize', 'tps_8x8'])
def test_tps_backward_stages(name):
    """A5 stage by stage on the golden coordinates (gradients are discontinuous across
    sampling-cell boundaries, so they are compared on identical x, y)."""
    g = load_golden(name)
    oh, ow = (int(v) for v in g['out_size'])
    g_im, gx, gy = O.tps_interpolate_bwd(g['u'], g['x'], g['y'], oh, ow, g['g_out'])
    scale = np.abs(g['grad_u']).max()
    assert np.abs(g_im - g['grad_u']).max() <= 1e-5 * scale + 1e-6
    g_t = O.tps_grid_bwd(g['coord'], oh, ow, gx + g['g_x'], gy + g['g_y'])
    _, w_inv = O.tps_solve(g['coord'], g['coord'] + g['second'], return_inverse=True)
    g_vec = O.tps_solve_bwd(w_inv, g_t)
    ref = g['grad_second']
    assert np.abs(g_vec - ref).max() <= 1e-4 * np.abs(ref).max()


def test_tps_mask_all_ones_image():
    g = load_golden('tps_mask')
    out, _, _ = O.thin_plate_spline(g['u'], g['coord'], g['second'], g['out_size'])
    assert np.abs(out - g['out']).max() <= 1e-6
    assert out.max() <= 1.0 + 1e-6 and out.min() >= -1e-5


def test_st_meshgrid_exact():
    g = load_golden('meshgrid')
    np.testing.assert_array_equal(O.st_meshgrid(g['out_size']), g['grid'])
    np.testing.assert_array_equal(O.st_meshgrid(g['out_size2'])[::97], g['grid2'])


@pytest.mark.parametrize('name', ['bilinear_c18', 'bilinear_c3'])
def test_bilinear_interp(name):
    g = load_golden(name)
    out = O.bilinear_interp(g['im'], g['x'], g['y'], g['out_size'])
    np.testing.assert_array_equal(out, g['out'])
    g_im, gx, gy = O.bilinear_interp_bwd(g['im'], g['x'], g['y'], g['out_size'], g['g_out'])
    assert np.abs(g_im - g['grad_im']).max() <= 2e-6
    assert np.abs(gx - g['grad_x']).max() <= 1e-6 * np.abs(g['grad_x']).max() + 1e-6
    assert np.abs(gy - g['grad_y']).max() <= 1e-6 * np.abs(g['grad_y']).max() + 1e-6


def test_projective_and_affine_transformers():
    g = load_golden('projective')
    assert np.abs(O.projective_transform(g['im'], g['theta'], g['out_size']) - g['out']).max() <= 1e-5
    g = load_golden('affine')
    assert np.abs(O.affine_transform(g['im'], g['theta'], g['out_size']) - g['out']).max() <= 1e-5


@pytest.mark.parametrize('name', ['flow_small', 'flow_large'])
def test_tf_warp(name):
    g = load_golden(name)
    h, w = g['im'].shape[1:3]
    np.testing.assert_array_equal(O.tf_warp(g['im'], g['flow'], h, w), g['out'])
    g_im, g_flow = O.tf_warp_bwd(g['im'], g['flow'], h, w, g['g_out'])
    assert np.abs(g_im - g['grad_im']).max() <= 2e-6
    assert np.abs(g_flow - g['grad_flow']).max() <= 2e-6


# ---- N3: the loss consumers of the warp outputs, pinned by trainer.py's own method bodies (make_golden.loss_cases) ----
def test_masked_mse_vs_reference_golden():
    g = load_golden('loss_masked_mse')
    loss, sq, ms = O.masked_mse(g['pred'], g['gt'], g['mask'])
    assert abs(loss - float(g['loss'])) <= 1e-6 * abs(float(g['loss']))
    assert ms[1] == 0.0 and sq[1] == 0.0            # the fully masked frame: div_no_nan -> 0


def test_temporal_loss_vs_reference_golden():
    g = load_golden('loss_temporal')
    h, w = g['pred'].shape[1:3]
    pw = O.tf_warp(g['pred'], g['flow'], h, w)
    mw = O.tf_warp(g['mask_pred'], g['flow'], h, w)
    loss, _, _ = O.masked_mse(pw, g['gt'], (mw * g['mask_gt']).astype(np.float32))
    assert abs(loss - float(g['loss'])) <= 1e-6 * abs(float(g['loss']))


def test_surf_loss_vs_reference_golden():
    g = load_golden('loss_surf')
    h, w = (int(v) for v in g['hw'])
    b = g['surf'].shape[0]
    loss, idx = O.surf_loss(g['surf'], g['x'], g['y'], g['max_dim'], b, w, h)
    assert abs(loss - float(g['loss'])) <= 1e-6 * abs(float(g['loss']))
    assert idx.max() == h * w                        # padded features hit the appended -1 entry (trainer.py:364-365)


# ---- N4: ElasticTransformer (spatial_transformer.py:93-362 executed unmodified over the shim) ----
@pytest.mark.parametrize('name', ['elastic_4x4', 'elastic_3x3_resize'])
def test_elastic_transformer_vs_reference_golden(name):
    g = load_golden(name)
    gs, osz = int(g['grid_size']), tuple(int(v) for v in g['out_size'])
    out, x, y = O.elastic_transform(g['im'], g['theta'], gs, osz)
    assert max(np.abs(x - g['x']).max(), np.abs(y - g['y']).max()) <= 2e-5
    assert np.abs(out - g['out']).max() <= 1e-4
    _, _, _, l_inv = O.elastic_initialize(gs, osz)
    assert np.abs(l_inv - g['l_inv']).max() <= 1e-5 * np.abs(g['l_inv']).max()
    # the sampler on the golden's own coordinates is bit-exact (bilinear_interp, B2)
    np.testing.assert_array_equal(O.bilinear_interp(g['im'], g['x'], g['y'], osz).reshape(g['out'].shape), g['out'])


# ---- backward of the projective / affine transformers (make_golden.transformer_grad_cases) ----
@pytest.mark.parametrize('name', ['projective_grad', 'projective_grad_c18', 'affine_grad'])
def test_transformer_backward_vs_reference_golden(name):
    """The reference's transform() with autograd over its own op sequence: grid-stage gradient rel <= 1e-5 (fp64 oracle
    against the fp32 graph), end-to-end gradients w.r.t. theta and the input rel <= 1e-4 of the max-norm."""
    g = load_golden(name)
    proj = name.startswith('projective')
    osz = tuple(int(v) for v in g['out_size'])
    xs, ys = (O.projective_grid if proj else O.affine_grid)(g['theta'], osz)
    assert max(np.abs(xs - g['x']).max(), np.abs(ys - g['y']).max()) <= 2e-6
    gt = O.homography_grid_bwd(g['theta'], osz, g['g_x'], g['g_y'], proj)
    assert gt.shape == g['grad_theta_grid'].shape
    assert np.abs(gt - g['grad_theta_grid']).max() <= 1e-5 * np.abs(g['grad_theta_grid']).max()
    g_im, g_theta = O.homography_transform_bwd(g['im'], g['theta'], osz, g['g_out'], proj)
    assert np.abs(g_im - g['grad_im']).max() <= 1e-4 * np.abs(g['grad_im']).max()
    assert np.abs(g_theta - g['grad_theta']).max() <= 1e-4 * np.abs(g['grad_theta']).max()


# ---- gradient w.r.t. the control-point positions (make_golden.coord_grad_cases) ----
def _oracle_coord_grad(g):
    osz = tuple(int(v) for v in g['out_size'])
    coord = g['coord']
    if int(g['variant']) == 1:
        _, g_second, g_coord = O.thin_plate_spline_bwd(g['u'], coord, g['second'], osz, g['g_out'], g['g_x'], g['g_y'], want_coord=True)
        return g_second, g_coord
    t, w_inv = O.tps_solve(coord, g['second'], return_inverse=True)                  # ThinPlateSpline2: the targets are a separate input
    x, y = O.tps_grid(t, coord, osz[0], osz[1])
    _, gx, gy = O.tps_interpolate_bwd(g['u'], x, y, osz[0], osz[1], g['g_out'])
    gx, gy = gx + g['g_x'], gy + g['g_y']
    g_t = O.tps_grid_bwd(coord, osz[0], osz[1], gx, gy)
    return O.tps_solve_bwd(w_inv, g_t), O.tps_coord_bwd(coord, t, w_inv, g_t, osz[0], osz[1], gx, gy)


@pytest.mark.parametrize('name', ['tps_coord_grad', 'tps2_coord_grad'])
def test_tps_coord_gradient_vs_reference_golden(name):
    """ThinPlateSpline / ThinPlateSpline2 executed unmodified with coord.requires_grad: the oracle's restatement of that
    gradient (grid radial terms + system matrix + right-hand side) is within 2e-5 of the fp64 run of the reference graph
    and 1e-4 of its fp32 run (max-norm relative)."""
    g = load_golden(name)
    g_second, g_coord = _oracle_coord_grad(g)
    assert np.abs(g_coord - g['grad_coord64']).max() <= 2e-5 * np.abs(g['grad_coord64']).max()
    assert np.abs(g_coord - g['grad_coord']).max() <= 1e-4 * np.abs(g['grad_coord']).max()
    assert np.abs(g_second - g['grad_second64']).max() <= 2e-5 * np.abs(g['grad_second64']).max()
