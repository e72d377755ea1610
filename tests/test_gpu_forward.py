"""GPU parity tests, forward path -- every call goes through the C ABI (ctypes) via the
drop-in modules.  Oracle = oracle/dvsg_oracle.py; golden vectors = unmodified reference
sources run over the TF1 shim (tests/golden/make_golden.py).

Stated tolerances (fp32):
  * integer sample corners, bilinear weights and the blend: BIT-EXACT given identical x, y
    (the kernel's own x, y output is fed to the oracle's sampler stage);
  * TPS sampling coordinates x, y (normalised): max-abs <= 2e-5 for meshes up to 5x5 --
    the reference's own fp32 inverse-then-multiply carries ~3e-6 of noise (fp64 run of the
    same graph, tests/golden/*.npz x64), MUFU.LG2 adds ~1e-6;
  * warped pixels: max-abs <= 1e-4 on band-limited [0,1] frames.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import smooth_flow, smooth_image, tiled_mesh
from oracle import dvsg_oracle as O

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'
FORCE_DIRECT = 1
TPS_EXACT = 2      # DVSG_FLAG_TPS_EXACT: every radial term per pixel (the tile kernels' default is the tile-node evaluation)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def run_tps(u, coord, second, out_size, variant=1, flags=0, want_mask=False):
    """Through ops (so the staged/direct kernel can be selected with flags)."""
    from coupe.dvsg_b200 import ops
    U, C_, S = cu(u), cu(coord), cu(second)
    target = C_ + S if variant == 1 else S
    T = ops.tps_solve(C_, target)
    out, x, y, mask = ops.tps_warp_fwd(U, C_, T, tuple(int(v) for v in out_size), want_grid=True, want_mask=want_mask, flags=flags)
    torch.cuda.synchronize()
    res = [out.cpu().numpy(), x.cpu().numpy(), y.cpu().numpy(), T.cpu().numpy()]
    if want_mask:
        res.append(mask.cpu().numpy())
    return res


TPS_GOLDEN = ['tps_4x4', 'tps_5x5', 'tps_4x4_big', 'tps_resize', 'tps2_4x4', 'tps_8x8']


@pytest.mark.parametrize('flags', [0, FORCE_DIRECT], ids=['strip', 'direct'])
@pytest.mark.parametrize('name', TPS_GOLDEN)
def test_tps_forward_vs_golden(name, flags):
    g = load_golden(name)
    out, x, y, T = run_tps(g['u'], g['coord'], g['second'], g['out_size'], int(g['variant']), flags)
    tol = 1e-4 if name == 'tps_8x8' else 2e-5
    ex, ey = np.abs(x - g['x']).max(), np.abs(y - g['y']).max()
    print('%s coord err vs fp32 golden %.2e %.2e, vs fp64 %.2e (golden fp32 vs fp64 %.2e)' %
          (name, ex, ey, np.abs(x - g['x64']).max(), np.abs(g['x'] - g['x64']).max()))
    assert ex <= tol and ey <= tol
    assert np.abs(out - g['out']).max() <= 1e-4
    # the sampler stage (index, weights, blend) is bit-exact on the kernel's own coordinates
    oh, ow = (int(v) for v in g['out_size'])
    ref = O.tps_interpolate(g['u'], x, y, oh, ow).reshape(out.shape)
    np.testing.assert_array_equal(out, ref)


@pytest.mark.parametrize('name', ['tps_4x4', 'tps_5x5', 'tps2_4x4'])
def test_dropin_signature_matches_reference(name):
    """ThinPlateSpline(U, coord, vector, out_size) -> (output, x, y), called as model.py:81 does."""
    from coupe.dvsg_b200.ThinPlateSpline import ThinPlateSpline as stn
    from coupe.dvsg_b200.ThinPlateSpline2 import ThinPlateSpline2
    g = load_golden(name)
    fn = stn if int(g['variant']) == 1 else ThinPlateSpline2
    out, x, y = fn(cu(g['u']), cu(g['coord']), cu(g['second']), [int(v) for v in g['out_size']])
    assert out.shape == g['out'].shape and x.shape == g['x'].shape and y.shape == g['y'].shape
    assert np.abs(out.cpu().numpy() - g['out']).max() <= 1e-4
    out2, x2, y2 = fn(cu(g['u']), cu(g['coord']), cu(g['second']), [int(v) for v in g['out_size']], return_grid=False)
    assert x2 is None and y2 is None
    np.testing.assert_array_equal(out2.cpu().numpy(), out.cpu().numpy())


@pytest.mark.parametrize('n,tol32,tol64', [(4, 5e-6, 5e-6), (5, 5e-6, 1e-5), (8, 1e-4, 2e-4), (16, 2e-3, 1e-2)])
def test_tps_solve_coefficients(n, tol32, tol64):
    """K1 against the fp32 oracle (the reference's arithmetic) and against the fp64 run of the
    same algorithm.  The kernel forms the matrix entries in fp32 exactly like the reference, so it
    solves the reference's system; the distance of BOTH fp32 results to the fp64 run is set by
    the fp32 rounding of the entries times cond(W) (2e-2 .. 4e4 from 4x4 to 16x16),
    which is why the stated tolerance grows with the mesh.  Measured on B200 (|T-fp32 oracle|,
    |T-fp64|, |fp32 oracle-fp64|): 4x4 2.6e-7/2.6e-7/3.0e-7, 5x5 3.1e-7/1.2e-6/1.1e-6,
    8x8 5.7e-6/2.2e-5/1.7e-5, 16x16 2.5e-4/1.5e-3/1.5e-3."""
    from coupe.dvsg_b200 import ops
    rng = np.random.default_rng(n)
    B = 3
    coord = tiled_mesh(n, n, B)
    vec = rng.uniform(-0.1, 0.1, coord.shape).astype(np.float32)
    target = (coord + vec).astype(np.float32)
    T = ops.tps_solve(cu(coord), cu(target)).cpu().numpy()
    T32 = O.tps_solve(coord, target)
    T64 = O.tps_solve(coord.astype(np.float64), target.astype(np.float64), dtype=np.float64)
    e32, e64 = np.abs(T - T32).max(), np.abs(T - T64).max()
    print('mesh %dx%d: |T - fp32 oracle| %.2e, |T - fp64| %.2e, |fp32 oracle - fp64| %.2e' % (n, n, e32, e64, np.abs(T32 - T64).max()))
    assert e32 <= tol32 and e64 <= tol64
    # shared-mesh (batch stride 0) path gives the same coefficients
    Ts = ops.tps_solve(cu(coord[0]).unsqueeze(0).expand(B, -1, -1), cu(target)).cpu().numpy()
    assert np.abs(Ts - T).max() <= 1e-6


def test_prepared_solve_of_a_constant_mesh_matches_the_plain_solve():
    """Shared meshes above 29 control points: the inverse is computed once per mesh (dvsg_tps_prepare) and reused;
    the coefficients equal the plain solve bit for bit, and an in-place change of the mesh invalidates the cache."""
    from coupe.dvsg_b200 import _lib, ops
    lib = _lib.load()
    m, b = 7, 5
    rng = np.random.default_rng(2)
    mesh = cu(tiled_mesh(m, m, 1)[0])
    target = mesh.unsqueeze(0) + cu(rng.uniform(-0.05, 0.05, (b, m * m, 2)).astype(np.float32))

    def plain(mesh_t):
        T = torch.empty((b, 2, m * m + 3), dtype=torch.float32, device=DEV)
        nbytes = lib.dvsg_tps_solve_workspace_bytes(b, m * m, 0)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
        rc = lib.dvsg_tps_solve(mesh_t.data_ptr(), 0, target.data_ptr(), T.data_ptr(), b, m * m, ws.data_ptr(), nbytes,
                                torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        return T

    first = ops.tps_solve(mesh.unsqueeze(0).expand(b, -1, -1), target)
    n0 = _lib.launch_count()
    second = ops.tps_solve(mesh.unsqueeze(0).expand(b, -1, -1), target)
    assert _lib.launch_count() - n0 == 1                      # apply only: the inverse was reused
    assert torch.equal(first, plain(mesh)) and torch.equal(second, first)
    g = torch.rand_like(first)
    assert ops.tps_solve_bwd(mesh.unsqueeze(0).expand(b, -1, -1), g).shape == (b, m * m, 2)
    mesh.mul_(0.9)                                            # in-place update -> new tensor version -> fresh inverse
    third = ops.tps_solve(mesh.unsqueeze(0).expand(b, -1, -1), target)
    assert torch.equal(third, plain(mesh)) and not torch.equal(third, first)


def test_tps_identity_and_affine_known_answers():
    from coupe.dvsg_b200 import ops
    coord = tiled_mesh(4, 4, 1)
    T = ops.tps_solve(cu(coord), cu(coord)).cpu().numpy()
    expect = np.zeros_like(T)
    expect[0, 0, 1] = expect[0, 1, 2] = 1.0
    assert np.abs(T - expect).max() < 1e-6
    a = np.array([[1.1, 0.2], [-0.15, 0.9]], np.float32)
    off = np.array([0.05, -0.07], np.float32)
    T = ops.tps_solve(cu(coord), cu(coord @ a.T + off)).cpu().numpy()
    assert np.abs(T[0, :, 3:]).max() < 1e-6 and np.abs(T[0, :, 0] - off).max() < 1e-6 and np.abs(T[0, :, 1:3] - a).max() < 1e-6


# (1, 44, 52, 3, m): tile kernel with a ragged last strip (4 of 8 rows) and a ragged last tile (20 of 32 columns)
@pytest.mark.parametrize('shape', [(2, 288, 512, 3, 4), (2, 288, 512, 3, 5), (1, 37, 53, 3, 4), (2, 40, 64, 1, 3), (1, 33, 48, 18, 4), (1, 44, 52, 3, 4), (1, 44, 52, 3, 6)])
@pytest.mark.parametrize('flags', [0, FORCE_DIRECT], ids=['strip', 'direct'])
def test_tps_forward_vs_oracle_seeded(shape, flags):
    b, h, w, c, m = shape
    rng = np.random.default_rng(h * 1000 + w + m)
    u = smooth_image(rng, b, h, w, c)
    coord = tiled_mesh(m, m, b)
    vec = rng.uniform(-0.1, 0.1, coord.shape).astype(np.float32)
    out, x, y, T = run_tps(u, coord, vec, (h, w), 1, flags)
    r_out, r_x, r_y = O.thin_plate_spline(u, coord, vec, (h, w))
    ex = max(np.abs(x - r_x).max(), np.abs(y - r_y).max())
    eo = np.abs(out - r_out).max()
    xp, yp, x0, x1, y0, y1 = O.tps_sample_indices(x, y, h, w)
    _, _, rx0, rx1, ry0, ry1 = O.tps_sample_indices(r_x, r_y, h, w)
    flips = float(np.mean((x0 != rx0) | (y0 != ry0) | (x1 != rx1) | (y1 != ry1)))
    # The reference sampler is DISCONTINUOUS at the edge of its support: crossing x_pix = W-1 (or 0)
    # switches from "interpolate pixels W-2, W-1" to "clamped corners whose weights cancel" (~0).
    # A pixel whose coordinate differs by one fp32 ulp can land on either side, so pixels whose
    # integer corner differs between the two coordinate sets are excluded from the value tolerance
    # (they are still covered by the bit-exact sampler check below, on the kernel's own x, y).
    same = ((x0 == rx0) & (y0 == ry0) & (x1 == rx1) & (y1 == ry1)).reshape(b, h, w)
    eo_same = np.abs(out - r_out)[same].max()
    print('%s coord err %.2e  pixel err %.2e (all) %.2e (same corners)  corner flips vs oracle coords %.4f%%' %
          (shape, ex, eo, eo_same, 100 * flips))
    assert ex <= 2e-5 and eo_same <= 1e-4 and flips <= 5e-3
    np.testing.assert_array_equal(out, O.tps_interpolate(u, x, y, h, w).reshape(out.shape))


@pytest.mark.parametrize('m', [4, 5, 6])
def test_tps_irregular_and_mixed_meshes(m):
    """Per-frame meshes: frame 0 keeps the regular (separable) mesh, frame 1 an irregular one, frame 2 a separable but
    non-uniform one.  The tile kernel's compact separable-mesh tables and its generic records, and the direct kernel,
    perform the same arithmetic: identical bits; values against the oracle within the stated tolerances."""
    from coupe.dvsg_b200 import ops
    b, h, w = 3, 96, 160
    rng = np.random.default_rng(7 + m)
    u = smooth_image(rng, b, h, w, 3)
    coord = tiled_mesh(m, m, b)
    coord[1] += rng.uniform(-0.05, 0.05, coord[1].shape).astype(np.float32)
    reg = np.linspace(-1, 1, m)
    gx = (reg + rng.uniform(-0.05, 0.05, m)).astype(np.float32)
    gy = (reg + rng.uniform(-0.05, 0.05, m)).astype(np.float32)
    coord[2] = np.stack(np.meshgrid(gx, gy), -1).reshape(m * m, 2)
    vec = rng.uniform(-0.08, 0.08, coord.shape).astype(np.float32)
    U, C_, S = cu(u), cu(coord), cu(vec)
    T = ops.tps_solve(C_, C_ + S)
    a = ops.tps_warp_fwd(U, C_, T, (h, w), want_grid=True, want_mask=True)
    d = ops.tps_warp_fwd(U, C_, T, (h, w), want_grid=True, want_mask=True, flags=FORCE_DIRECT)
    for p_, q_ in zip(a, d):
        assert torch.equal(p_, q_)
    for i in range(b):      # a frame's result does not depend on which table variant its batch neighbours take
        one = ops.tps_warp_fwd(U[i:i + 1].contiguous(), C_[i:i + 1].contiguous(), T[i:i + 1].contiguous(), (h, w), want_grid=True)
        assert torch.equal(one[0][0], a[0][i]) and torch.equal(one[1], a[1][i * h * w:(i + 1) * h * w])
    r_out, r_x, r_y = O.thin_plate_spline(u, coord, vec, (h, w))
    x, y = a[1].cpu().numpy(), a[2].cpu().numpy()
    assert max(np.abs(x - r_x).max(), np.abs(y - r_y).max()) <= 2e-5
    np.testing.assert_array_equal(a[0].cpu().numpy(), O.tps_interpolate(u, x, y, h, w).reshape(b, h, w, 3))


def test_staged_and_direct_kernels_agree_bitwise_at_720p():
    """Full-size property: the shared-memory-staged kernel and the direct-gather kernel perform
    the same arithmetic, so their outputs are identical bit for bit (white-noise frames)."""
    from coupe.dvsg_b200 import ops
    torch.manual_seed(0)
    B, H, W = 4, 720, 1280
    U = torch.rand((B, H, W, 3), device=DEV)
    coord = cu(tiled_mesh(4, 4, 1)[0]).unsqueeze(0).expand(B, -1, -1)
    vec = (torch.rand((B, 16, 2), device=DEV) - 0.5) * 0.2
    T = ops.tps_solve(coord, coord + vec)
    a = ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=True, want_mask=True, flags=TPS_EXACT)
    b = ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=True, want_mask=True, flags=FORCE_DIRECT)
    for p, q in zip(a, b):
        assert torch.equal(p, q)
    # shard invariance: per-frame results do not depend on what else is in the batch (exact and node evaluation)
    c = ops.tps_warp_fwd(U[2:].contiguous(), coord[2:], T[2:].contiguous(), (H, W), want_grid=False, flags=TPS_EXACT)
    assert torch.equal(c[0], a[0][2:])
    n_all = ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=True)
    n_sub = ops.tps_warp_fwd(U[2:].contiguous(), coord[2:], T[2:].contiguous(), (H, W), want_grid=True)
    assert torch.equal(n_sub[0], n_all[0][2:]) and torch.equal(n_sub[1], n_all[1][2 * H * W:])
    # large offsets: footprints that do not fit the staging buffer take the in-kernel fallback
    vec2 = (torch.rand((B, 16, 2), device=DEV) - 0.5) * 1.6
    T2 = ops.tps_solve(coord, coord + vec2)
    a2 = ops.tps_warp_fwd(U, coord, T2, (H, W), want_grid=False, flags=TPS_EXACT)
    b2 = ops.tps_warp_fwd(U, coord, T2, (H, W), want_grid=False, flags=FORCE_DIRECT)
    assert torch.equal(a2[0], b2[0])


def test_4k_16x16_mesh_properties():
    """BASELINE configs[4] shape (2160x3840 frames, 16x16 mesh, N = 259: prepared solve, generic 48-byte tables, XU-bound
    kernel): tile kernel == direct kernel bit for bit; an identity mesh reproduces the A4 sampler's own identity map
    (x_pix = j * W / (W - 1), last row / column zero) against the oracle's sampler on the kernel's coordinates."""
    from coupe.dvsg_b200 import ops
    torch.manual_seed(1)
    B, H, W, m = 1, 2160, 3840, 16
    U = torch.rand((B, H, W, 3), device=DEV)
    mesh = cu(tiled_mesh(m, m, 1)[0])
    coord = mesh.unsqueeze(0).expand(B, -1, -1)
    vec = (torch.rand((B, m * m, 2), device=DEV) - 0.5) * 0.04
    T = ops.tps_solve(coord, coord + vec)
    a = ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=True, flags=TPS_EXACT)
    b = ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=True, flags=FORCE_DIRECT)
    for p_, q_ in zip(a[:3], b[:3]):
        assert torch.equal(p_, q_)
    T0 = ops.tps_solve(coord, coord)
    out, x, y, _ = ops.tps_warp_fwd(U, coord, T0, (H, W), want_grid=True)
    xs = x.reshape(H, W)[::270, ::480].cpu().numpy()
    assert np.abs(xs - O.tf_linspace(-1.0, 1.0, W)[::480][None]).max() <= 2e-4      # stated looser bound for N = 259
    # the identity map zeroes the last row / column: the clamped corners' paired weights cancel (ThinPlateSpline.py:57-60,81-88)
    assert float(out[0, -1].abs().max()) <= 1e-5 and float(out[0, :, -1].abs().max()) <= 1e-5


def test_mask_output_equals_warp_of_ones():
    """N1: mask_out == ThinPlateSpline(ones_like(U), ...) (model.py:82,85), exactly."""
    from coupe.dvsg_b200.ThinPlateSpline import ThinPlateSpline, ThinPlateSplineWithMask
    g = load_golden('tps_mask')
    U = cu(smooth_image(np.random.default_rng(0), 1, 24, 32, 3))
    out, mask, x, y = ThinPlateSplineWithMask(U, cu(g['coord']), cu(g['second']), [24, 32])
    ones_out, _, _ = ThinPlateSpline(torch.ones_like(U), cu(g['coord']), cu(g['second']), [24, 32])
    assert torch.equal(mask.contiguous(), ones_out)
    assert np.abs(ones_out.cpu().numpy() - g['out']).max() <= 1e-5
    plain, _, _ = ThinPlateSpline(U, cu(g['coord']), cu(g['second']), [24, 32])
    assert torch.equal(out, plain)


@pytest.mark.parametrize('name', ['bilinear_c18', 'bilinear_c3'])
def test_bilinear_interp_vs_golden_bit_exact(name):
    from coupe.dvsg_b200.spatial_transformer import _interpolate
    g = load_golden(name)
    out = _interpolate(cu(g['im']), cu(g['x']), cu(g['y']), [int(v) for v in g['out_size']], 'bilinear')
    assert tuple(out.shape) == g['out'].shape
    np.testing.assert_array_equal(out.cpu().numpy(), g['out'])


@pytest.mark.parametrize('shape', [(2, 64, 96, 3), (1, 45, 67, 3), (2, 30, 40, 18), (1, 288, 512, 3), (2, 31, 45, 4), (1, 17, 23, 7), (2, 44, 52, 3)])
def test_bilinear_interp_vs_oracle_bit_exact(shape):
    from coupe.dvsg_b200.spatial_transformer import bilinear_interp
    from coupe.dvsg_b200 import _lib, ops
    b, h, w, c = shape
    rng = np.random.default_rng(sum(shape))
    im = rng.random(shape, dtype=np.float32)
    oh, ow = h, w
    x = rng.uniform(-1.2, 1.2, b * oh * ow).astype(np.float32)
    y = rng.uniform(-1.2, 1.2, b * oh * ow).astype(np.float32)
    # a smooth part so that the staged path sees compact footprints too
    gx, gy = np.meshgrid(np.linspace(-1, 1, ow), np.linspace(-1, 1, oh))
    half = b * oh * ow // 2
    x[:half] = (np.tile(gx.reshape(-1), b)[:half] * 1.05 + 0.01).astype(np.float32)
    y[:half] = (np.tile(gy.reshape(-1), b)[:half] * 0.95 - 0.02).astype(np.float32)
    ref = O.bilinear_interp(im, x, y, (oh, ow))
    out = bilinear_interp(cu(im), cu(x), cu(y), [oh, ow])
    np.testing.assert_array_equal(out.cpu().numpy(), ref)
    lib = _lib.load()
    out2 = torch.empty_like(out)
    I, X, Y = cu(im), cu(x), cu(y)
    rc = lib.dvsg_bilinear_fwd(I.data_ptr(), X.data_ptr(), Y.data_ptr(), out2.data_ptr(), b, h, w, c, oh, ow, FORCE_DIRECT, 0)
    assert rc == 0
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out2.cpu().numpy(), ref)
    del ops


@pytest.mark.parametrize('name', ['flow_small', 'flow_large'])
def test_tf_warp_vs_golden_bit_exact(name):
    from coupe.dvsg_b200.warp_with_optical_flow import tf_warp
    g = load_golden(name)
    h, w = g['im'].shape[1:3]
    out = tf_warp(cu(g['im']), cu(g['flow']), h, w)
    np.testing.assert_array_equal(out.cpu().numpy(), g['out'])


@pytest.mark.parametrize('shape', [(2, 128, 192, 3), (1, 51, 75, 3), (1, 40, 44, 5), (2, 37, 50, 6), (2, 44, 52, 3)])
def test_tf_warp_vs_oracle_bit_exact(shape):
    from coupe.dvsg_b200.warp_with_optical_flow import tf_warp
    b, h, w, c = shape
    rng = np.random.default_rng(sum(shape))
    im = rng.random(shape, dtype=np.float32)
    flow = smooth_flow(rng, b, h, w)
    flow[0, :3, :3] = [[[-2.0, -2.0]] * 3, [[0.0, 0.0]] * 3, [[w + 1.0, h + 1.0]] * 3]
    out = tf_warp(cu(im), cu(flow), h, w)
    np.testing.assert_array_equal(out.cpu().numpy(), O.tf_warp(im, flow, h, w))


def test_tf_warp_1080p_properties():
    """Full-size (cfg4 frame shape) properties: zero flow is the identity, an integer shift moves
    pixels exactly, staged == direct bitwise."""
    from coupe.dvsg_b200 import _lib
    from coupe.dvsg_b200.warp_with_optical_flow import tf_warp
    torch.manual_seed(1)
    B, H, W = 2, 1080, 1920
    im = torch.rand((B, H, W, 3), device=DEV)
    assert torch.equal(tf_warp(im, torch.zeros((B, H, W, 2), device=DEV), H, W), im)
    flow = torch.zeros((B, H, W, 2), device=DEV)
    flow[..., 0] = 3.0
    flow[..., 1] = -2.0
    out = tf_warp(im, flow, H, W)
    assert torch.equal(out[:, 2:, :-3], im[:, :-2, 3:])
    assert out[:, :2].abs().max() == 0 and out[:, :, -3:].abs().max() == 0
    flow = (torch.rand((B, H, W, 2), device=DEV) - 0.5) * 16
    a = tf_warp(im, flow, H, W)
    b = torch.empty_like(a)
    rc = _lib.load().dvsg_flow_warp_fwd(im.data_ptr(), flow.data_ptr(), b.data_ptr(), B, H, W, 3, FORCE_DIRECT, 0)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    # and against the oracle itself at this size: one frame, smooth flow with jitter, bit-exact
    rng = np.random.default_rng(4)
    fl = smooth_flow(rng, 1, H, W).astype(np.float32)
    u = im[:1].cpu().numpy()
    got = tf_warp(im[:1].contiguous(), cu(fl), H, W).cpu().numpy()
    np.testing.assert_array_equal(got, O.tf_warp(u, fl, H, W))


def test_st_meshgrid_and_transformers_vs_golden():
    from coupe.dvsg_b200.spatial_transformer import AffineTransformer, ProjectiveTransformer, _meshgrid
    g = load_golden('meshgrid')
    np.testing.assert_array_equal(_meshgrid([int(v) for v in g['out_size']]).cpu().numpy(), g['grid'])
    np.testing.assert_array_equal(_meshgrid([int(v) for v in g['out_size2']]).cpu().numpy()[::97], g['grid2'])
    g = load_golden('projective')
    pt = ProjectiveTransformer([int(v) for v in g['out_size']])
    out = pt.transform(cu(g['im']), cu(g['theta']))
    assert np.abs(out.cpu().numpy() - g['out']).max() <= 1e-5
    np.testing.assert_array_equal(out.cpu().numpy(), O.projective_transform(g['im'], g['theta'], g['out_size']))
    xs, ys = pt._transform(cu(g['im']), cu(g['theta']))
    rx, ry = O.projective_grid(g['theta'], g['out_size'])
    np.testing.assert_array_equal(xs.cpu().numpy(), rx)
    g = load_golden('affine')
    out = AffineTransformer([int(v) for v in g['out_size']]).transform(cu(g['im']), cu(g['theta']))
    assert np.abs(out.cpu().numpy() - g['out']).max() <= 1e-5


def test_empty_and_ragged_inputs():
    from coupe.dvsg_b200.ThinPlateSpline import ThinPlateSpline
    from coupe.dvsg_b200.warp_with_optical_flow import tf_warp
    coord = cu(tiled_mesh(4, 4, 1)[0])
    out, x, y = ThinPlateSpline(torch.zeros((0, 8, 8, 3), device=DEV), coord.unsqueeze(0).expand(0, -1, -1),
                                torch.zeros((0, 16, 2), device=DEV), [8, 8])
    assert out.shape == (0, 8, 8, 3) and x.numel() == 0
    out = tf_warp(torch.zeros((0, 8, 8, 3), device=DEV), torch.zeros((0, 8, 8, 2), device=DEV), 8, 8)
    assert out.shape == (0, 8, 8, 3)
    # 1x1 output and 1-pixel-wide frames: linspace degenerates to the single value -1
    u = torch.rand((1, 5, 1, 3), device=DEV)
    out, x, y = ThinPlateSpline(u, coord.unsqueeze(0), torch.zeros((1, 16, 2), device=DEV), [1, 1])
    r_out, r_x, r_y = O.thin_plate_spline(u.cpu().numpy(), coord.cpu().numpy()[None], np.zeros((1, 16, 2), np.float32), (1, 1))
    assert np.abs(out.cpu().numpy() - r_out).max() <= 1e-5 and abs(float(x[0]) - r_x[0]) <= 1e-5


def test_errors_are_loud():
    from coupe.dvsg_b200.ThinPlateSpline import ThinPlateSpline
    from coupe.dvsg_b200.warp_with_optical_flow import tf_warp
    coord = cu(tiled_mesh(4, 4, 2))
    with pytest.raises(ValueError):
        ThinPlateSpline(torch.zeros((2, 8, 8, 3), device=DEV), coord, torch.zeros((2, 9, 2), device=DEV), [8, 8])
    with pytest.raises(ValueError):
        ThinPlateSpline(torch.zeros((2, 8, 8, 3)), coord, torch.zeros((2, 16, 2), device=DEV), [8, 8])
    with pytest.raises(ValueError):
        tf_warp(torch.zeros((1, 8, 8, 3), device=DEV), torch.zeros((1, 8, 8, 2), device=DEV), 4, 4)
    with pytest.raises(ValueError):
        tf_warp(torch.zeros((1, 8, 8, 3), device=DEV), torch.zeros((1, 8, 9, 2), device=DEV), 8, 8)


def test_host_pipeline_matches_device_path():
    from coupe.dvsg_b200 import ops
    rng = np.random.default_rng(11)
    B, H, W = 7, 48, 64
    u = smooth_image(rng, B, H, W, 3)
    mesh = tiled_mesh(4, 4, 1)[0]
    vec = rng.uniform(-0.1, 0.1, (B, 16, 2)).astype(np.float32)
    pipe = ops.HostPipeline(H, W, 3, 16, frames_per_chunk=2, n_slots=3, device=0)
    out_h = pipe.thin_plate_spline(torch.from_numpy(u).pin_memory(), torch.from_numpy(mesh), torch.from_numpy(vec))
    pipe.close()
    out_d, _, _, _ = run_tps(u, np.tile(mesh[None], (B, 1, 1)), vec, (H, W))
    np.testing.assert_array_equal(out_h.numpy(), out_d)


# ---- N4: frame ingest / egress (eval.py:76-81, 112-113) ---------------------------------------------
@pytest.mark.parametrize('shape', [(2, 48, 64), (1, 37, 53), (3, 1, 5)])
@pytest.mark.parametrize('swap', [True, False])
def test_frame_ingest_egress_bit_exact(shape, swap):
    """u/255 and uint8(x*255.) on the device are bit-exact against the reference's NumPy arithmetic: every uint8 value,
    random fp32 values, values on and next to every k/255 boundary, negatives, > 1, inf and NaN."""
    from coupe.dvsg_b200 import frame_io
    b, h, w = shape
    rng = np.random.default_rng(3)
    f8 = rng.integers(0, 256, (b, h, w, 3), dtype=np.uint8)
    f8.reshape(-1)[:min(256, f8.size)] = np.arange(min(256, f8.size), dtype=np.uint8)
    got = frame_io.frames_u8_to_f32(cu(f8), swap_rb=swap).cpu().numpy()
    np.testing.assert_array_equal(got, O.frames_u8_to_f32(f8, swap))
    x = rng.uniform(-0.1, 1.1, (b, h, w, 3)).astype(np.float32)
    flat = x.reshape(-1)
    k = (np.arange(flat.size) % 256).astype(np.float32)
    edge = (k / np.float32(255)).astype(np.float32)
    third = flat.size // 3
    flat[:third] = np.nextafter(edge[:third], np.float32(0))
    flat[third:2 * third] = np.nextafter(edge[third:2 * third], np.float32(2))
    if flat.size > 8:
        flat[-8:] = [np.nan, np.inf, -np.inf, 1e10, -1e10, 300.0 / 255.0, -3.7, 1.0]
    got8 = frame_io.frames_f32_to_u8(cu(x), swap_rb=swap).cpu().numpy()
    np.testing.assert_array_equal(got8, O.frames_f32_to_u8(x, swap))
    # unaligned views take the element-wise kernels
    if f8.size > 16:
        base = cu(np.concatenate([np.zeros(1, np.uint8), f8.reshape(-1)]))
        v = base[1:].reshape(b, h, w, 3)
        np.testing.assert_array_equal(frame_io.frames_u8_to_f32(v, swap_rb=swap).cpu().numpy(), O.frames_u8_to_f32(f8, swap))


def test_host_pipeline_u8_matches_reference_loop():
    """uint8 BGR frames in, uint8 BGR frames out: identical to ingest -> fp32 ThinPlateSpline (device path) -> egress
    computed step by step with the oracle's eval.py restatement around the kernel's own fp32 warp."""
    from coupe.dvsg_b200 import ops
    rng = np.random.default_rng(12)
    B, H, W = 5, 48, 64
    f8 = (smooth_image(rng, B, H, W, 3) * 255).astype(np.uint8)
    mesh = tiled_mesh(4, 4, 1)[0]
    vec = rng.uniform(-0.1, 0.1, (B, 16, 2)).astype(np.float32)
    pipe = ops.HostPipeline(H, W, 3, 16, frames_per_chunk=2, n_slots=3, device=0)
    out8 = pipe.thin_plate_spline_u8(torch.from_numpy(f8).pin_memory(), torch.from_numpy(mesh), torch.from_numpy(vec))
    again = pipe.thin_plate_spline_u8(torch.from_numpy(f8).pin_memory(), torch.from_numpy(mesh), torch.from_numpy(vec))
    pipe.close()
    u = O.frames_u8_to_f32(f8, True)
    out_d, x, y, _ = run_tps(u, np.tile(mesh[None], (B, 1, 1)), vec, (H, W))
    np.testing.assert_array_equal(out8.numpy(), O.frames_f32_to_u8(out_d, True))
    np.testing.assert_array_equal(out8.numpy(), again.numpy())
    # and against the oracle end to end: the sampler stage is bit-exact on the kernel's coordinates
    ref = O.tps_interpolate(u, x, y, H, W).reshape(B, H, W, 3)
    np.testing.assert_array_equal(out8.numpy(), O.frames_f32_to_u8(ref, True))


@pytest.mark.parametrize('sizes', [(720, 1280, 288, 512), (480, 640, 288, 512), (100, 150, 288, 512), (288, 512, 288, 512), (37, 53, 61, 19), (2, 2, 5, 7)])
def test_read_frames_with_resize_matches_cv2(sizes):
    """eval.py:76-81 including cv2.resize: the device ingest equals cv2's result after the fp32 feed cast.  Stated bar:
    max-abs <= 1.2e-7 (one fp32 ulp at 1.0: the sums are formed in double in both, a different last double bit can move
    the fp32 rounding); in practice bit-identical."""
    import cv2
    from coupe.dvsg_b200 import frame_io
    hs, ws, h, w = sizes
    rng = np.random.default_rng(sum(sizes))
    f8 = rng.integers(0, 256, (2, hs, ws, 3), dtype=np.uint8)
    got = frame_io.read_frames(cu(f8), (h, w)).cpu().numpy()
    for i in range(2):
        ref = cv2.resize(cv2.cvtColor(f8[i], cv2.COLOR_BGR2RGB) / 255., (w, h)).astype(np.float32)
        assert np.abs(got[i] - ref).max() <= 1.2e-7
        assert (got[i] != ref).mean() <= 1e-3
        np.testing.assert_array_equal(got[i], O.read_frame(f8[i], w, h))      # the oracle's restatement: bit-exact
    with pytest.raises(ValueError):
        frame_io.read_frames(cu(f8[:, :1]), (h, w))          # one-row sources are outside the restated algorithm


def test_online_warper_equals_the_dropin():
    """ops.OnlineWarper (prepared mesh, preallocated buffers, one C-ABI call per frame) == ThinPlateSpline, bit for bit."""
    from coupe.dvsg_b200 import ops
    from coupe.dvsg_b200.ThinPlateSpline import ThinPlateSpline
    rng = np.random.default_rng(21)
    B, H, W = 1, 96, 160
    mesh = cu(tiled_mesh(5, 5, 1)[0])
    warper = ops.OnlineWarper(mesh, B, H, W)
    for _ in range(3):
        U = cu(smooth_image(rng, B, H, W, 3))
        vec = cu(rng.uniform(-0.1, 0.1, (B, 25, 2)).astype(np.float32))
        ref, _, _ = ThinPlateSpline(U, mesh, vec, [H, W], return_grid=False)
        assert torch.equal(warper.warp(U, vec), ref)
    with pytest.raises(ValueError):
        warper.warp(U[:, :10], vec)


def test_results_do_not_depend_on_the_tiles_per_cta_cut(monkeypatch):
    """The tile kernels cut every strip of tiles into CTAs with a cost model (tile_pick_seg_len); the cut is
    scheduling only.  Forward outputs are bit-identical for any cut; the backward's grad_image goes through
    per-tile TMA reduce-adds, so only its fp32 summation order may change."""
    from coupe.dvsg_b200 import ops
    torch.manual_seed(5)
    B, H, W = 3, 288, 512
    U = torch.rand((B, H, W, 3), device=DEV)
    coord = cu(tiled_mesh(4, 4, 1)[0]).unsqueeze(0).expand(B, -1, -1)
    vec = (torch.rand((B, 16, 2), device=DEV) - 0.5) * 0.2
    T = ops.tps_solve(coord, coord + vec)
    g = torch.randn((B, H, W, 3), device=DEV)
    flow = cu(smooth_flow(np.random.default_rng(3), B, H, W).astype(np.float32))
    from coupe.dvsg_b200.warp_with_optical_flow import tf_warp

    def run():
        f = ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=True, want_mask=True)
        b = ops.tps_warp_bwd(U, coord, T, (H, W), g, None, None, need_grad_U=True, want_grid_grad=True)
        w = tf_warp(U, flow, H, W)
        torch.cuda.synchronize()
        return f, b, w

    monkeypatch.delenv('DVSG_FWD_SEGLEN', raising=False)
    monkeypatch.delenv('DVSG_BWD_SEGLEN', raising=False)
    f0, b0, w0 = run()
    for seg in ('1', '4', '7', '16'):
        monkeypatch.setenv('DVSG_FWD_SEGLEN', seg)
        monkeypatch.setenv('DVSG_BWD_SEGLEN', seg)
        f1, b1, w1 = run()
        for p, q in zip(f0, f1):
            assert torch.equal(p, q), seg
        assert torch.equal(w0, w1), seg
        gU0, gT0, gx0, gy0 = b0
        gU1, gT1, gx1, gy1 = b1
        assert torch.equal(gx0, gx1) and torch.equal(gy0, gy1), seg
        assert float((gU0 - gU1).abs().max()) <= 2e-5 * float(gU0.abs().max()), seg
        assert float((gT0 - gT1).abs().max()) <= 2e-5 * float(gT0.abs().max()), seg


def test_tps_forward_vs_oracle_at_the_north_star_shape():
    """One frame at the headline shape (1080 x 1920 x 3, 4x4 mesh, offsets +-0.1; BASELINE north_star) against the
    oracle itself, not only through size-independent properties: coordinates <= 2e-5, pixels <= 1e-4 where both
    coordinate sets pick the same corners, sampler stage bit-exact on the kernel's own coordinates."""
    b, h, w, c, m = 1, 1080, 1920, 3, 4
    rng = np.random.default_rng(1080)
    # the coordinate noise (2e-6 normalised) is 4x larger in pixels than at 288 x 512: periods >= 64 px keep the pixel bar meaningful
    u = smooth_image(rng, b, h, w, c, period=64.0)
    coord = tiled_mesh(m, m, b)
    vec = rng.uniform(-0.1, 0.1, coord.shape).astype(np.float32)
    out, x, y, T = run_tps(u, coord, vec, (h, w), 1, 0)
    r_out, r_x, r_y = O.thin_plate_spline(u, coord, vec, (h, w))
    ex = max(np.abs(x - r_x).max(), np.abs(y - r_y).max())
    # all four clamped corners must agree: at x_pix = 0- / 0+ the low corner is pixel 0 on both sides (clamped from -1 or
    # not) while the high one flips between 0 (weights cancel) and 1 (interpolation)
    _, _, x0, x1, y0, y1 = O.tps_sample_indices(x, y, h, w)
    _, _, rx0, rx1, ry0, ry1 = O.tps_sample_indices(r_x, r_y, h, w)
    same = ((x0 == rx0) & (y0 == ry0) & (x1 == rx1) & (y1 == ry1)).reshape(b, h, w)
    flips = 1.0 - float(same.mean())
    eo_same = np.abs(out - r_out)[same].max()
    print('1080p coord err %.2e  pixel err %.2e (same corners)  corner flips %.4f%%' % (ex, eo_same, 100 * flips))
    assert ex <= 2e-5 and eo_same <= 1e-4 and flips <= 5e-3
    np.testing.assert_array_equal(out, O.tps_interpolate(u, x, y, h, w).reshape(out.shape))


@pytest.mark.parametrize('m,size', [(16, (135, 240)), (8, (90, 160))])
def test_large_mesh_grid_stage_vs_oracle_on_the_kernels_coefficients(m, size):
    """Large meshes (16x16 = BASELINE configs[4], N = 259): the solve is ill-conditioned, so the grid stage
    is compared on its own -- the oracle's `_transform` arithmetic (ThinPlateSpline.py:92-133) fed with the KERNEL's
    coefficients T.  Bar: the kernel's coordinates are no farther from the fp64 evaluation of that stage than twice the
    fp32 oracle's own distance to it (sequential fp32 sum of N terms), and <= 2e-4 absolute."""
    from coupe.dvsg_b200 import ops
    h, w = size
    rng = np.random.default_rng(m)
    coord = tiled_mesh(m, m, 1)
    vec = rng.uniform(-0.02, 0.02, coord.shape).astype(np.float32)
    U = cu(smooth_image(rng, 1, h, w, 3))
    T = ops.tps_solve(cu(coord), cu(coord + vec))
    _, x, y, _ = ops.tps_warp_fwd(U, cu(coord), T, (h, w), want_grid=True)
    x, y, Tn = x.cpu().numpy(), y.cpu().numpy(), T.cpu().numpy()
    x32, y32 = O.tps_grid(Tn, coord, h, w)
    x64, y64 = O.tps_grid(Tn.astype(np.float64), coord.astype(np.float64), h, w, dtype=np.float64)
    e_k = max(np.abs(x - x64.reshape(-1)).max(), np.abs(y - y64.reshape(-1)).max())
    e_o = max(np.abs(x32.reshape(-1) - x64.reshape(-1)).max(), np.abs(y32.reshape(-1) - y64.reshape(-1)).max())
    print('mesh %dx%d grid stage: |kernel - fp64| %.2e, |fp32 oracle - fp64| %.2e, max |T| %.1f' % (m, m, e_k, e_o, np.abs(Tn).max()))
    assert e_k <= max(2.0 * e_o, 2e-5) and e_k <= 2e-4


@pytest.mark.parametrize('m', [4, 5, 16])
def test_solve_from_offsets_equals_solve_from_the_sum(m):
    """dvsg_tps_solve_offsets_prepared forms target = coord + vector (ThinPlateSpline.py:161) inside the apply kernel:
    the same fp32 add, so T is bit-identical to adding first; per-frame meshes fall back to the explicit add."""
    from coupe.dvsg_b200 import ops
    B = 5
    rng = np.random.default_rng(m)
    mesh = cu(tiled_mesh(m, m, 1)[0])
    vec = cu(rng.uniform(-0.1, 0.1, (B, m * m, 2)).astype(np.float32))
    shared = mesh.unsqueeze(0).expand(B, -1, -1)
    assert torch.equal(ops.tps_solve(shared, vec, offsets=True), ops.tps_solve(shared, shared + vec))
    per_frame = shared.contiguous()
    assert torch.equal(ops.tps_solve(per_frame, vec, offsets=True), ops.tps_solve(per_frame, per_frame + vec))


def _same_corner_mask(x, y, rx, ry, h, w):
    _, _, x0, x1, y0, y1 = O.tps_sample_indices(x, y, h, w)
    _, _, a0, a1, b0, b1 = O.tps_sample_indices(rx, ry, h, w)
    return (x0 == a0) & (y0 == b0) & (x1 == a1) & (y1 == b1)


def test_tps_forward_vs_oracle_one_full_cfg2_frame():
    """BASELINE configs[1] shape, one whole 720 x 1280 frame, 4x4 mesh, offsets +-0.1, against O.thin_plate_spline ITSELF:
    coordinates <= 2e-5, pixels <= 1e-4 where both coordinate sets pick the same corners, flips <= 0.5 % and
    the flipped pixels' error bounded too (bilinear interpolation is C0 across a flip); sampler bit-exact."""
    b, h, w, c, m = 1, 720, 1280, 3, 4
    rng = np.random.default_rng(720)
    u = smooth_image(rng, b, h, w, c, period=48.0)
    coord = tiled_mesh(m, m, b)
    vec = rng.uniform(-0.1, 0.1, coord.shape).astype(np.float32)
    out, x, y, T = run_tps(u, coord, vec, (h, w), 1, 0)
    r_out, r_x, r_y = O.thin_plate_spline(u, coord, vec, (h, w))
    ex = max(np.abs(x - r_x).max(), np.abs(y - r_y).max())
    same = _same_corner_mask(x, y, r_x, r_y, h, w).reshape(b, h, w)
    err = np.abs(out - r_out).max(axis=-1)
    # A flip INSIDE the frame moves a sample across a cell boundary of a C0 surface: the pixel changes continuously.  A flip
    # AT the frame border is different: the A4 sampler weights from the clamped corners (ThinPlateSpline.py:57-60,81-88), so
    # its output jumps from ~0 (both corners clamped onto one pixel, weights cancel) to the image value when x_pix crosses
    # 0 or W-1 -- the edge of the validity mask is discontinuous in the reference itself.
    xp, yp, _, _, _, _ = O.tps_sample_indices(r_x, r_y, h, w)
    inside = ((xp > 0.5) & (xp < w - 1.5) & (yp > 0.5) & (yp < h - 1.5)).reshape(b, h, w)
    flip_in, flip_edge = ~same & inside, ~same & ~inside
    e_same = err[same].max()
    e_flip = err[flip_in].max() if flip_in.any() else 0.0
    print('720p full frame: coord err %.2e, pixel err %.2e on same corners, %.2e on the %d flipped px inside the frame, %d flips on the '
          'mask edge (%.4f%% flipped in all)' % (ex, e_same, e_flip, int(flip_in.sum()), int(flip_edge.sum()), 100.0 * (1.0 - same.mean())))
    assert ex <= 2e-5 and e_same <= 1e-4 and e_flip <= 2e-4 and (1.0 - same.mean()) <= 5e-3
    np.testing.assert_array_equal(out, O.tps_interpolate(u, x, y, h, w).reshape(out.shape))


def test_tps_forward_vs_oracle_one_full_cfg5_frame():
    """BASELINE configs[4] shape: ONE WHOLE 2160 x 3840 frame with the 16x16 mesh (N = 259) against the oracle, in row
    bands (the reference's basis for this frame is 8.6 GB; O.tps_grid(rows=...) evaluates the same element-wise ops band by
    band).  The fp32 SOLVE of the reference is ill-conditioned here (cond 3.9e4: its coefficients carry 1.6e-3 of noise,
    SURVEY H4), so the grid stage is fed with the kernel's own coefficients, as in
    test_large_mesh_grid_stage_vs_oracle_on_the_kernels_coefficients.  Bars, every pixel of the frame:
      * coordinates no farther from the fp64 evaluation than 2x the fp32 oracle's own distance to it, and <= 3e-5;
      * pixels on same-corner samples <= 1e-4 + 2x the fp32 oracle's own pixel distance to the fp64 run (one fp32 ulp of a
        coordinate is 2.4e-4 px at W = 3840: SURVEY H3);
      * sampler stage bit-exact on the kernel's own coordinates."""
    from concurrent.futures import ThreadPoolExecutor
    import os
    from coupe.dvsg_b200 import ops
    b, h, w, c, m = 1, 2160, 3840, 3, 16
    rng = np.random.default_rng(2160)
    # periods >= 512 px: |dI/dx| <~ 0.006 per px, so that the 1e-4 pixel bar is about the arithmetic and not about the
    # 0.02 px of fp32 noise any 259-term fp32 sum carries at this width
    u = smooth_image(rng, b, h, w, c, period=512.0)
    coord = tiled_mesh(m, m, b)
    vec = rng.uniform(-0.02, 0.02, coord.shape).astype(np.float32)
    U, C_ = cu(u), cu(coord)
    T = ops.tps_solve(C_, C_ + cu(vec))
    out, x, y, _ = ops.tps_warp_fwd(U, C_, T, (h, w), want_grid=True)
    torch.cuda.synchronize()
    out, x, y, Tn = out.cpu().numpy(), x.cpu().numpy().reshape(h, w), y.cpu().numpy().reshape(h, w), T.cpu().numpy()
    np.testing.assert_array_equal(out, O.tps_interpolate(u, x.reshape(-1), y.reshape(-1), h, w).reshape(out.shape))
    T64, c64 = Tn.astype(np.float64), coord.astype(np.float64)
    band = 36

    def one(r0):
        rows = (r0, min(r0 + band, h))
        x32, y32 = O.tps_grid(Tn, coord, h, w, rows=rows)
        x64, y64 = O.tps_grid(T64, c64, h, w, dtype=np.float64, rows=rows)
        xk, yk = x[rows[0]:rows[1]].reshape(-1), y[rows[0]:rows[1]].reshape(-1)
        e_k = max(np.abs(xk - x64).max(), np.abs(yk - y64).max())
        e_o = max(np.abs(x32 - x64).max(), np.abs(y32 - y64).max())
        nb = rows[1] - rows[0]
        o32 = O.tps_interpolate(u, x32, y32, nb, w)
        o64 = O.tps_interpolate(u.astype(np.float64), x64, y64, nb, w, dtype=np.float64)
        ok = out[0, rows[0]:rows[1]].reshape(-1, c)
        s_k = _same_corner_mask(xk, yk, x32, y32, h, w)
        s_o = _same_corner_mask(x32, y32, x64.astype(np.float32), y64.astype(np.float32), h, w)
        p_k = np.abs(ok - o32).max(axis=1)
        p_o = np.abs(o32 - o64).max(axis=1)
        return e_k, e_o, (p_k[s_k].max() if s_k.any() else 0.0), (p_o[s_o].max() if s_o.any() else 0.0), int((~s_k).sum())

    with ThreadPoolExecutor(max_workers=max(1, min(os.cpu_count() or 1, 12))) as ex:
        res = list(ex.map(one, range(0, h, band)))
    e_k, e_o = max(r[0] for r in res), max(r[1] for r in res)
    p_k, p_o = max(r[2] for r in res), max(r[3] for r in res)
    flips = sum(r[4] for r in res)
    print('4K 16x16 full frame: |kernel - fp64| %.2e vs |fp32 oracle - fp64| %.2e; pixels (same corners) kernel-oracle %.2e, '
          'oracle fp32-fp64 %.2e; corner flips %d of %d px' % (e_k, e_o, p_k, p_o, flips, h * w))
    assert e_k <= max(2.0 * e_o, 1e-5) and e_k <= 3e-5
    assert p_k <= 1e-4 + 2.0 * p_o
    assert flips <= 5e-3 * h * w


# ---- tile-node evaluation of the TPS map (the tile kernels' default) against the per-pixel evaluation and the oracle ----
NODE_CASES = [
    # H, W, mesh, offset amplitude, jittered mesh
    (288, 512, 16, 0.02, False), (200, 400, 8, 0.05, True), (288, 512, 5, 0.1, False), (288, 512, 5, 0.1, True), (540, 960, 4, 0.1, False), (540, 960, 5, 0.3, False),
    (720, 1280, 4, 0.1, False), (720, 1280, 5, 0.3, True), (1080, 1920, 5, 0.1, False), (1080, 1920, 8, 0.05, True),
    (1080, 1920, 16, 0.03, False),      # two-level evaluation (16 x 16 mesh, <= 32 super-near control points per super-tile)
]


@pytest.mark.parametrize('case', NODE_CASES, ids=lambda c: '%dx%d_m%d_a%g%s' % (c[0], c[1], c[2], c[3], '_jit' if c[4] else ''))
def test_node_evaluation_vs_exact_evaluation_and_fp64_oracle(case):
    """The gate of the tile-node evaluation (far field of the spline on 6 x 5 Chebyshev nodes per 32 x 8 tile, near control
    points per pixel): on EVERY pixel of the frame -- the ones next to control points included, reported separately --
      * |node coords - fp64 evaluation| <= max(2x the fp32 oracle's own distance to it, 2e-6) -- no farther from the true
        map than twice the reference's own fp32 arithmetic -- and <= 2x the exact kernel's distance + 3e-7 (the fp32 rounding
        noise of the node values passes through interpolation weights whose absolute sum is <= 4.2; the APPROXIMATION error
        itself is pinned in fp64 by tests/test_node_model.py: <= 2e-7);
      * the sampler stage is bit-exact on the node coordinates; pixels <= 1e-4 vs the exact kernel on same-corner samples."""
    from coupe.dvsg_b200 import _lib, ops
    h, w, m, amp, jit = case
    rng = np.random.default_rng(h + 7 * m)
    u = smooth_image(rng, 1, h, w, 3, period=max(32.0, w / 30.0))
    coord = tiled_mesh(m, m, 1)
    if jit:
        coord = (coord + rng.uniform(-0.3, 0.3, coord.shape) / m).astype(np.float32)
    vec = rng.uniform(-amp, amp, coord.shape).astype(np.float32)
    assert _lib.load().dvsg_tps_coords_mode(h, w, 3, h, w, m * m, 0) == 1
    U, C_ = cu(u), cu(coord)
    T = ops.tps_solve(C_, C_ + cu(vec))
    on, xn, yn, _ = ops.tps_warp_fwd(U, C_, T, (h, w), want_grid=True)
    oe, xe, ye, _ = ops.tps_warp_fwd(U, C_, T, (h, w), want_grid=True, flags=TPS_EXACT)
    torch.cuda.synchronize()
    on, xn, yn, oe, xe, ye, Tn = (t.cpu().numpy() for t in (on, xn, yn, oe, xe, ye, T))
    np.testing.assert_array_equal(on, O.tps_interpolate(u, xn, yn, h, w).reshape(on.shape))
    x64 = np.empty(h * w); y64 = np.empty(h * w); x32 = np.empty(h * w, np.float32); y32 = np.empty(h * w, np.float32)
    band = max(8, (1 << 22) // (w * (m * m + 3)))
    for r0 in range(0, h, band):
        rows = (r0, min(r0 + band, h))
        sl = slice(rows[0] * w, rows[1] * w)
        x64[sl], y64[sl] = O.tps_grid(Tn.astype(np.float64), coord.astype(np.float64), h, w, dtype=np.float64, rows=rows)
        x32[sl], y32[sl] = O.tps_grid(Tn, coord, h, w, rows=rows)
    e_n = np.maximum(np.abs(xn - x64), np.abs(yn - y64))
    e_e = np.maximum(np.abs(xe - x64), np.abs(ye - y64))
    e_o = np.maximum(np.abs(x32 - x64), np.abs(y32 - y64))
    # pixels within 3 px of a control point
    ccol, crow = (coord[0, :, 0] + 1) * (w - 1) / 2, (coord[0, :, 1] + 1) * (h - 1) / 2
    cols, rows_ = np.meshgrid(np.arange(w), np.arange(h))
    at_cp = np.zeros((h, w), bool)
    for a, b in zip(ccol, crow):
        at_cp |= (np.abs(cols - a) <= 3) & (np.abs(rows_ - b) <= 3)
    at_cp = at_cp.reshape(-1)
    print('%s: |nodes-fp64| %.2e (next to control points %.2e), |exact-fp64| %.2e, |fp32 oracle-fp64| %.2e, |nodes-exact| %.2e' %
          (case, e_n.max(), e_n[at_cp].max() if at_cp.any() else 0.0, e_e.max(), e_o.max(), max(np.abs(xn - xe).max(), np.abs(yn - ye).max())))
    assert e_n.max() <= max(2.0 * e_o.max(), 2e-6)
    assert e_n.max() <= 2.0 * e_e.max() + 3e-7
    same = _same_corner_mask(xn, yn, xe, ye, h, w).reshape(1, h, w)
    # two fp32 evaluations of the map differ by their rounding noise (~ the fp32 oracle's distance to fp64, which grows with
    # the mesh: 1e-5 at 16x16); in pixels that is noise * W/2 * |dI/dx|
    gmax = max(np.abs(np.diff(u, axis=2)).max(), np.abs(np.diff(u, axis=1)).max())
    allow = 1e-4 + 2.0 * e_o.max() * max(w, h) / 2.0 * gmax
    d_px = np.abs(on - oe).max(axis=-1)[same].max()
    print('   pixels nodes vs exact kernel on same corners: %.2e (allowance %.2e), corner flips %.4f%%' % (d_px, allow, 100.0 * (1.0 - same.mean())))
    assert d_px <= allow
    assert (1.0 - same.mean()) <= 5e-3 + 2e3 * e_o.max()


@pytest.mark.parametrize('case', [(2, 300, 500, (722, 1284), 4, True), (3, 724, 1276, (362, 1412), 5, False)],
                         ids=['upscale_722x1284_per_frame_meshes', 'resize_362x1412_shared_mesh'])
def test_node_evaluation_with_resize_ragged_edges_batch_and_mask(case):
    """Node mode away from the square cases above: output size != input size, ow not a multiple of the 32-pixel tile, oh not
    a multiple of the 8-row strip, several frames (per-frame meshes / one shared mesh), mask output.  Against the fp32 oracle
    on every pixel: coordinates <= 2e-5 (the bar of the exact path), sampler and mask bit-exact on the kernel's own
    coordinates; against the per-pixel kernel: coordinates <= 3e-6."""
    from coupe.dvsg_b200 import _lib, ops
    b, h, w, (oh, ow), m, per_frame = case
    rng = np.random.default_rng(oh + ow)
    u = smooth_image(rng, b, h, w, 3, period=40.0)
    coord = tiled_mesh(m, m, b)
    if per_frame:
        coord = (coord + rng.uniform(-0.2, 0.2, coord.shape) / m).astype(np.float32)
    vec = rng.uniform(-0.1, 0.1, coord.shape).astype(np.float32)
    assert _lib.load().dvsg_tps_coords_mode(h, w, 3, oh, ow, m * m, 0) == 1
    U = cu(u)
    C_ = cu(coord) if per_frame else cu(coord[0])
    T = ops.tps_solve(C_, cu(coord + vec))
    on, xn, yn, mn = ops.tps_warp_fwd(U, C_, T, (oh, ow), want_grid=True, want_mask=True)
    oe, xe, ye, me = ops.tps_warp_fwd(U, C_, T, (oh, ow), want_grid=True, want_mask=True, flags=TPS_EXACT)
    torch.cuda.synchronize()
    on, xn, yn, mn, xe, ye, Tn = (t.cpu().numpy() for t in (on, xn, yn, mn, xe, ye, T))
    r_x, r_y = O.tps_grid(Tn, coord, oh, ow)
    e = max(np.abs(xn - r_x).max(), np.abs(yn - r_y).max())
    d = max(np.abs(xn - xe).max(), np.abs(yn - ye).max())
    print('%s: |nodes - fp32 oracle| %.2e, |nodes - exact kernel| %.2e' % (case, e, d))
    assert e <= 2e-5 and d <= 3e-6
    np.testing.assert_array_equal(on, O.tps_interpolate(u, xn, yn, oh, ow).reshape(on.shape))
    np.testing.assert_array_equal(mn, O.tps_interpolate(np.ones((b, h, w, 1), np.float32), xn, yn, oh, ow).reshape(mn.shape))


@pytest.mark.parametrize('shape', [(1, 288, 512, 4), (3, 288, 512, 5), (2, 720, 1280, 4), (1, 96, 160, 5), (2, 40, 50, 4), (1, 288, 512, 6)],
                         ids=lambda s: 'B%d_%dx%d_m%d' % s)
def test_online_call_is_one_launch_and_bit_identical_to_solve_then_warp(shape):
    """dvsg_tps_warp_frames_offsets (the per-frame call of eval.py:101-124): for pn + 3 <= 32 on the tile path the prepared
    solve runs in the prologue of the warp kernel -- ONE launch -- with the arithmetic of tps_apply_kernel; coefficients,
    frames, grid and mask are bit-identical to dvsg_tps_solve_offsets_prepared followed by dvsg_tps_warp_fwd.  Larger
    systems (6x6: N = 39) and shapes off the tile path (40 x 50) keep the two launches."""
    from coupe.dvsg_b200 import _lib
    lib = _lib.load()
    B, H, W, m = shape
    pn = m * m
    rng = np.random.default_rng(sum(shape))
    U = cu(smooth_image(rng, B, H, W, 3))
    mesh = cu(tiled_mesh(m, m, 1)[0])
    vec = cu(rng.uniform(-0.1, 0.1, (B, pn, 2)).astype(np.float32))
    nbytes = lib.dvsg_tps_prepare_workspace_bytes(B, pn, 0)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    assert lib.dvsg_tps_prepare(mesh.data_ptr(), 0, B, pn, ws.data_ptr(), nbytes, 0) == 0

    def bufs():
        return (torch.empty((B, 2, pn + 3), device=DEV), torch.empty((B, H, W, 3), device=DEV), torch.empty(B * H * W, device=DEV),
                torch.empty(B * H * W, device=DEV), torch.empty((B, H, W), device=DEV))
    T1, o1, x1, y1, m1 = bufs()
    n0 = _lib.launch_count()
    rc = lib.dvsg_tps_warp_frames_offsets(U.data_ptr(), mesh.data_ptr(), vec.data_ptr(), ws.data_ptr(), nbytes, T1.data_ptr(), o1.data_ptr(),
                                          x1.data_ptr(), y1.data_ptr(), m1.data_ptr(), B, H, W, 3, H, W, pn, 0)
    assert rc == 0
    launches = _lib.launch_count() - n0
    fused = pn + 3 <= 32 and W % 4 == 0 and W >= 32 and H >= 8
    assert launches == (1 if fused else 2)
    T2, o2, x2, y2, m2 = bufs()
    assert lib.dvsg_tps_solve_offsets_prepared(mesh.data_ptr(), 0, vec.data_ptr(), T2.data_ptr(), B, pn, ws.data_ptr(), nbytes, 0) == 0
    assert lib.dvsg_tps_warp_fwd(U.data_ptr(), mesh.data_ptr(), 0, T2.data_ptr(), o2.data_ptr(), x2.data_ptr(), y2.data_ptr(), m2.data_ptr(),
                                 B, H, W, 3, H, W, pn, 0, 0) == 0
    torch.cuda.synchronize()
    for a, b_ in ((T1, T2), (o1, o2), (x1, x2), (y1, y2), (m1, m2)):
        assert torch.equal(a, b_)


# ---- N4: flow ingest and ElasticTransformer ------------------------------------------------------------------
@pytest.mark.parametrize('sizes', [(72, 128, 288, 512), (288, 512, 288, 512), (180, 320, 288, 512), (360, 640, 288, 512), (100, 150, 61, 19), (1, 1, 5, 7)])
def test_read_flow_matches_cv2(sizes):
    """data_loader.py:239 on the device: bit-identical to `cv2.resize(flow, (w, h)) * [w, h]` (cast by the float32 feed) and to
    the oracle's restatement, single fields and batches."""
    cv2 = pytest.importorskip('cv2')
    from coupe.dvsg_b200 import frame_io
    hs, ws, h, w = sizes
    rng = np.random.default_rng(hs + w)
    flows = rng.uniform(-0.05, 0.05, (3, hs, ws, 2)).astype(np.float32)
    got = frame_io.read_flow(cu(flows), (h, w)).cpu().numpy()
    for i in range(3):
        ref = (cv2.resize(flows[i], (w, h)).reshape(h, w, 2) * [w, h]).astype(np.float32)
        np.testing.assert_array_equal(got[i], ref)
        np.testing.assert_array_equal(got[i], O.flow_ingest(flows[i], w, h))
    np.testing.assert_array_equal(frame_io.read_flow(cu(flows[1]), (h, w)).cpu().numpy(), got[1])
    # the ingested field is what tf_warp takes (warp_with_optical_flow.py:117-120): channel 0 = dx, 1 = dy, pixels
    if h >= 8 and w >= 8:
        from coupe.dvsg_b200.warp_with_optical_flow import tf_warp
        im = rng.random((3, h, w, 3), dtype=np.float32)
        out = tf_warp(cu(im), cu(got), h, w).cpu().numpy()
        np.testing.assert_array_equal(out, O.tf_warp(im, got, h, w))


@pytest.mark.parametrize('name', ['elastic_4x4', 'elastic_3x3_resize'])
def test_elastic_transformer_vs_reference_golden(name):
    """ElasticTransformer(out_size, 2*g*g, g).transform(inp, theta) against spatial_transformer.py:93-362 run unmodified
    (make_golden.elastic_cases): coordinates <= 2e-5, pixels <= 1e-4, sampler bit-exact on the kernel's coordinates,
    gradients w.r.t. the input and theta rel <= 1e-4 (corner flips removed exactly would need the TPS helper: the frames
    are small and no corner flips at this noise level; asserted)."""
    from coupe.dvsg_b200.spatial_transformer import ElasticTransformer
    g = load_golden(name)
    gs, osz = int(g['grid_size']), [int(v) for v in g['out_size']]
    et = ElasticTransformer(osz, param_dim=2 * gs * gs, param_dim_per_side=gs)
    im = cu(g['im']).requires_grad_(True)
    theta = cu(g['theta']).requires_grad_(True)
    out, x, y = et.transform(im, theta)
    assert out.shape == tuple(g['out'].shape)
    xn, yn = x.detach().cpu().numpy(), y.detach().cpu().numpy()
    ex = max(np.abs(xn - g['x']).max(), np.abs(yn - g['y']).max())
    eo = np.abs(out.detach().cpu().numpy() - g['out']).max()
    print('%s: coord err %.2e pixel err %.2e' % (name, ex, eo))
    assert ex <= 2e-5 and eo <= 1e-4
    np.testing.assert_array_equal(out.detach().cpu().numpy().reshape(-1, g['im'].shape[3]), O.bilinear_interp(g['im'], xn, yn, osz))
    (out * cu(g['g_out'])).sum().backward()
    H, W = g['im'].shape[1:3]
    same = np.array_equal(np.floor((xn + 1) / 2 * (W - 1)), np.floor((g['x'] + 1) / 2 * (W - 1))) and \
        np.array_equal(np.floor((yn + 1) / 2 * (H - 1)), np.floor((g['y'] + 1) / 2 * (H - 1)))
    assert same, 'a sampling corner flipped: compare stage-wise instead'
    gi, gt = im.grad.cpu().numpy(), theta.grad.cpu().numpy()
    assert np.abs(gi - g['grad_im']).max() <= 1e-4 * np.abs(g['grad_im']).max()
    assert np.abs(gt - g['grad_theta']).max() <= 1e-4 * np.abs(g['grad_theta']).max()
    np.testing.assert_allclose(et.get_abs_theta(cu(g['theta'])).cpu().numpy(), g['abs_theta'], atol=1e-6)
    np.testing.assert_allclose(et.get_abs_src_points(g['im'].shape[0]).cpu().numpy(), g['abs_src'], atol=1e-6)


def test_elastic_transformer_vs_oracle_at_the_training_shape():
    """288 x 512, 4x4 control mesh (the constructor defaults), offsets +-0.1: coordinates vs the oracle's restatement of
    _initialize_tps / _transform <= 2e-5, sampler bit-exact; identity theta reproduces the identity grid."""
    from coupe.dvsg_b200.spatial_transformer import ElasticTransformer
    b, h, w = 2, 288, 512
    rng = np.random.default_rng(93)
    im = smooth_image(rng, b, h, w, 3)
    theta = rng.uniform(-0.1, 0.1, (b, 32)).astype(np.float32)
    et = ElasticTransformer([h, w])
    out, x, y = et.transform(cu(im), cu(theta))
    xn, yn = x.cpu().numpy(), y.cpu().numpy()
    r_out, r_x, r_y = O.elastic_transform(im, theta, 4, (h, w))
    assert max(np.abs(xn - r_x).max(), np.abs(yn - r_y).max()) <= 2e-5
    np.testing.assert_array_equal(out.cpu().numpy().reshape(-1, 3), O.bilinear_interp(im, xn, yn, (h, w)))
    _, x0, y0 = et.transform(cu(im), cu(np.zeros((b, 32), np.float32)))
    gx = np.tile(O.tf_linspace(-1.0, 1.0, w)[None], (h, 1)).reshape(-1)
    assert np.abs(x0.cpu().numpy()[:h * w] - gx).max() <= 2e-5


def test_host_pipeline_streaming_mesh_change_and_tf_warp():
    """HostPipeline: (1) asynchronous (streaming) submission gives the same frames as blocking calls; (2) a second mesh makes
    the pipeline invert the new system (the inverse is cached per mesh, not per call); (3) tf_warp on host buffers equals the
    device tf_warp bit for bit; (4) a 6x6 mesh (N = 39 > 32: two launches per chunk) works through the same pipeline."""
    from coupe.dvsg_b200 import ops
    from coupe.dvsg_b200.ThinPlateSpline import ThinPlateSpline
    from coupe.dvsg_b200.warp_with_optical_flow import tf_warp
    rng = np.random.default_rng(77)
    B, H, W = 7, 96, 160
    u = smooth_image(rng, B, H, W, 3)
    U_h = torch.from_numpy(u).pin_memory()
    for m in (4, 6):
        pn = m * m
        pipe = ops.HostPipeline(H, W, 3, pn, frames_per_chunk=2, n_slots=3)
        for trial in range(2):      # two different meshes through the same pipeline
            mesh = tiled_mesh(m, m, 1)[0] + (rng.uniform(-0.02, 0.02, (pn, 2)).astype(np.float32) if trial else 0)
            vec = rng.uniform(-0.1, 0.1, (B, pn, 2)).astype(np.float32)
            ref, _, _ = ThinPlateSpline(cu(u), cu(mesh), cu(vec), [H, W], return_grid=False)
            out_b = pipe.thin_plate_spline(U_h, torch.from_numpy(mesh), torch.from_numpy(vec))
            assert torch.equal(out_b, ref.cpu())
            pipe.set_async(True)
            outs = [torch.empty_like(U_h).pin_memory() for _ in range(3)]
            for o in outs:
                pipe.thin_plate_spline(U_h, torch.from_numpy(mesh), torch.from_numpy(vec), o)
            pipe.sync()
            pipe.set_async(False)
            for o in outs:
                assert torch.equal(o, ref.cpu())
        pipe.close()
    flow = smooth_flow(rng, B, H, W).astype(np.float32)
    pipe = ops.HostPipeline(H, W, 3, 16, frames_per_chunk=3, n_slots=2)
    got = pipe.tf_warp(U_h, torch.from_numpy(flow).pin_memory())
    assert torch.equal(got, tf_warp(cu(u), cu(flow), H, W).cpu())
    np.testing.assert_array_equal(got.numpy(), O.tf_warp(u, flow, H, W))
    pipe.close()
