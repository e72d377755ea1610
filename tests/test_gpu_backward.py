"""GPU parity tests, backward path (K4, K5-bwd, K6-bwd) through the C ABI.

Gradients are discontinuous across sampling-cell boundaries (the bilinear surface is C0), so
stage-wise comparisons are made on IDENTICAL coordinates: the kernel's own x, y are fed to the
oracle's backward formulas.  Stated tolerance: rel <= 1e-4 of the max-norm (fp32; grad_image
and grad_T are accumulated with atomics, so their summation order is not reproducible).
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import smooth_flow, smooth_image, tiled_mesh
from oracle import dvsg_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


@pytest.mark.parametrize('merge', [1, 0], ids=['merged', 'plain'])
@pytest.mark.parametrize('name', ['tps_4x4', 'tps_5x5', 'tps_4x4_big', 'tps_resize', 'tps_8x8'])
def test_tps_backward_stages_vs_golden(name, merge):
    from coupe.dvsg_b200 import _lib, ops
    _lib.load().dvsg_set_bwd_tuning(merge)
    try:
        g = load_golden(name)
        oh, ow = (int(v) for v in g['out_size'])
        U, C_ = cu(g['u']), cu(g['coord'])
        target = C_ + cu(g['second'])
        T = ops.tps_solve(C_, target)
        _, x, y, _ = ops.tps_warp_fwd(U, C_, T, (oh, ow))
        gU, gT, gxs, gys = ops.tps_warp_bwd(U, C_, T, (oh, ow), cu(g['g_out']), cu(g['g_x']), cu(g['g_y']), want_grid_grad=True)
        gvec = ops.tps_solve_bwd(C_, gT)
        torch.cuda.synchronize()
        x, y = x.cpu().numpy(), y.cpu().numpy()
        # stage 1: sampler backward on the kernel's own coordinates
        r_gim, r_gx, r_gy = O.tps_interpolate_bwd(g['u'], x, y, oh, ow, g['g_out'])
        r_gx, r_gy = r_gx + g['g_x'], r_gy + g['g_y']
        assert rel(gU.cpu().numpy(), r_gim) <= 1e-4
        assert rel(gxs.cpu().numpy(), r_gx) <= 1e-4 and rel(gys.cpu().numpy(), r_gy) <= 1e-4
        # stage 2: grad_T = sum_pix grad * basis (oracle accumulates in fp64)
        r_gT = O.tps_grid_bwd(g['coord'], oh, ow, gxs.cpu().numpy(), gys.cpu().numpy())
        assert rel(gT.cpu().numpy(), r_gT) <= 1e-4
        # stage 3: chain through W^-T to the offsets
        _, w_inv = O.tps_solve(g['coord'].astype(np.float64), (g['coord'] + g['second']).astype(np.float64), dtype=np.float64, return_inverse=True)
        r_gvec = O.tps_solve_bwd(w_inv, gT.cpu().numpy().astype(np.float64))
        assert rel(gvec.cpu().numpy(), r_gvec) <= 1e-4
        # end to end against TF-autodiff-equivalent goldens, when no sampling corner flipped
        _, _, x0, _, y0, _ = O.tps_sample_indices(x, y, g['u'].shape[1], g['u'].shape[2])
        _, _, gx0, _, gy0, _ = O.tps_sample_indices(g['x'], g['y'], g['u'].shape[1], g['u'].shape[2])
        if np.array_equal(x0, gx0) and np.array_equal(y0, gy0):
            tol = 2e-3 if name == 'tps_8x8' else 3e-4
            assert rel(gvec.cpu().numpy(), g['grad_second']) <= tol
            assert rel(gU.cpu().numpy(), g['grad_u']) <= tol
    finally:
        _lib.load().dvsg_set_bwd_tuning(1)


def test_tps_autograd_dropin():
    """loss.backward() through the drop-in ThinPlateSpline reaches U and vector (trainer.py:108-111)."""
    from coupe.dvsg_b200.ThinPlateSpline import ThinPlateSpline
    g = load_golden('tps_4x4')
    U = cu(g['u']).requires_grad_(True)
    V = cu(g['second']).requires_grad_(True)
    out, x, y = ThinPlateSpline(U, cu(g['coord']), V, [int(v) for v in g['out_size']])
    loss = (out * cu(g['g_out'])).sum() + (x * cu(g['g_x'])).sum() + (y * cu(g['g_y'])).sum()
    loss.backward()
    assert rel(V.grad.cpu().numpy(), g['grad_second']) <= 5e-3     # e2e: a flipped corner moves this by O(1e-3)
    assert rel(U.grad.cpu().numpy(), g['grad_u']) <= 5e-3
    # grid-only loss (surf loss shape): no image gradient requested
    V2 = cu(g['second']).requires_grad_(True)
    _, x2, y2 = ThinPlateSpline(cu(g['u']), cu(g['coord']), V2, [int(v) for v in g['out_size']])
    (x2 * cu(g['g_x'])).sum().backward()
    assert torch.isfinite(V2.grad).all() and V2.grad.abs().max() > 0


# (2, 288, 512, 3, 4): the training shape of BASELINE configs[2] (two frames of it: the oracle's backward is NumPy)
# (1, 44, 52, 3, 4): tile kernel with a ragged last strip (4 of 8 rows) and a ragged last tile (20 of 32 columns)
@pytest.mark.parametrize('shape', [(2, 96, 128, 3, 4), (1, 45, 50, 3, 5), (1, 32, 48, 2, 4), (2, 288, 512, 3, 4), (1, 44, 52, 3, 4)])
def test_tps_backward_vs_oracle_seeded(shape):
    from coupe.dvsg_b200 import ops
    b, h, w, c, m = shape
    rng = np.random.default_rng(sum(shape))
    u = smooth_image(rng, b, h, w, c)
    coord = tiled_mesh(m, m, b)
    vec = rng.uniform(-0.1, 0.1, coord.shape).astype(np.float32)
    g_out = rng.standard_normal((b, h, w, c)).astype(np.float32)
    U, C_ = cu(u), cu(coord)
    T = ops.tps_solve(C_, C_ + cu(vec))
    _, x, y, _ = ops.tps_warp_fwd(U, C_, T, (h, w))
    # upstream gradients on the returned x, y (surf loss, trainer.py:363-386) add to the sampler's own
    gx_in = rng.standard_normal(b * h * w).astype(np.float32)
    gy_in = rng.standard_normal(b * h * w).astype(np.float32)
    gU, gT, gxs, gys = ops.tps_warp_bwd(U, C_, T, (h, w), cu(g_out), cu(gx_in), cu(gy_in), want_grid_grad=True)
    x, y = x.cpu().numpy(), y.cpu().numpy()
    r_gim, r_gx, r_gy = O.tps_interpolate_bwd(u, x, y, h, w, g_out)
    r_gx, r_gy = r_gx.reshape(-1) + gx_in, r_gy.reshape(-1) + gy_in
    assert rel(gU.cpu().numpy(), r_gim) <= 1e-4
    assert rel(gxs.cpu().numpy(), r_gx) <= 1e-4 and rel(gys.cpu().numpy(), r_gy) <= 1e-4
    assert rel(gT.cpu().numpy(), O.tps_grid_bwd(coord, h, w, gxs.cpu().numpy(), gys.cpu().numpy())) <= 1e-4


@pytest.mark.parametrize('name', ['bilinear_c18', 'bilinear_c3'])
def test_bilinear_backward_vs_golden(name):
    from coupe.dvsg_b200.spatial_transformer import bilinear_interp
    g = load_golden(name)
    im = cu(g['im']).requires_grad_(True)
    x = cu(g['x']).requires_grad_(True)
    y = cu(g['y']).requires_grad_(True)
    out = bilinear_interp(im, x, y, [int(v) for v in g['out_size']])
    (out * cu(g['g_out'])).sum().backward()
    assert rel(im.grad.cpu().numpy(), g['grad_im']) <= 1e-4
    assert rel(x.grad.cpu().numpy(), g['grad_x']) <= 1e-4 and rel(y.grad.cpu().numpy(), g['grad_y']) <= 1e-4


@pytest.mark.parametrize('name', ['flow_small', 'flow_large'])
def test_flow_warp_backward_vs_golden(name):
    from coupe.dvsg_b200.warp_with_optical_flow import tf_warp
    g = load_golden(name)
    h, w = g['im'].shape[1:3]
    im = cu(g['im']).requires_grad_(True)
    flow = cu(g['flow']).requires_grad_(True)
    (tf_warp(im, flow, h, w) * cu(g['g_out'])).sum().backward()
    assert rel(im.grad.cpu().numpy(), g['grad_im']) <= 1e-4
    assert rel(flow.grad.cpu().numpy(), g['grad_flow']) <= 1e-4


@pytest.mark.parametrize('shape', [(2, 72, 100, 3), (1, 44, 52, 3)], ids=['72x100', '44x52'])
@pytest.mark.parametrize('merge', [1, 0], ids=['merged', 'plain'])
def test_flow_and_bilinear_backward_vs_oracle_seeded(merge, shape):
    from coupe.dvsg_b200 import _lib
    from coupe.dvsg_b200.spatial_transformer import bilinear_interp
    from coupe.dvsg_b200.warp_with_optical_flow import tf_warp
    _lib.load().dvsg_set_bwd_tuning(merge)
    try:
        rng = np.random.default_rng(7)
        b, h, w, c = shape
        im = rng.random((b, h, w, c), dtype=np.float32)
        flow = smooth_flow(rng, b, h, w)
        g_out = rng.standard_normal((b, h, w, c)).astype(np.float32)
        I, F = cu(im).requires_grad_(True), cu(flow).requires_grad_(True)
        (tf_warp(I, F, h, w) * cu(g_out)).sum().backward()
        r_gim, r_gflow = O.tf_warp_bwd(im, flow, h, w, g_out)
        assert rel(I.grad.cpu().numpy(), r_gim) <= 1e-4 and rel(F.grad.cpu().numpy(), r_gflow) <= 1e-4
        gx, gy = np.meshgrid(np.linspace(-1.1, 1.1, w), np.linspace(-1.05, 1.05, h))
        x = np.tile(gx.reshape(-1), b).astype(np.float32) + rng.uniform(-0.01, 0.01, b * h * w).astype(np.float32)
        y = np.tile(gy.reshape(-1), b).astype(np.float32)
        I2, X, Y = cu(im).requires_grad_(True), cu(x).requires_grad_(True), cu(y).requires_grad_(True)
        (bilinear_interp(I2, X, Y, [h, w]) * cu(g_out).reshape(-1, c)).sum().backward()
        r_gim, r_gx, r_gy = O.bilinear_interp_bwd(im, x, y, (h, w), g_out)
        assert rel(I2.grad.cpu().numpy(), r_gim) <= 1e-4
        assert rel(X.grad.cpu().numpy(), r_gx) <= 1e-4 and rel(Y.grad.cpu().numpy(), r_gy) <= 1e-4
    finally:
        _lib.load().dvsg_set_bwd_tuning(1)


def test_backward_linearity_at_training_shape():
    """cfg3 shape (batch 32, 288x512, 4x4 mesh): the backward is linear in grad_out, and the
    image gradient of an all-ones grad_out sums to the sum of the validity mask (every output
    pixel distributes exactly its four weights)."""
    from coupe.dvsg_b200 import ops
    torch.manual_seed(3)
    B, H, W = 32, 288, 512
    U = torch.rand((B, H, W, 3), device=DEV)
    coord = cu(tiled_mesh(4, 4, 1)[0]).unsqueeze(0).expand(B, -1, -1)
    vec = (torch.rand((B, 16, 2), device=DEV) - 0.5) * 0.2
    T = ops.tps_solve(coord, coord + vec)
    g1 = torch.rand((B, H, W, 3), device=DEV)
    g2 = torch.rand((B, H, W, 3), device=DEV)
    a = ops.tps_warp_bwd(U, coord, T, (H, W), g1, None, None)
    b = ops.tps_warp_bwd(U, coord, T, (H, W), g2, None, None)
    c = ops.tps_warp_bwd(U, coord, T, (H, W), g1 + 2 * g2, None, None)
    # the frame's border ring collects the clamped corners of every out-of-frame sample, whose
    # weights are large and cancel pairwise (|w| ~ distance outside the frame): its rounding noise
    # scales with sum|w*g|, not with the net value, in the reference too -- compare the interior
    inner = (slice(None), slice(1, -1), slice(1, -1))
    assert float(((a[0] + 2 * b[0]) - c[0])[inner].abs().max()) <= 1e-4 * float(c[0][inner].abs().max())
    assert float(((a[1] + 2 * b[1]) - c[1]).abs().max()) <= 2e-4 * float(c[1].abs().max())
    ones = torch.ones((B, H, W, 3), device=DEV)
    gU = ops.tps_warp_bwd(U, coord, T, (H, W), ones, None, None)[0]
    _, _, _, mask = ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=False, want_mask=True)
    assert abs(float(gU.double().sum()) - 3 * float(mask.double().sum())) <= 1e-5 * 3 * float(mask.double().sum())


def test_tps_solve_bwd_shared_mesh_matches_per_frame_solve():
    """The batched shared-mesh kernel (one CTA eliminates [W^T | grad_T of 32 frames]) and the
    one-warp-per-frame kernel run the same fp64 Gauss-Jordan: identical results."""
    from coupe.dvsg_b200 import ops
    rng = np.random.default_rng(5)
    for n, B in ((4, 70), (5, 3)):
        coord = tiled_mesh(n, n, B)
        gT = rng.standard_normal((B, 2, n * n + 3)).astype(np.float32)
        dense = ops.tps_solve_bwd(cu(coord), cu(gT)).cpu().numpy()
        shared = ops.tps_solve_bwd(cu(coord[0]).unsqueeze(0).expand(B, -1, -1), cu(gT)).cpu().numpy()
        assert np.abs(dense - shared).max() <= 1e-6 * np.abs(dense).max()
        target = (coord + rng.uniform(-0.1, 0.1, coord.shape)).astype(np.float32)
        Td = ops.tps_solve(cu(coord), cu(target)).cpu().numpy()
        Ts = ops.tps_solve(cu(coord[0]).unsqueeze(0).expand(B, -1, -1), cu(target)).cpu().numpy()
        assert np.abs(Td - Ts).max() <= 1e-6


@pytest.mark.parametrize('amp', [0.0, 0.2, 1.2], ids=['identity', 'offsets', 'folding'])
def test_tile_and_generic_backward_kernels_agree_at_training_shape(amp):
    """Full-size property (batch 8 at 288x512, the training shape): the warp-autonomous tile kernel
    (TMA-staged gather, shared-memory accumulation, TMA reduce-add) and the generic kernel
    (global gathers, red.global) compute the same gradients; only the floating-point summation
    order of grad_image / grad_T differs.  amp=1.2 folds the mesh so that the per-pixel paths of the
    tile kernel (oversize footprints, non-monotone rows, frame border) are exercised too."""
    from coupe.dvsg_b200 import _lib, ops
    torch.manual_seed(3)
    B, H, W = 8, 288, 512
    U = torch.rand((B, H, W, 3), device=DEV)
    coord = cu(tiled_mesh(4, 4, 1)[0]).unsqueeze(0).expand(B, -1, -1)
    vec = (torch.rand((B, 16, 2), device=DEV) - 0.5) * amp
    T = ops.tps_solve(coord, coord + vec)
    g = torch.randn((B, H, W, 3), device=DEV)
    gx_in, gy_in = torch.randn(B * H * W, device=DEV), torch.randn(B * H * W, device=DEV)
    lib = _lib.load()
    try:
        lib.dvsg_set_bwd_tuning(1)
        a = ops.tps_warp_bwd(U, coord, T, (H, W), g, gx_in, gy_in, need_grad_U=True, want_grid_grad=True)
        lib.dvsg_set_bwd_tuning(1 | 2)      # force the generic kernel
        b = ops.tps_warp_bwd(U, coord, T, (H, W), g, gx_in, gy_in, need_grad_U=True, want_grid_grad=True)
    finally:
        lib.dvsg_set_bwd_tuning(1)
    for name, p, q in zip(('grad_U', 'grad_T', 'grad_xs', 'grad_ys'), a, b):
        if name == 'grad_U':
            # frame-border pixels collect the clamped out-of-frame samples, whose weights are large and cancel
            # (+-|distance|): their sum depends on the summation order at the 1e-4 level in ANY implementation
            # (the generic kernel with and without its shuffle merge differ by as much); strict inside
            inner = (p - q)[:, 2:-2, 2:-2].abs().max()
            assert float(inner) / float(q.abs().max()) <= 2e-5, (name, float(inner))
            if amp < 1.0:      # with the mesh folded, samples land hundreds of pixels outside: border sums are pure cancellation
                assert float((p - q).abs().max()) / float(q.abs().max()) <= 2e-3
            continue
        err = float((p - q).abs().max()) / max(float(q.abs().max()), 1e-30)
        assert err <= 2e-5, (name, err)
    # flow warp and given-grid sampler
    flow = (torch.rand((B, H, W, 2), device=DEV) - 0.5) * 6.0
    x = torch.rand(B * H * W, device=DEV) * 2.2 - 1.1
    y = torch.rand(B * H * W, device=DEV) * 2.2 - 1.1
    xs = torch.linspace(-1.02, 1.02, W, device=DEV).repeat(B * H) + (torch.rand(B * H * W, device=DEV) - 0.5) * 0.004
    ys = torch.linspace(-1.02, 1.02, H, device=DEV).repeat_interleave(W).repeat(B)
    res = []
    for mode in (1, 1 | 2):
        try:
            lib.dvsg_set_bwd_tuning(mode)
            gi = torch.zeros_like(U); gf = torch.empty_like(flow)
            rc = lib.dvsg_flow_warp_bwd(U.data_ptr(), flow.data_ptr(), g.data_ptr(), gi.data_ptr(), gf.data_ptr(), B, H, W, 3, 0)
            assert rc == 0
            outs = [gi, gf]
            for xx, yy in ((x, y), (xs.contiguous(), ys.contiguous())):
                gi2 = torch.zeros_like(U); gxx = torch.empty_like(xx); gyy = torch.empty_like(yy)
                rc = lib.dvsg_bilinear_bwd(U.data_ptr(), xx.data_ptr(), yy.data_ptr(), g.data_ptr(), gi2.data_ptr(), gxx.data_ptr(), gyy.data_ptr(),
                                           B, H, W, 3, H, W, 0)
                assert rc == 0
                outs += [gi2, gxx, gyy]
            torch.cuda.synchronize()
            res.append(outs)
        finally:
            lib.dvsg_set_bwd_tuning(1)
    for p, q in zip(*res):
        err = float((p - q).abs().max()) / max(float(q.abs().max()), 1e-30)
        assert err <= 2e-5, err


def test_flow_warp_backward_vs_oracle_at_720p():
    """One 720 x 1280 frame (many strips, several CTAs per strip): tf_warp's gradients w.r.t. image and flow against the
    oracle's scatter-add, rel <= 1e-4."""
    from coupe.dvsg_b200.warp_with_optical_flow import tf_warp
    rng = np.random.default_rng(720)
    b, h, w, c = 1, 720, 1280, 3
    im = rng.random((b, h, w, c), dtype=np.float32)
    flow = smooth_flow(rng, b, h, w).astype(np.float32)
    g_out = rng.standard_normal((b, h, w, c)).astype(np.float32)
    I, F = cu(im).requires_grad_(True), cu(flow).requires_grad_(True)
    (tf_warp(I, F, h, w) * cu(g_out)).sum().backward()
    r_gim, r_gflow = O.tf_warp_bwd(im, flow, h, w, g_out)
    assert rel(I.grad.cpu().numpy(), r_gim) <= 1e-4 and rel(F.grad.cpu().numpy(), r_gflow) <= 1e-4
