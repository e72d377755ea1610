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
from helpers import close_elementwise, rel_max, smooth_flow, smooth_image, tiled_mesh, tps_flip_correction
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
        # element-wise beside the max-norm: |a-b| <= 1e-4 |b| + 1e-5 max|b| on every element of every stage
        for nm, a, b in (('grad_U', gU, r_gim), ('grad_xs', gxs, r_gx), ('grad_ys', gys, r_gy), ('grad_T', gT, r_gT), ('grad_vector', gvec, r_gvec)):
            ok, worst, at = close_elementwise(a.cpu().numpy(), b, rtol=1e-4, atol_frac=1e-5)
            print('%s %s: element-wise worst error / allowance = %.3f' % (name, nm, worst))
            assert ok, (nm, worst, at)
        # ---- end to end against the reference's own autodiff (goldens), UNCONDITIONALLY ----
        # Pixels whose sampling corner differs between the two evaluations (floor of a coordinate within rounding noise
        # of an integer) change d out/d x discontinuously; their exact effect on grad_vector is computed and removed
        # (helpers.tps_flip_correction), everything else is compared on all elements.  Yardstick: the fp64 run of the
        # reference graph -- this kernel must be no farther from it than max(1e-4, 2x the reference's own fp32 run).
        tgt = g['coord'] + g['second']
        x64, y64 = g['x64'].astype(np.float32), g['y64'].astype(np.float32)
        n_k, corr_k = tps_flip_correction(O, g['u'], g['coord'], tgt, (oh, ow), x, y, x64, y64, g['g_out'])
        n_r, corr_r = tps_flip_correction(O, g['u'], g['coord'], tgt, (oh, ow), g['x'], g['y'], x64, y64, g['g_out'])
        for nm, ours, ref32, ref64 in (('grad_vector', gvec.cpu().numpy() + corr_k, g['grad_second'] + corr_r, g['grad_second64']),
                                       ('grad_U', gU.cpu().numpy(), g['grad_u'], g['grad_u64'])):
            d_ours, d_ref = rel_max(ours, ref64), rel_max(ref32, ref64)
            print('%s %s e2e: kernel-fp64 %.2e, reference fp32-fp64 %.2e (corner flips excluded exactly: kernel %d, reference %d of %d px)'
                  % (name, nm, d_ours, d_ref, n_k, n_r, x.size))
            assert d_ours <= max(1e-4, 2.0 * d_ref), (nm, d_ours, d_ref)
            ok, worst, at = close_elementwise(ours, ref64, rtol=1e-4, atol_frac=max(1e-4, 2.0 * d_ref))
            assert ok, (nm, worst, at)
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
    # end to end through autograd: corner flips removed exactly (see test_tps_backward_stages_vs_golden), then rel <= 1e-4
    oh, ow = (int(v) for v in g['out_size'])
    n, corr = tps_flip_correction(O, g['u'], g['coord'], g['coord'] + g['second'], (oh, ow), x.detach().cpu().numpy(), y.detach().cpu().numpy(),
                                  g['x'], g['y'], g['g_out'])
    print('dropin e2e: %d corner flips of %d px' % (n, x.numel()))
    assert rel(V.grad.cpu().numpy() + corr, g['grad_second']) <= 1e-4
    assert rel(U.grad.cpu().numpy(), g['grad_u']) <= 1e-4
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
    r_gT = O.tps_grid_bwd(coord, h, w, gxs.cpu().numpy(), gys.cpu().numpy())
    assert rel(gT.cpu().numpy(), r_gT) <= 1e-4
    for nm, a, b_ in (('grad_U', gU, r_gim), ('grad_xs', gxs, r_gx), ('grad_ys', gys, r_gy), ('grad_T', gT, r_gT)):
        ok, worst, at = close_elementwise(a.cpu().numpy().reshape(-1), np.asarray(b_).reshape(-1), rtol=1e-4, atol_frac=1e-5)
        assert ok, (nm, worst, at)


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
        # DVSG_FLAG_TPS_EXACT: the generic kernel evaluates every radial term per pixel; the tile kernel's default (tile
        # nodes) yields coordinates ~1e-6 away, i.e. different corners on a few pixels
        a = ops.tps_warp_bwd(U, coord, T, (H, W), g, gx_in, gy_in, need_grad_U=True, want_grid_grad=True, flags=2)
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


def test_mask_variant_is_differentiable_like_the_plain_dropin():
    """N1 (model.py:81-85): ThinPlateSplineWithMask returns output, x, y WITH a graph -- gradients to U and vector are
    bit-identical to ThinPlateSpline's (same kernels, same arguments) -- and the mask without one (its gradient is
    identically zero: the four weights sum to a piecewise constant of the coordinates)."""
    from coupe.dvsg_b200.ThinPlateSpline import ThinPlateSpline, ThinPlateSplineWithMask
    g = load_golden('tps_4x4')
    osz = [int(v) for v in g['out_size']]
    grads = []
    for fn in (ThinPlateSpline, ThinPlateSplineWithMask):
        U = cu(g['u']).requires_grad_(True)
        V = cu(g['second']).requires_grad_(True)
        res = fn(U, cu(g['coord']), V, osz)
        out, x, y = (res[0], res[1], res[2]) if fn is ThinPlateSpline else (res[0], res[2], res[3])
        if fn is ThinPlateSplineWithMask:
            mask = res[1]
            assert not mask.requires_grad and out.requires_grad and x.requires_grad
            ones, _, _ = ThinPlateSpline(torch.ones_like(U), cu(g['coord']), cu(g['second']), osz)
            assert torch.equal(mask.contiguous(), ones.detach())
            loss = (out * mask * cu(g['g_out'])).sum()           # the mask gates a loss, as in trainer.py:232-243
        else:
            ones, _, _ = ThinPlateSpline(torch.ones_like(U), cu(g['coord']), cu(g['second']), osz)
            loss = (out * ones.detach() * cu(g['g_out'])).sum()
        (loss + (x * cu(g['g_x'])).sum() + (y * cu(g['g_y'])).sum()).backward()
        grads.append((U.grad.clone(), V.grad.clone()))
    # grad_U is accumulated with atomics (order-dependent at the ulp level): compare within fp32 noise; grad_vector too
    assert rel(grads[1][0].cpu().numpy(), grads[0][0].cpu().numpy()) <= 1e-6
    assert rel(grads[1][1].cpu().numpy(), grads[0][1].cpu().numpy()) <= 1e-5


def test_grad_image_fixed_point_bound_with_1e6_dynamic_range_inside_a_tile():
    """grad_image of the tile kernel is accumulated in 32-bit fixed point scaled per 32x8 tile from max|grad_out|
    (DESIGN.md K4): each contribution is rounded to 2^-23 of the TILE's largest gradient -- an ABSOLUTE bound, stated
    and checked here against the generic kernel's fp32 scatter with grad_out spanning 1e6 inside every tile."""
    from coupe.dvsg_b200 import _lib, ops
    torch.manual_seed(11)
    B, H, W = 2, 64, 128
    U = torch.rand((B, H, W, 3), device=DEV)
    coord = cu(tiled_mesh(4, 4, 1)[0]).unsqueeze(0).expand(B, -1, -1)
    vec = (torch.rand((B, 16, 2), device=DEV) - 0.5) * 0.1
    T = ops.tps_solve(coord, coord + vec)
    g = torch.randn((B, H, W, 3), device=DEV) * torch.pow(10.0, -6.0 * torch.rand((B, H, W, 1), device=DEV))
    lib = _lib.load()
    try:
        lib.dvsg_set_bwd_tuning(1)
        a = ops.tps_warp_bwd(U, coord, T, (H, W), g, None, None)[0]
        lib.dvsg_set_bwd_tuning(1 | 2)
        b = ops.tps_warp_bwd(U, coord, T, (H, W), g, None, None)[0]
    finally:
        lib.dvsg_set_bwd_tuning(1)
    # up to ~16 contributions per source pixel, each within 2^-23 of the tile maximum (<= the global maximum here)
    bound = 16 * 2.0 ** -23 * float(g.abs().max())
    err = float((a - b).abs().max())
    print('fixed-point grad_image: max abs deviation %.3e (bound %.3e, max|grad_out| %.3e)' % (err, bound, float(g.abs().max())))
    assert err <= bound


def test_shared_mesh_solve_equals_per_frame_solve_bitwise_on_the_5x5_mesh():
    """the batched shared-mesh elimination (one CTA, row groups) had a racing store at pivot steps with
    piv % 8 != 0 (5x5 mesh, k = 1).  Same fp64 Gauss-Jordan, same pivot rule, same operation order as the one-warp-per-
    frame kernel: the coefficients must agree BIT FOR BIT, run after run."""
    from coupe.dvsg_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(55)
    for n in (5, 4, 3):
        B, pn = 67, n * n
        mesh = tiled_mesh(n, n, 1)[0]
        target = (mesh[None] + rng.uniform(-0.1, 0.1, (B, pn, 2))).astype(np.float32)
        C1, Cb, Tg = cu(mesh), cu(np.tile(mesh[None], (B, 1, 1))), cu(target)
        outs = []
        for rep in range(8):
            Ts = torch.empty((B, 2, pn + 3), device=DEV)
            assert lib.dvsg_tps_solve(C1.data_ptr(), 0, Tg.data_ptr(), Ts.data_ptr(), B, pn, 0, 0, 0) == 0       # shared kernel
            outs.append(Ts)
        Tw = torch.empty((B, 2, pn + 3), device=DEV)
        assert lib.dvsg_tps_solve(Cb.data_ptr(), 2 * pn, Tg.data_ptr(), Tw.data_ptr(), B, pn, 0, 0, 0) == 0          # warp kernel
        torch.cuda.synchronize()
        for Ts in outs:
            assert torch.equal(Ts, outs[0])
        # the shared kernel takes 1/pivot by Newton steps and compares pivot magnitudes in fp32: equal up to the last fp32 bit
        assert float((outs[0] - Tw).abs().max()) <= 2e-7 * float(Tw.abs().max())


@pytest.mark.parametrize('shape', [(1, 540, 960, 5, 0.1), (1, 720, 1280, 4, 0.1), (2, 288, 512, 8, 0.04), (2, 300, 500, 4, 0.1, 722, 1284), (3, 288, 512, 5, 0.1)],
                         ids=lambda s: '%dx%dx%d_m%d' % s[:4] + ('_to_%dx%d' % s[5:] if len(s) > 5 else ''))
def test_tps_backward_in_node_mode_vs_oracle(shape):
    """Shapes for which the tile kernels evaluate the spline on tile nodes (>= 0.5 Mpix, or >= 25 control points): forward and
    backward must see the SAME coordinates (the sampler backward is compared on the forward's x, y: one flipped corner would
    show as an O(1) error), and grad_T -- computed as the adjoint of the node interpolation -- must equal the oracle's
    sum_pix grad * basis within rel 1e-4, element-wise too; the chain to the offsets through W^-T likewise."""
    from coupe.dvsg_b200 import _lib, ops
    b, h, w, m, amp = shape[:5]
    oh, ow = shape[5:] if len(shape) > 5 else (h, w)      # output size != input size, ragged against the 32 x 8 tiles
    assert _lib.load().dvsg_tps_coords_mode(h, w, 3, oh, ow, m * m, 0) == 1
    rng = np.random.default_rng(h + m)
    u = smooth_image(rng, b, h, w, 3, period=48.0)
    coord = tiled_mesh(m, m, b)
    vec = rng.uniform(-amp, amp, coord.shape).astype(np.float32)
    g_out = rng.standard_normal((b, oh, ow, 3)).astype(np.float32)
    gx_in = rng.standard_normal(b * oh * ow).astype(np.float32)
    gy_in = rng.standard_normal(b * oh * ow).astype(np.float32)
    U, C_ = cu(u), cu(coord)
    T = ops.tps_solve(C_, C_ + cu(vec))
    _, x, y, _ = ops.tps_warp_fwd(U, C_, T, (oh, ow))
    gU, gT, gxs, gys = ops.tps_warp_bwd(U, C_, T, (oh, ow), cu(g_out), cu(gx_in), cu(gy_in), want_grid_grad=True)
    gvec = ops.tps_solve_bwd(C_, gT)
    x, y = x.cpu().numpy(), y.cpu().numpy()
    r_gim, r_gx, r_gy = O.tps_interpolate_bwd(u, x, y, oh, ow, g_out)
    r_gx, r_gy = r_gx.reshape(-1) + gx_in, r_gy.reshape(-1) + gy_in
    assert rel(gU.cpu().numpy(), r_gim) <= 1e-4
    assert rel(gxs.cpu().numpy(), r_gx) <= 1e-4 and rel(gys.cpu().numpy(), r_gy) <= 1e-4
    r_gT = O.tps_grid_bwd(coord.astype(np.float64), oh, ow, gxs.cpu().numpy(), gys.cpu().numpy(), dtype=np.float64)
    print('%s: node-mode grad_T rel %.2e' % (shape, rel(gT.cpu().numpy(), r_gT)))
    assert rel(gT.cpu().numpy(), r_gT) <= 1e-4
    ok, worst, at = close_elementwise(gT.cpu().numpy(), r_gT, rtol=1e-4, atol_frac=1e-5)
    assert ok, (worst, at)
    _, w_inv = O.tps_solve(coord.astype(np.float64), (coord + vec).astype(np.float64), dtype=np.float64, return_inverse=True)
    assert rel(gvec.cpu().numpy(), O.tps_solve_bwd(w_inv, gT.cpu().numpy().astype(np.float64))) <= 1e-4
    # the per-pixel evaluation (DVSG_FLAG_TPS_EXACT in both calls) passes the same stage-wise checks on ITS coordinates; the two
    # grad_T differ by the slope jumps of the few pixels whose sampling corner differs (white-noise grad_out: ~1 % here)
    _, xe, ye, _ = ops.tps_warp_fwd(U, C_, T, (oh, ow), flags=2)
    _, gTe, gxe, gye = ops.tps_warp_bwd(U, C_, T, (oh, ow), cu(g_out), cu(gx_in), cu(gy_in), want_grid_grad=True, flags=2)
    _, e_gx, e_gy = O.tps_interpolate_bwd(u, xe.cpu().numpy(), ye.cpu().numpy(), oh, ow, g_out)
    assert rel(gxe.cpu().numpy(), e_gx.reshape(-1) + gx_in) <= 1e-4 and rel(gye.cpu().numpy(), e_gy.reshape(-1) + gy_in) <= 1e-4
    print('   node vs per-pixel evaluation: grad_T differs by %.2e (corner flips)' % rel(gT.cpu().numpy(), gTe.cpu().numpy()))


# ---- backward of ProjectiveTransformer / AffineTransformer (VERDICT r1 "missing" 6; no reference caller differentiates them) ----
@pytest.mark.parametrize('name', ['projective_grad', 'projective_grad_c18', 'affine_grad'])
def test_transformer_backward_vs_reference_golden(name):
    """Fixtures: the reference's transform() executed unmodified with autograd over its op sequence
    (make_golden.transformer_grad_cases).  Grid stage through the raw C ABI: rel <= 1e-5 (fp64 partial sums); end to end
    through the drop-in classes: max-norm rel <= 1e-4 and element-wise |a-b| <= 1e-4|b| + 2e-5 max|b|."""
    from coupe.dvsg_b200 import _lib
    from coupe.dvsg_b200.spatial_transformer import AffineTransformer, ProjectiveTransformer
    lib = _lib.load()
    g = load_golden(name)
    proj = name.startswith('projective')
    osz = [int(v) for v in g['out_size']]
    b = g['im'].shape[0]
    th, gx, gy = cu(g['theta']), cu(g['g_x']), cu(g['g_y'])
    gt = torch.full_like(th, float('nan'))
    rc = lib.dvsg_homography_grid_bwd(th.data_ptr(), gx.data_ptr(), gy.data_ptr(), 1 if proj else 0, gt.data_ptr(), b, osz[0], osz[1], 0)
    assert rc == 0, lib.dvsg_last_error()
    torch.cuda.synchronize()
    assert rel(gt.cpu().numpy(), g['grad_theta_grid']) <= 1e-5
    tr = (ProjectiveTransformer if proj else AffineTransformer)(osz)
    im, theta = cu(g['im']).requires_grad_(True), cu(g['theta']).requires_grad_(True)
    out = tr.transform(im, theta)
    assert np.abs(out.detach().cpu().numpy() - g['out']).max() <= 1e-5
    (out * cu(g['g_out'])).sum().backward()
    for nm, got, want in (('grad_im', im.grad, g['grad_im']), ('grad_theta', theta.grad, g['grad_theta'])):
        assert rel(got.cpu().numpy(), want) <= 1e-4, nm
        ok, worst, at = close_elementwise(got.cpu().numpy(), want)
        assert ok, (nm, worst, at)
    # the coordinates themselves are differentiable too (_transform)
    theta2 = cu(g['theta']).requires_grad_(True)
    xs, ys = tr._transform(cu(g['im']), theta2)
    ((xs * gx).sum() + (ys * gy).sum()).backward()
    assert rel(theta2.grad.cpu().numpy(), g['grad_theta_grid']) <= 1e-5


@pytest.mark.parametrize('case', [(2, 72, 128, 3, (72, 128), True), (2, 40, 64, 18, (36, 60), True), (3, 50, 70, 3, (64, 96), False)],
                         ids=['projective_c3_tile_path', 'projective_c18_wide_path', 'affine_resize'])
def test_transformer_backward_vs_oracle_seeded(case):
    """Seeded, larger than the fixtures, against oracle.homography_transform_bwd (fp64 grid-stage chain on the oracle's own
    sampler backward): rel <= 1e-4 of the max-norm; theta as model.py:161-163 draws it."""
    from coupe.dvsg_b200.spatial_transformer import AffineTransformer, ProjectiveTransformer
    b, h, w, c, osz, proj = case
    rng = np.random.default_rng(h * w + c)
    im = smooth_image(rng, b, h, w, c)
    if proj:
        theta = (rng.uniform(-1, 1, (b, 8)) * np.array([0.1, 0.1, 0.5, 0.1, 0.1, 0.5, 0.1, 0.1]) + np.array([1.0, 0, 0, 0, 1.0, 0, 0, 0])).astype(np.float32)
    else:
        theta = (rng.uniform(-0.2, 0.2, (b, 6)) + np.array([1.0, 0, 0, 0, 1.0, 0])).astype(np.float32)
    g_out = rng.standard_normal((b, osz[0], osz[1], c)).astype(np.float32)
    tr = (ProjectiveTransformer if proj else AffineTransformer)(list(osz))
    I, TH = cu(im).requires_grad_(True), cu(theta).requires_grad_(True)
    (tr.transform(I, TH) * cu(g_out)).sum().backward()
    r_gim, r_gth = O.homography_transform_bwd(im, theta, osz, g_out, proj)
    assert rel(I.grad.cpu().numpy(), r_gim) <= 1e-4
    assert rel(TH.grad.cpu().numpy(), r_gth) <= 1e-4
    # only the input needs a gradient: theta's chain is skipped, the result is the same
    I2 = cu(im).requires_grad_(True)
    (tr.transform(I2, cu(theta)) * cu(g_out)).sum().backward()
    assert rel(I2.grad.cpu().numpy(), I.grad.cpu().numpy()) <= 1e-6      # (float atomics in the generic kernel: the order may differ)


# ---- gradient w.r.t. the control-point positions (VERDICT r1 "missing" 6; no reference caller takes it) ----
@pytest.mark.parametrize('name', ['tps_coord_grad', 'tps2_coord_grad'])
def test_tps_coord_gradient_vs_reference_golden(name):
    """Stage-wise through the raw C ABI (dvsg_tps_coord_bwd on the oracle's own T, grad_T and coordinate gradients: no
    corner can flip): rel <= 1e-5 against oracle.tps_coord_bwd.  End to end through the drop-ins with coord.requires_grad
    against the reference's own fp64 run: rel <= 1e-4 (max-norm) for coord, the second argument and the image."""
    from coupe.dvsg_b200 import _lib
    from coupe.dvsg_b200.ThinPlateSpline import ThinPlateSpline
    from coupe.dvsg_b200.ThinPlateSpline2 import ThinPlateSpline2
    lib = _lib.load()
    g = load_golden(name)
    osz = [int(v) for v in g['out_size']]
    coord, variant = g['coord'], int(g['variant'])
    b, pn = coord.shape[:2]
    target = (coord + g['second']).astype(np.float32) if variant == 1 else g['second']
    t, w_inv = O.tps_solve(coord, target, return_inverse=True)
    x, y = O.tps_grid(t, coord, osz[0], osz[1])
    _, gx, gy = O.tps_interpolate_bwd(g['u'], x, y, osz[0], osz[1], g['g_out'])
    gx, gy = (gx + g['g_x']).astype(np.float32), (gy + g['g_y']).astype(np.float32)
    g_t = O.tps_grid_bwd(coord, osz[0], osz[1], gx, gy)
    want = O.tps_coord_bwd(coord, t, w_inv, g_t, osz[0], osz[1], gx, gy)
    C_ = cu(coord)
    nbytes = lib.dvsg_tps_prepare_workspace_bytes(b, pn, pn * 2)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    assert lib.dvsg_tps_prepare(C_.data_ptr(), pn * 2, b, pn, ws.data_ptr(), nbytes, 0) == 0, lib.dvsg_last_error()
    got = torch.full((b, pn, 2), float('nan'), device=DEV)
    T_, gT_, gx_, gy_ = cu(t), cu(g_t), cu(gx), cu(gy)
    rc = lib.dvsg_tps_coord_bwd(C_.data_ptr(), pn * 2, T_.data_ptr(), gT_.data_ptr(), gx_.data_ptr(), gy_.data_ptr(), got.data_ptr(),
                                b, osz[0], osz[1], pn, ws.data_ptr(), nbytes, 0)
    assert rc == 0, lib.dvsg_last_error()
    torch.cuda.synchronize()
    assert rel(got.cpu().numpy(), want) <= 1e-5
    # solve part alone (grad_x = grad_y = null)
    rc = lib.dvsg_tps_coord_bwd(C_.data_ptr(), pn * 2, T_.data_ptr(), gT_.data_ptr(), 0, 0, got.data_ptr(), b, osz[0], osz[1], pn, ws.data_ptr(), nbytes, 0)
    assert rc == 0, lib.dvsg_last_error()
    assert rel(got.cpu().numpy(), O.tps_coord_bwd(coord, t, w_inv, g_t, osz[0], osz[1], None, None)) <= 1e-5
    # end to end
    U = cu(g['u']).requires_grad_(True)
    Cg = cu(coord).requires_grad_(True)
    S = cu(g['second']).requires_grad_(True)
    out, xs, ys = (ThinPlateSpline if variant == 1 else ThinPlateSpline2)(U, Cg, S, osz)
    ((out * cu(g['g_out'])).sum() + (xs * cu(g['g_x'])).sum() + (ys * cu(g['g_y'])).sum()).backward()
    print('coord-grad rel vs fp64 reference run:', rel(Cg.grad.cpu().numpy(), g['grad_coord64']))
    assert rel(Cg.grad.cpu().numpy(), g['grad_coord64']) <= 1e-4
    assert rel(S.grad.cpu().numpy(), g['grad_second64']) <= 1e-4
    assert rel(U.grad.cpu().numpy(), g['grad_u64']) <= 1e-4


def test_tps_coord_gradient_of_a_mesh_shared_by_the_batch():
    """coord of shape [pn, 2] (one mesh for every frame, as model.py:68 builds it) with requires_grad: its gradient is the sum
    over frames of the per-frame gradients (same call with the mesh repeated per frame)."""
    from coupe.dvsg_b200.ThinPlateSpline import ThinPlateSpline
    rng = np.random.default_rng(3)
    b, h, w = 3, 40, 56
    mesh = (tiled_mesh(4, 4, 1)[0] + rng.uniform(-0.05, 0.05, (16, 2))).astype(np.float32)
    vec = rng.uniform(-0.1, 0.1, (b, 16, 2)).astype(np.float32)
    U = cu(smooth_image(rng, b, h, w, 3))
    g_out = cu(rng.standard_normal((b, h, w, 3)).astype(np.float32))
    C1 = cu(mesh).requires_grad_(True)
    out, _, _ = ThinPlateSpline(U, C1, cu(vec), [h, w])
    (out * g_out).sum().backward()
    C2 = cu(np.tile(mesh[None], (b, 1, 1))).requires_grad_(True)
    out2, _, _ = ThinPlateSpline(U, C2, cu(vec), [h, w])
    (out2 * g_out).sum().backward()
    assert torch.equal(out, out2) or float((out - out2).abs().max()) <= 1e-5
    assert rel(C1.grad.cpu().numpy(), C2.grad.sum(dim=0).cpu().numpy()) <= 1e-4
    # against the oracle
    _, _, want = O.thin_plate_spline_bwd(U.cpu().numpy(), np.tile(mesh[None], (b, 1, 1)), vec, (h, w), g_out.cpu().numpy(), want_coord=True)
    assert rel(C2.grad.cpu().numpy(), want) <= 2e-4


def test_two_level_node_evaluation_forward_and_backward_see_the_same_coordinates():
    """16 x 16 mesh at 1080p: both tile kernels take the two-level node evaluation (tile_common.cuh).  The sampler backward is
    compared on the FORWARD's x, y (one corner seen differently by the two kernels would show as an O(1) error in grad_xs),
    and grad_T -- the adjoint of the single-level node interpolation applied to the coordinate gradients -- against the
    oracle's sum over pixels, accumulated in row bands (the full basis would be 4 GB)."""
    from coupe.dvsg_b200 import _lib, ops
    b, h, w, m, amp = 1, 1080, 1920, 16, 0.03
    assert _lib.load().dvsg_tps_coords_mode(h, w, 3, h, w, m * m, 0) == 1
    rng = np.random.default_rng(16)
    u = smooth_image(rng, b, h, w, 3, period=48.0)
    coord = tiled_mesh(m, m, b)
    vec = rng.uniform(-amp, amp, coord.shape).astype(np.float32)
    g_out = rng.standard_normal((b, h, w, 3)).astype(np.float32)
    U, C_ = cu(u), cu(coord)
    T = ops.tps_solve(C_, C_ + cu(vec))
    _, x, y, _ = ops.tps_warp_fwd(U, C_, T, (h, w))
    gU, gT, gxs, gys = ops.tps_warp_bwd(U, C_, T, (h, w), cu(g_out), want_grid_grad=True)
    x, y = x.cpu().numpy(), y.cpu().numpy()
    r_gim, r_gx, r_gy = O.tps_interpolate_bwd(u, x, y, h, w, g_out)
    assert rel(gxs.cpu().numpy(), r_gx.reshape(-1)) <= 1e-4 and rel(gys.cpu().numpy(), r_gy.reshape(-1)) <= 1e-4
    assert rel(gU.cpu().numpy(), r_gim) <= 1e-4
    gx, gy = gxs.cpu().numpy().astype(np.float64).reshape(h, w), gys.cpu().numpy().astype(np.float64).reshape(h, w)
    r_gT = np.zeros((2, m * m + 3))
    for r0 in range(0, h, 60):
        basis = O.tps_basis(coord.astype(np.float64), h, w, dtype=np.float64, rows=(r0, r0 + 60))[0]      # [N, 60 * w]
        r_gT[0] += basis @ gx[r0:r0 + 60].reshape(-1)
        r_gT[1] += basis @ gy[r0:r0 + 60].reshape(-1)
    print('two-level shape: grad_T rel %.2e' % rel(gT.cpu().numpy()[0], r_gT))
    assert rel(gT.cpu().numpy()[0], r_gT) <= 1e-4
