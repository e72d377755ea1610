"""GPU parity tests of the N3 row: sparse TPS evaluation at feature points and the masked MSE, through
the C ABI.  Oracle = oracle/dvsg_oracle.py (restatement of trainer.py:232-243, 363-386).

Stated tolerances: sparse evaluation == gather of the dense grid, BIT-EXACT (same arithmetic); losses rel <= 1e-5
against the float64-accumulating oracle; gradients rel <= 1e-4 of the max-norm against float64 torch autograd of the
same formula."""
import numpy as np
import pytest
import torch

from helpers import smooth_image, tiled_mesh
from oracle import dvsg_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def make_surf(rng, b, p, h, w, n_pad):
    surf = np.zeros((b, 2, p, 2), np.int32)
    surf[:, :, :, 0] = rng.integers(0, w, (b, 2, p))
    surf[:, :, :, 1] = rng.integers(0, h, (b, 2, p))
    if n_pad:       # padded features: stable index h*w -> the appended -1 entry; unstable (0, 0) -> (-1, -1)
        surf[:, 0, -n_pad:, :] = 0
        surf[:, 1, -n_pad:, 0] = 0
        surf[:, 1, -n_pad:, 1] = h
    return surf


@pytest.mark.parametrize('m', [4, 5, 7])
def test_sparse_tps_evaluation_equals_dense_grid_gather(m):
    from coupe.dvsg_b200 import losses, ops
    b, h, w, p = 3, 96, 160, 205
    rng = np.random.default_rng(m)
    coord = tiled_mesh(m, m, b)
    coord[1] += rng.uniform(-0.04, 0.04, coord[1].shape).astype(np.float32)      # one irregular mesh in the batch
    vec = rng.uniform(-0.1, 0.1, coord.shape).astype(np.float32)
    C_, S = cu(coord), cu(vec)
    T = ops.tps_solve(C_, C_ + S)
    U = cu(smooth_image(rng, b, h, w, 3))
    _, x, y, _ = ops.tps_warp_fwd(U, C_, T, (h, w), want_grid=True)
    idx = rng.integers(0, h * w + 1, (b, p)).astype(np.int32)
    idx[:, :4] = [0, w - 1, h * w - 1, h * w]
    sx, sy = losses.tps_eval_points(C_, C_ + S, cu(idx), (h, w))
    xd = torch.cat([x.reshape(b, -1), -torch.ones((b, 1), device=DEV)], 1)
    yd = torch.cat([y.reshape(b, -1), -torch.ones((b, 1), device=DEV)], 1)
    assert torch.equal(sx, torch.gather(xd, 1, cu(idx).long()))
    assert torch.equal(sy, torch.gather(yd, 1, cu(idx).long()))
    r_x, r_y = O.tps_grid(O.tps_solve(coord, coord + vec), coord, h, w)[:2]
    rx = np.concatenate([np.asarray(r_x).reshape(b, -1), -np.ones((b, 1), np.float32)], 1)
    assert np.abs(sx.cpu().numpy() - np.take_along_axis(rx, idx.astype(np.int64), 1)).max() <= 2e-5


def test_surf_loss_sparse_dense_and_oracle_agree_with_gradients():
    from coupe.dvsg_b200 import losses
    from coupe.dvsg_b200.ThinPlateSpline import ThinPlateSpline
    b, h, w, p = 4, 72, 128, 150
    rng = np.random.default_rng(5)
    coord = tiled_mesh(4, 4, b)
    vec = rng.uniform(-0.1, 0.1, coord.shape).astype(np.float32)
    surf = make_surf(rng, b, p, h, w, n_pad=17)
    max_dim = np.array([p - 17, p - 17, 0, p], np.float32)                      # one frame without features: div_no_nan
    U = cu(smooth_image(rng, b, h, w, 3))
    C_ = cu(coord)
    v1 = cu(vec).requires_grad_(True)
    _, x, y = ThinPlateSpline(U, C_, v1, [h, w])
    dense = losses.get_surf_loss(cu(surf), x, y, cu(max_dim), b, w, h)
    dense.backward()
    v2 = cu(vec).requires_grad_(True)
    sparse = losses.get_surf_loss_sparse(cu(surf), C_, C_ + v2, cu(max_dim), w, h)
    sparse.backward()
    assert torch.equal(dense.detach(), sparse.detach())                          # same numbers entering the same torch ops
    ref, _ = O.surf_loss(surf, x.detach().cpu().numpy(), y.detach().cpu().numpy(), max_dim, b, w, h)
    assert abs(float(sparse.detach()) - ref) <= 1e-5 * max(abs(ref), 1e-6)
    g1, g2 = v1.grad.cpu().numpy(), v2.grad.cpu().numpy()
    assert np.abs(g1 - g2).max() <= 1e-4 * np.abs(g1).max()


@pytest.mark.parametrize('shape', [(3, 72, 128, 3, 3), (2, 33, 47, 3, 1), (2, 288, 512, 3, 3), (1, 16, 16, 1, 1)])
def test_masked_mse_vs_oracle_and_autograd(shape):
    from coupe.dvsg_b200 import losses
    b, h, w, c, mc = shape
    rng = np.random.default_rng(h + w)
    pred = rng.random((b, h, w, c), dtype=np.float32)
    gt = rng.random((b, h, w, c), dtype=np.float32)
    mask = (rng.random((b, h, w, mc)) > 0.3).astype(np.float32) * rng.random((b, h, w, mc), dtype=np.float32)
    if b > 1:
        mask[1] = 0.0                                                            # empty mask: div_no_nan -> 0, zero gradients
    P, G, M = (cu(a).requires_grad_(True) for a in (pred, gt, mask))
    loss = losses.masked_MSE(P, G, M)
    again = losses.masked_MSE(cu(pred), cu(gt), cu(mask))
    assert torch.equal(loss.detach(), again)                                     # deterministic reduction order
    ref, _, _ = O.masked_mse(pred, gt, mask)
    # the oracle sums the mask over ITS OWN elements (tf.reduce_sum(mask)), like the kernel
    assert abs(float(loss.detach()) - ref) <= 1e-5 * max(abs(ref), 1e-12)
    (loss * 1.7).backward()
    p64, g64, m64 = (torch.from_numpy(a).double().requires_grad_(True) for a in (pred, gt, mask))
    sq = ((p64 * m64 - g64 * m64) ** 2).sum(dim=(1, 2, 3))
    ms = m64.sum(dim=(1, 2, 3))
    l64 = torch.where(ms != 0, sq / torch.where(ms != 0, ms, torch.ones_like(ms)), torch.zeros_like(sq)).mean() * 1.7
    l64.backward()
    for got, want in ((P.grad, p64.grad), (G.grad, g64.grad), (M.grad, m64.grad)):
        want = want.numpy()
        assert np.abs(got.cpu().numpy() - want).max() <= 1e-4 * max(np.abs(want).max(), 1e-12)


def test_temporal_loss_composes_flow_warp_and_masked_mse():
    from coupe.dvsg_b200 import losses
    b, h, w = 2, 64, 96
    rng = np.random.default_rng(9)
    pred, gt = smooth_image(rng, b, h, w, 3), smooth_image(rng, b, h, w, 3)
    mp = np.ones((b, h, w, 3), np.float32)
    mg = (rng.random((b, h, w, 3)) > 0.2).astype(np.float32)
    flow = rng.uniform(-3, 3, (b, h, w, 2)).astype(np.float32)
    got = losses.temporal_loss(cu(pred), cu(gt), cu(mp), cu(mg), cu(flow), h, w)
    pw, mw = O.tf_warp(pred, flow, h, w), O.tf_warp(mp, flow, h, w)
    ref, _, _ = O.masked_mse(pw, gt, mw * mg)
    assert abs(float(got) - ref) <= 1e-5 * abs(ref)


# ---- against fixtures generated by the reference's own method bodies (tests/golden/make_golden.py::loss_cases) ----
def _close(got, want, rtol=1e-4, atol_frac=1e-6):
    want = np.asarray(want)
    return bool(np.all(np.abs(got - want) <= rtol * np.abs(want) + atol_frac * max(np.abs(want).max(), 1e-30)))


def test_masked_mse_vs_reference_golden():
    """Trainer.masked_MSE (trainer.py:232-243) executed unmodified over the TF1 shim: loss rel <= 1e-5, the three
    gradients element-wise |a-b| <= 1e-4|b| + 1e-6 max|b|."""
    from conftest import load_golden
    from coupe.dvsg_b200 import losses
    g = load_golden('loss_masked_mse')
    P, G, M = (cu(g[k]).requires_grad_(True) for k in ('pred', 'gt', 'mask'))
    loss = losses.masked_MSE(P, G, M)
    assert abs(float(loss.detach()) - float(g['loss'])) <= 1e-5 * abs(float(g['loss']))
    (loss * float(g['grad_scale'])).backward()
    assert _close(P.grad.cpu().numpy(), g['grad_pred'])
    assert _close(G.grad.cpu().numpy(), g['grad_gt'])
    assert _close(M.grad.cpu().numpy(), g['grad_mask'])


def test_temporal_loss_vs_reference_golden():
    """Trainer.temporal_loss (trainer.py:245-250): tf_warp of the prediction and its mask + masked MSE; gradients flow
    into BOTH warped inputs (the image gradient of tf_warp)."""
    from conftest import load_golden
    from coupe.dvsg_b200 import losses
    g = load_golden('loss_temporal')
    h, w = g['pred'].shape[1:3]
    P, MP = cu(g['pred']).requires_grad_(True), cu(g['mask_pred']).requires_grad_(True)
    loss = losses.temporal_loss(P, cu(g['gt']), MP, cu(g['mask_gt']), cu(g['flow']), h, w)
    assert abs(float(loss.detach()) - float(g['loss'])) <= 1e-5 * abs(float(g['loss']))
    loss.backward()
    assert _close(P.grad.cpu().numpy(), g['grad_pred'], atol_frac=1e-5)
    assert _close(MP.grad.cpu().numpy(), g['grad_mask_pred'], atol_frac=1e-5)


def test_surf_loss_vs_reference_golden():
    """Trainer.get_surf_loss (trainer.py:363-386) on dense grids, incl. the sentinel index h*w and a frame with
    max_dim = 0 (div_no_nan)."""
    from conftest import load_golden
    from coupe.dvsg_b200 import losses
    g = load_golden('loss_surf')
    h, w = (int(v) for v in g['hw'])
    b = g['surf'].shape[0]
    X, Y = cu(g['x']).requires_grad_(True), cu(g['y']).requires_grad_(True)
    loss = losses.get_surf_loss(cu(g['surf']), X, Y, cu(g['max_dim']), b, w, h)
    assert abs(float(loss.detach()) - float(g['loss'])) <= 1e-5 * abs(float(g['loss']))
    loss.backward()
    assert _close(X.grad.cpu().numpy(), g['grad_x'])
    assert _close(Y.grad.cpu().numpy(), g['grad_y'])
