"""The consumers of the warp path's outputs in the reference's training graph, on the sm_100a kernels.

    masked_MSE(pred, gt, mask)                      trainer.py:232-243   (method of Trainer in the reference)
    get_surf_loss(surf, x_offset, y_offset, ...)    trainer.py:363-386   same signature, dense grid as in the reference
    get_surf_loss_sparse(surf, coord, target, ...)  same loss, but the spline is evaluated AT the feature points
                                                    (dvsg_tps_eval_points) instead of gathering a dense [B*h*w] grid:
                                                    identical numbers, no 8 B/pixel grid in training
Tensors are CUDA torch tensors; the reductions, the sparse evaluation and their backward passes are CUDA kernels behind
the C ABI (csrc/losses.cu).  The few [B, P, 2]-sized elementwise steps of the surf loss are torch ops (plumbing).
"""
import torch

from . import _lib, ops
from ._tensors import as_cuda_f32, ptr, stream_ptr


# ---- masked MSE ---------------------------------------------------------------------------------------
class _MaskedMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, mask):
        lib = _lib.load()
        B, C = pred.shape[0], pred.shape[-1]
        n_px = pred[0].numel() // C
        mc = mask.shape[-1]
        dev = pred.device
        sq = torch.empty(B, dtype=torch.float32, device=dev)
        ms = torch.empty(B, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        nbytes = lib.dvsg_masked_mse_workspace_bytes(B, n_px * C)
        ws = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.dvsg_masked_mse_fwd(ptr(pred), ptr(gt), ptr(mask), mc, ptr(sq), ptr(ms), ptr(loss), ptr(ws), nbytes, B, n_px, C, stream_ptr(dev))
        _lib.check(rc, 'dvsg_masked_mse_fwd')
        ctx.save_for_backward(pred, gt, mask, sq, ms)
        return loss[0]

    @staticmethod
    def backward(ctx, grad_loss):
        pred, gt, mask, sq, ms = ctx.saved_tensors
        lib = _lib.load()
        B, C = pred.shape[0], pred.shape[-1]
        n_px = pred[0].numel() // C
        need = ctx.needs_input_grad
        gp = torch.empty_like(pred) if need[0] else None
        gg = torch.empty_like(gt) if need[1] else None
        gm = torch.empty_like(mask) if need[2] else None
        with torch.cuda.device(pred.device):
            rc = lib.dvsg_masked_mse_bwd(ptr(pred), ptr(gt), ptr(mask), mask.shape[-1], ptr(sq), ptr(ms), float(grad_loss), ptr(gp), ptr(gg), ptr(gm),
                                         B, n_px, C, stream_ptr(pred.device))
        _lib.check(rc, 'dvsg_masked_mse_bwd')
        return gp, gg, gm


def masked_MSE(pred, gt, mask, name=None):
    """mean over the batch of  sum((pred*mask - gt*mask)^2) / sum(mask)  with tf.div_no_nan (trainer.py:232-243).
    pred, gt [B,H,W,C]; mask [B,H,W,C] or [B,H,W,1].  Returns a 0-d tensor; differentiable w.r.t. all three."""
    pred = as_cuda_f32(pred, 'pred')
    gt = as_cuda_f32(gt, 'gt', like=pred)
    mask = as_cuda_f32(mask, 'mask', like=pred)
    if pred.dim() != 4 or gt.shape != pred.shape:
        raise ValueError('pred and gt must both be [B,H,W,C], got %s and %s' % (tuple(pred.shape), tuple(gt.shape)))
    if tuple(mask.shape[:3]) != tuple(pred.shape[:3]) or mask.dim() != 4 or mask.shape[3] not in (1, pred.shape[3]):
        raise ValueError('mask must be [B,H,W,C] or [B,H,W,1], got %s' % (tuple(mask.shape),))
    if pred.shape[0] == 0 or pred[0].numel() == 0:
        raise ValueError('masked_MSE of an empty batch')
    return _MaskedMSE.apply(pred, gt, mask)


def temporal_loss(pred, gt, mask_pred, mask_gt, of, h, w, name=None):
    """trainer.py:245-250: warp pred and its mask with the flow, then the masked MSE."""
    from .warp_with_optical_flow import tf_warp
    pred_warped = tf_warp(pred, of, h, w)
    mask_pred_warped = tf_warp(mask_pred, of, h, w)
    return masked_MSE(pred_warped, gt, mask_pred_warped * mask_gt, name)


# ---- sparse TPS evaluation ----------------------------------------------------------------------------
class _TpsEvalPoints(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coord, target, idx, oh, ow):
        lib = _lib.load()
        T = ops.tps_solve(coord, target)
        B, P = idx.shape
        cbuf, cstride, pn = ops._mesh_args(coord, B, T.shape[2] - 3)
        x = torch.empty((B, P), dtype=torch.float32, device=T.device)
        y = torch.empty((B, P), dtype=torch.float32, device=T.device)
        with torch.cuda.device(T.device):
            rc = lib.dvsg_tps_eval_points(ptr(cbuf), cstride, ptr(T), ptr(idx), ptr(x), ptr(y), B, oh, ow, pn, P, stream_ptr(T.device))
        _lib.check(rc, 'dvsg_tps_eval_points')
        ctx.save_for_backward(coord, idx)
        ctx.dims = (oh, ow, pn)
        return x, y

    @staticmethod
    def backward(ctx, gx, gy):
        coord, idx = ctx.saved_tensors
        oh, ow, pn = ctx.dims
        lib = _lib.load()
        B, P = idx.shape
        cbuf, cstride, _ = ops._mesh_args(coord, B, pn)
        gT = torch.empty((B, 2, pn + 3), dtype=torch.float32, device=idx.device)
        gx, gy = gx.contiguous(), gy.contiguous()      # keep the (possibly freshly copied) buffers alive across the call
        with torch.cuda.device(idx.device):
            rc = lib.dvsg_tps_eval_points_bwd(ptr(cbuf), cstride, ptr(idx), ptr(gx), ptr(gy), ptr(gT), B, oh, ow, pn, P,
                                              stream_ptr(idx.device))
        _lib.check(rc, 'dvsg_tps_eval_points_bwd')
        return None, ops.tps_solve_bwd(coord, gT), None, None, None


def tps_eval_points(coord, target, idx, out_size):
    """x_s, y_s of ThinPlateSpline(U, coord, target - coord, out_size) at the flat pixel indices idx [B, P] (int32,
    idx = col + row*w; idx == h*w yields -1, the extra entry trainer.py:364-365 appends).  Bit-identical to gathering
    the dense grid the warp returns.  Differentiable w.r.t. target."""
    target = as_cuda_f32(target, 'target')
    coord = ops._as_mesh(coord, target)
    if idx.dtype != torch.int32 or not idx.is_cuda or idx.dim() != 2 or idx.shape[0] != target.shape[0]:
        raise ValueError('idx must be an int32 CUDA tensor [B, P]')
    return _TpsEvalPoints.apply(coord, target, idx.contiguous(), int(out_size[0]), int(out_size[1]))


def _surf_terms(surf, w, h):
    # trainer.py:367-378
    unstab = surf[:, 0, :, :].float()
    unstab_norm = torch.stack([(unstab[:, :, 0] / (w - 1)) * 2 - 1, (unstab[:, :, 1] / (h - 1)) * 2 - 1], dim=2)
    stab = surf[:, 1, :, :]
    idx = (stab[:, :, 0].float() + stab[:, :, 1].float() * w).to(torch.int32)
    return unstab_norm, idx


def get_surf_loss(surf, x_offset, y_offset, max_dim_per_batch, batch_size, w, h):
    """trainer.py:363-386 with the reference's signature: x_offset, y_offset are the dense flat grids ThinPlateSpline
    returned.  surf [B, 2, P, 2] (unstable / stable feature positions in pixels), max_dim_per_batch [B]."""
    x_offset = torch.cat([x_offset.reshape(batch_size, -1), -torch.ones((batch_size, 1), device=x_offset.device)], dim=1)
    y_offset = torch.cat([y_offset.reshape(batch_size, -1), -torch.ones((batch_size, 1), device=y_offset.device)], dim=1)
    unstab_norm, idx = _surf_terms(surf, w, h)
    tx = torch.gather(x_offset, 1, idx.long())
    ty = torch.gather(y_offset, 1, idx.long())
    return _surf_mse(torch.stack([tx, ty], dim=2), unstab_norm, max_dim_per_batch)


def get_surf_loss_sparse(surf, coord, target, max_dim_per_batch, w, h):
    """The same loss without the dense grid: the spline is evaluated at the stable feature points."""
    unstab_norm, idx = _surf_terms(surf, w, h)
    tx, ty = tps_eval_points(coord, target, idx, (h, w))
    return _surf_mse(torch.stack([tx, ty], dim=2), unstab_norm, max_dim_per_batch)


def _surf_mse(transformed, unstab_norm, max_dim_per_batch):
    mse = ((transformed - unstab_norm) ** 2).sum(dim=(1, 2))                 # :382
    md = max_dim_per_batch.to(mse.dtype)
    mse = torch.where(md != 0, mse / torch.where(md != 0, md, torch.ones_like(md)), torch.zeros_like(mse))   # tf.div_no_nan, :383
    return mse.mean()                                                          # :385
