"""Drop-in for the live part of the reference's spatial_transformer.py:
_meshgrid / _repeat / _interpolate / bilinear_interp (:460-563) and the two transformer
classes built on them that have callers or trivially share the kernel (ProjectiveTransformer
:364-452 -- model.py:156-167 -- and AffineTransformer :5-91).

ElasticTransformer (:93-362), the reference's second TPS formulation (no caller in the model; SURVEY.md
8(f) N4 / Appendix A), is provided on the prepared-inverse path: L^-1 once at construction, per
call one small solve, the dense grid and bilinear_interp.

Not provided: bicubic_interp (:565-673) is unreachable and broken in the reference (NameError
at :633).
"""
import math

import torch

from . import _lib, ops
from ._tensors import as_cuda_f32, ptr, stream_ptr


def _meshgrid(out_size, device=None):
    """Flat [x_t ; y_t ; 1] sampling grid of 3*h*w floats (spatial_transformer.py:460-482)."""
    return ops.st_meshgrid(out_size, device)


def _repeat(x, n_repeats):
    """spatial_transformer.py:485-487 (index helper of the TF formulation; kept for API parity)."""
    return x.reshape(-1, 1).repeat(1, int(n_repeats)).reshape(-1)


def bilinear_interp(im, x, y, out_size):
    """Bilinear sampling with 1-px zero padding and (W-1)/2 scaling (spatial_transformer.py:496-563).
    im [B,H,W,C]; x, y flat normalised [B*h*w] -> [B*h*w, C]."""
    return ops.bilinear_interp(im, x, y, out_size)


def bicubic_interp(im, x, y, out_size):
    raise NotImplementedError("the reference's bicubic_interp cannot run (NameError at spatial_transformer.py:633) "
                              "and has no caller; only 'bilinear' is implemented")


def _interpolate(im, x, y, out_size, method):
    """spatial_transformer.py:489-494."""
    if method == 'bilinear':
        return bilinear_interp(im, x, y, out_size)
    if method == 'bicubic':
        return bicubic_interp(im, x, y, out_size)
    return None


class _HomographyTransformer(object):
    _projective = True
    param_dim = 8

    def __init__(self, out_size, name=None, interp_method='bilinear', **kwargs):
        self.name = name or type(self).__name__
        self.out_size = out_size
        self.interp_method = interp_method
        if interp_method != 'bilinear':
            bicubic_interp(None, None, None, None)
        self._pixel_grid = None

    @property
    def pixel_grid(self):
        if self._pixel_grid is None:
            self._pixel_grid = _meshgrid(self.out_size)
        return self._pixel_grid

    def transform(self, inp, theta):
        """inp [B,H,W,C], theta [B, param_dim] -> [B, out_h, out_w, C].  Grid generation and
        sampling are fused in one kernel; differentiable w.r.t. inp and theta (the only reference
        caller feeds random constants, model.py:156-167, so the backward is built for completeness:
        dvsg_bilinear_bwd + dvsg_homography_grid_bwd)."""
        return ops.homography_warp(inp, theta, self.out_size, self._projective)

    def _transform(self, inp, theta):
        """(x_s, y_s) flat, as the reference's _transform returns them."""
        _, x, y = ops.homography_warp(inp, theta, self.out_size, self._projective, want_grid=True)
        return x, y


class ProjectiveTransformer(_HomographyTransformer):
    """spatial_transformer.py:364-452: theta [B, 8] (3x3 homography with h33 = 1), div_no_nan."""
    _projective = True
    param_dim = 8

    def __init__(self, out_size, name='SpatialProjectiveTransformer', interp_method='bilinear', **kwargs):
        super(ProjectiveTransformer, self).__init__(out_size, name, interp_method, **kwargs)


class AffineTransformer(_HomographyTransformer):
    """spatial_transformer.py:5-91: theta [B, 6]."""
    _projective = False
    param_dim = 6

    def __init__(self, out_size, name='SpatialAffineTransformer', interp_method='bilinear', **kwargs):
        super(AffineTransformer, self).__init__(out_size, name, interp_method, **kwargs)


# ---- ElasticTransformer ------------------------------------------------------------------------------------
class _ElasticGrid(torch.autograd.Function):
    """theta_abs [B,2,pn] -> (x_s, y_s): theta @ L_inv, then coefficients @ right_mat (spatial_transformer.py:283-296)."""

    @staticmethod
    def forward(ctx, theta_abs, layer):
        lib = _lib.load()
        B, pn = theta_abs.shape[0], layer.num_control_points
        oh, ow = layer.out_size
        dev = theta_abs.device
        coef = torch.empty((B, 2, pn + 3), dtype=torch.float32, device=dev)
        x = torch.empty(B * oh * ow, dtype=torch.float32, device=dev)
        y = torch.empty(B * oh * ow, dtype=torch.float32, device=dev)
        sp, ws = layer._state(dev)
        with torch.cuda.device(dev):
            _lib.check(lib.dvsg_elastic_solve(ptr(theta_abs), ptr(coef), B, pn, ptr(ws), ws.numel(), stream_ptr(dev)), 'dvsg_elastic_solve')
            _lib.check(lib.dvsg_elastic_grid(ptr(sp), ptr(coef), ptr(x), ptr(y), B, oh, ow, pn, stream_ptr(dev)), 'dvsg_elastic_grid')
        ctx.layer, ctx.B = layer, B
        return x, y

    @staticmethod
    def backward(ctx, gx, gy):
        lib = _lib.load()
        layer, B = ctx.layer, ctx.B
        pn = layer.num_control_points
        oh, ow = layer.out_size
        dev = gx.device
        gx, gy = gx.contiguous(), gy.contiguous()
        sp, ws = layer._state(dev)
        g_coef = torch.empty((B, 2, pn + 3), dtype=torch.float32, device=dev)
        g_theta = torch.empty((B, 2, pn), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.dvsg_elastic_grid_bwd(ptr(sp), ptr(gx), ptr(gy), ptr(g_coef), B, oh, ow, pn, stream_ptr(dev)), 'dvsg_elastic_grid_bwd')
            _lib.check(lib.dvsg_elastic_solve_bwd(ptr(g_coef), ptr(g_theta), B, pn, ptr(ws), ws.numel(), stream_ptr(dev)), 'dvsg_elastic_solve_bwd')
        return g_theta, None


class ElasticTransformer(object):
    """Spatial Elastic Transformer Layer with Thin Plate Spline deformations (spatial_transformer.py:93-362): same
    constructor arguments, same `transform(inp, theta, forward=True) -> (output, x_s, y_s)`.

    The regular g x g control mesh and L^-1 are fixed at construction (_initialize_tps, :324-362: basis r2*log(r2) with
    log(0) -> 0, coefficient order (x, y, 1, w_k), the L_mid row as the reference writes it); the [pn+3, h*w] right_mat
    is never materialised.  Differentiable w.r.t. inp and theta."""

    def __init__(self, out_size, param_dim=2 * 16, param_dim_per_side=4, name='SpatialElasticTransformer', interp_method='bilinear', **kwargs):
        num_control_points = int(param_dim / 2)
        assert param_dim == 2 * num_control_points, 'param_dim must be 2 times a square of an integer.'
        self.name = name
        self.param_dim = param_dim
        self.interp_method = interp_method
        self.num_control_points = num_control_points
        self.num_control_points_per_side = param_dim_per_side
        self.out_size = (int(out_size[0]), int(out_size[1]))
        self.grid_size = math.floor(math.sqrt(self.num_control_points))
        assert self.grid_size * self.grid_size == self.num_control_points, 'num_control_points must be a square of an int'
        self.num_pixels = self.out_size[0] * self.out_size[1]
        if interp_method != 'bilinear':
            bicubic_interp(None, None, None, None)
        self._dev_state = {}

    def _state(self, device):
        """(source points [pn,2] (x, y), workspace holding L^-1) on `device`, built on first use."""
        st = self._dev_state.get(device)
        if st is None:
            lib = _lib.load()
            g, pn = self.grid_size, self.num_control_points
            # get_meshgrid(g, g) (:313-322): the x and y rows of the [x; y; 1] grid that _meshgrid builds with the same tf.linspace
            grid = ops.st_meshgrid((g, g), device)
            sp = grid[:2 * pn].reshape(2, pn).t().contiguous()
            nbytes = lib.dvsg_elastic_workspace_bytes(pn)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            with torch.cuda.device(device):
                _lib.check(lib.dvsg_elastic_prepare(ptr(sp), pn, ptr(ws), nbytes, stream_ptr(device)), 'dvsg_elastic_prepare')
            st = self._dev_state[device] = (sp, ws)
        return st

    @property
    def source_points(self):
        """[2, pn] (x row, y row), as the reference keeps it (:131)."""
        dev = next(iter(self._dev_state), torch.device('cuda', torch.cuda.current_device()))
        return self._state(dev)[0].t().contiguous()

    def _abs_theta(self, theta):
        theta = as_cuda_f32(theta, 'theta')
        sp = self._state(theta.device)[0]
        return sp.t().unsqueeze(0) + theta.reshape(-1, 2, self.num_control_points)      # :160-161

    def transform(self, inp, theta, forward=True, **kwargs):
        inp = as_cuda_f32(inp, 'inp')
        if inp.dim() != 4:
            raise ValueError('inp must have shape [batch_size, height, width, num_channels], got %s' % (tuple(inp.shape),))
        theta_abs = self._abs_theta(theta).contiguous()
        if theta_abs.shape[0] != inp.shape[0]:
            raise ValueError('theta holds %d transforms for a batch of %d' % (theta_abs.shape[0], inp.shape[0]))
        x_s, y_s = _ElasticGrid.apply(theta_abs, self)        # forward=False recomputes the same transform in the reference (:174-184)
        output = _interpolate(inp, x_s, y_s, self.out_size, method=self.interp_method)
        return output.reshape(-1, self.out_size[0], self.out_size[1], inp.shape[3]), x_s, y_s

    def get_abs_theta(self, theta):
        """:195-219: absolute control-point targets mapped to [0, 1], reshaped to the mesh."""
        t = self._abs_theta(theta).permute(0, 2, 1)
        xy = torch.clamp((t + 1.0) / 2.0, 0, 1)
        return xy.reshape(-1, self.num_control_points_per_side, self.num_control_points_per_side, 2)

    def get_abs_src_points(self, batch_size):
        """:258-272."""
        sp = self.source_points.unsqueeze(0).permute(0, 2, 1)
        xy = torch.clamp((sp + 1.0) / 2.0, 0, 1)
        return xy.reshape(-1, self.num_control_points_per_side, self.num_control_points_per_side, 2)
