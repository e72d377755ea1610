"""Drop-in for the live part of the reference's spatial_transformer.py:
_meshgrid / _repeat / _interpolate / bilinear_interp (:460-563) and the two transformer
classes built on them that have callers or trivially share the kernel (ProjectiveTransformer
:364-452 -- model.py:156-167 -- and AffineTransformer :5-91).

Not provided: bicubic_interp (:565-673) is unreachable and broken in the reference (NameError
at :633); ElasticTransformer (:93-362) has no caller.
"""
import torch

from . import ops


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
        sampling are fused in one kernel; differentiation is not provided (the only reference
        caller feeds random constants, model.py:156-167)."""
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
