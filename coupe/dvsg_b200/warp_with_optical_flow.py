"""Drop-in for the live part of the reference's warp_with_optical_flow.py
(/root/reference/warp_with_optical_flow.py:96-176).  `tf_warp_prev` / `get_pixel_value`
(:3-89) are dead code in the reference (no caller) and are not provided."""
from . import ops

__all__ = ['tf_warp']


def tf_warp(im, flow, out_height, out_width):
    """Dense backward warp out(p) = im(p + flow(p)).

    im   : float [batch, height, width, channels]
    flow : float [batch, height, width, 2], channel 0 = dx, 1 = dy, in pixels
    Returns [batch, out_height, out_width, channels].  Zero outside the frame (1-px fade),
    differentiable w.r.t. im and flow.
    """
    return ops.flow_warp(im, flow, out_height, out_width)
