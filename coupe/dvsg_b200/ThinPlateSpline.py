"""Drop-in for the reference's ThinPlateSpline.py (same function name, argument order and
return arity: /root/reference/ThinPlateSpline.py:4,168-170), backed by the sm_100a kernels.

    from coupe.dvsg_b200.ThinPlateSpline import ThinPlateSpline as stn      # model.py:9
    output, x, y = stn(U, coord, vector, [h, w])
"""
from . import ops


def _check_mesh(coord, vector):
    if coord.dim() != 3 or vector.dim() != 3 or coord.shape[2] != 2 or tuple(coord.shape) != tuple(vector.shape):
        raise ValueError('coord and vector must both have shape [num_batch, num_point, 2], got %s and %s'
                         % (tuple(coord.shape), tuple(vector.shape)))


def ThinPlateSpline(U, coord, vector, out_size, return_grid=True):
    """Thin Plate Spline Spatial Transformer Layer.

    U      : float [num_batch, height, width, num_channels] CUDA tensor (or DLPack producer).
    coord  : float [num_batch, num_point, 2] control points in [-1, 1], (x, y) order.  A
             batch-stride-0 view (`mesh.expand(B, -1, -1)`) or a [num_point, 2] tensor marks the
             mesh as shared by the batch (the only way the reference ever calls it, model.py:68).
    vector : float [num_batch, num_point, 2] offsets of the control points.
    out_size : (height, width) of the output.
    Returns (output [B,h,w,C], x [B*h*w], y [B*h*w]); x, y are the normalised source
    coordinates of every output pixel, exactly the reference's 2nd and 3rd results.
    `return_grid=False` (extension) skips materialising x, y -- inference callers such as
    eval.py:110 only fetch the warped frame -- and returns (output, None, None).
    Differentiable w.r.t. U and vector (and through x, y); coord is a constant.
    """
    from ._tensors import as_cuda_f32
    U = as_cuda_f32(U, 'U')
    vector = as_cuda_f32(vector, 'vector', like=U)
    coord_t = ops._as_mesh(coord, U)
    if coord_t.dim() == 2:
        coord_t = coord_t.unsqueeze(0).expand(U.shape[0], -1, -1)
    _check_mesh(coord_t, vector)
    # target = coord + vector (ThinPlateSpline.py:161) is formed inside the solve: same fp32 add, one launch less
    return ops.thin_plate_spline(U, coord_t, vector, out_size, want_grid=return_grid, offsets=True)


def ThinPlateSplineWithMask(U, coord, vector, out_size, return_grid=True):
    """One pass that yields both `stn(U, ...)` and `stn(ones_like(U), ...)` -- the pair every
    training call site computes back to back (model.py:81-85,120-121).
    Returns (output, mask, x, y); mask has U's shape (stride-0 over channels)."""
    from ._tensors import as_cuda_f32
    U = as_cuda_f32(U, 'U')
    vector = as_cuda_f32(vector, 'vector', like=U)
    coord_t = ops._as_mesh(coord, U)
    if coord_t.dim() == 2:
        coord_t = coord_t.unsqueeze(0).expand(U.shape[0], -1, -1)
    _check_mesh(coord_t, vector)
    return ops.thin_plate_spline_with_mask(U, coord_t, vector, out_size, want_grid=return_grid, offsets=True)
