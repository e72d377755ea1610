"""Host-side operators over the C ABI: shape checks, buffer allocation, autograd wiring.

Each public function here is the eager, CUDA-backed equivalent of one group of TF ops in
the reference; the drop-in modules (ThinPlateSpline.py, spatial_transformer.py,
warp_with_optical_flow.py) only rename these to the reference's signatures.
"""
import ctypes

import collections

import torch

from . import _lib
from ._tensors import as_cuda_f32, out_hw, ptr, stream_ptr

_workspaces = {}   # (device index) -> cached uint8 workspace for large TPS systems


def _workspace(device, nbytes):
    if nbytes == 0:
        return None
    key = device.index
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _as_mesh(coord, like):
    """coord -> fp32 CUDA tensor WITHOUT materialising a batch-stride-0 (expanded) view."""
    if not isinstance(coord, torch.Tensor):
        if not hasattr(coord, '__dlpack__'):
            raise TypeError('coord: expected a torch.Tensor or a DLPack-capable array, got %s' % type(coord).__name__)
        coord = torch.from_dlpack(coord)
    if not coord.is_cuda:
        raise ValueError('coord: tensor lives on %s -- this path runs on CUDA only (no CPU fallback)' % coord.device)
    if coord.device != like.device:
        raise ValueError('coord: device %s differs from %s' % (coord.device, like.device))
    return coord if coord.dtype == torch.float32 else coord.float()


def _mesh_args(coord, B, pn_expected=None):
    """coord [B,pn,2] (dense) or [pn,2] / batch-stride-0 view (one mesh shared by the batch,
    as model.py:68 builds it) -> (contiguous tensor, batch stride in floats, pn)."""
    if coord.dim() == 2:
        coord = coord.unsqueeze(0).expand(B, -1, -1)
    if coord.dim() != 3 or coord.shape[2] != 2 or coord.shape[0] != B:
        raise ValueError('coord must have shape [B, num_point, 2] with B=%d, got %s' % (B, tuple(coord.shape)))
    pn = coord.shape[1]
    if pn_expected is not None and pn != pn_expected:
        raise ValueError('coord has %d control points, expected %d' % (pn, pn_expected))
    if B > 0 and (coord.stride(0) == 0 or B == 1):      # a single frame's mesh is "shared" too: prepared solve, one table
        return coord[0].contiguous(), 0, pn
    return coord.contiguous(), pn * 2, pn


# Inverse of the system of a constant shared mesh, kept per mesh so that a clip inverts it once: per
# call only W^-1 is applied (4x4 mesh: a ~3 us product instead of a 23 us factorisation; 16x16: instead of 7.6 ms).
# The entry holds the mesh storage alive, so its address cannot be recycled, and is keyed on the tensor version, so an
# in-place update of the mesh through torch invalidates it.  Meshes that arrive through DLPack (cupy / jax / numba
# producers: every import is a fresh tensor with version 0, and the producer may rewrite the memory behind torch's back)
# are never cached: their system is inverted on every call.  The inverse is produced on the stream of the call that
# missed; a later call on another stream waits for that event first.
_prepared = collections.OrderedDict()


def _prepare(lib, cbuf, B, pn, nbytes):
    ws = torch.empty(nbytes, dtype=torch.uint8, device=cbuf.device)
    with torch.cuda.device(cbuf.device):
        sp = stream_ptr(cbuf.device)
        _lib.check(lib.dvsg_tps_prepare(ptr(cbuf), 0, B, pn, ptr(ws), nbytes, sp), 'dvsg_tps_prepare')
        singular = ctypes.c_int(0)
        _lib.check(lib.dvsg_tps_prepare_status(ptr(ws), nbytes, B, pn, 0, sp, ctypes.byref(singular)), 'dvsg_tps_prepare_status')
    if singular.value:
        # tf.matrix_inverse (ThinPlateSpline.py:159) raises InvalidArgumentError("Input is not invertible.") here
        raise ValueError('ThinPlateSpline: the TPS system of this control mesh is not invertible (duplicate control points?)')
    return ws


def _prepared_workspace(lib, cbuf, cstride, B, pn, cache=True):
    """(workspace holding W^-1 for this shared mesh, its size), or (None, 0) for per-frame meshes (plain solve)."""
    if cstride != 0:
        return None, 0
    nbytes = lib.dvsg_tps_prepare_workspace_bytes(B, pn, 0)
    if not cache:
        return _prepare(lib, cbuf, B, pn, nbytes), nbytes
    st = cbuf.untyped_storage()
    key = (cbuf.device.index, st.data_ptr(), cbuf.storage_offset(), cbuf._version, pn)
    hit = _prepared.get(key)
    cur = torch.cuda.current_stream(cbuf.device)
    if hit is None:
        ws = _prepare(lib, cbuf, B, pn, nbytes)
        ev = torch.cuda.Event()
        ev.record(cur)
        _prepared[key] = hit = (ws, st, ev, cur.cuda_stream)
        while len(_prepared) > 4:
            _prepared.popitem(last=False)
    else:
        _prepared.move_to_end(key)
        if hit[3] != cur.cuda_stream:
            cur.wait_event(hit[2])
    return hit[0], nbytes


# ---- K1 ------------------------------------------------------------------------------------
def tps_solve(coord, target, offsets=False, cache_mesh=None):
    """_solve_system (ThinPlateSpline.py:143-166) -> T [B, 2, pn+3].  offsets=True: `target` holds the regressed offsets
    `vector` and the right-hand side coord + vector (ThinPlateSpline.py:161) is formed inside the solve (same fp32 add,
    one launch less) when the mesh is shared by the batch; per-frame meshes add here."""
    lib = _lib.load()
    target = as_cuda_f32(target, 'target')
    B = target.shape[0]
    if cache_mesh is None:
        cache_mesh = isinstance(coord, torch.Tensor)      # DLPack imports are never cached (see _prepared)
    coord = _as_mesh(coord, target)
    cbuf, cstride, pn = _mesh_args(coord, B)
    if tuple(target.shape) != (B, pn, 2):
        raise ValueError('target must have shape [B, num_point, 2], got %s' % (tuple(target.shape),))
    if pn < 3:
        raise ValueError('TPS needs at least 3 control points, got %d' % pn)
    T = torch.empty((B, 2, pn + 3), dtype=torch.float32, device=target.device)
    pws, pbytes = _prepared_workspace(lib, cbuf, cstride, B, pn, cache_mesh) if B > 0 else (None, 0)
    if offsets and pws is None and B > 0:
        target = coord.to(target.device) + target
    with torch.cuda.device(target.device):
        if pws is not None and offsets:
            rc = lib.dvsg_tps_solve_offsets_prepared(ptr(cbuf), cstride, ptr(target), ptr(T), B, pn, ptr(pws), pbytes, stream_ptr(target.device))
        elif pws is not None:
            rc = lib.dvsg_tps_solve_prepared(ptr(cbuf), cstride, ptr(target), ptr(T), B, pn, ptr(pws), pbytes, stream_ptr(target.device))
        else:
            nbytes = lib.dvsg_tps_solve_workspace_bytes(B, pn, cstride)
            ws = _workspace(target.device, nbytes)
            rc = lib.dvsg_tps_solve(ptr(cbuf), cstride, ptr(target), ptr(T), B, pn, ptr(ws), nbytes, stream_ptr(target.device))
    _lib.check(rc, 'dvsg_tps_solve')
    return T


def tps_solve_bwd(coord, grad_T, cache_mesh=True):
    lib = _lib.load()
    grad_T = as_cuda_f32(grad_T, 'grad_T')
    B, _, N = grad_T.shape
    cbuf, cstride, pn = _mesh_args(coord, B, N - 3)
    g = torch.empty((B, pn, 2), dtype=torch.float32, device=grad_T.device)
    pws, pbytes = _prepared_workspace(lib, cbuf, cstride, B, pn, cache_mesh) if B > 0 else (None, 0)
    with torch.cuda.device(grad_T.device):
        if pws is not None:
            rc = lib.dvsg_tps_solve_bwd_prepared(ptr(cbuf), cstride, ptr(grad_T), ptr(g), B, pn, ptr(pws), pbytes, stream_ptr(grad_T.device))
        else:
            nbytes = lib.dvsg_tps_solve_workspace_bytes(B, pn, cstride)
            ws = _workspace(grad_T.device, nbytes)
            rc = lib.dvsg_tps_solve_bwd(ptr(cbuf), cstride, ptr(grad_T), ptr(g), B, pn, ptr(ws), nbytes, stream_ptr(grad_T.device))
    _lib.check(rc, 'dvsg_tps_solve_bwd')
    return g


def tps_coord_bwd(coord, T, grad_T, grad_x, grad_y, out_size):
    """Gradient w.r.t. the control points through the dense grid and the system matrix (dvsg_tps_coord_bwd; the right-hand
    side's share is the caller's).  Returns [B, pn, 2]; the mesh's inverse is computed here, uncached."""
    lib = _lib.load()
    B, _, N = T.shape
    oh, ow = out_hw(out_size)
    cbuf, cstride, pn = _mesh_args(coord.detach(), B, N - 3)
    dev = T.device
    nbytes = lib.dvsg_tps_prepare_workspace_bytes(B, pn, cstride)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    g = torch.empty((B, pn, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        sp = stream_ptr(dev)
        _lib.check(lib.dvsg_tps_prepare(ptr(cbuf), cstride, B, pn, ptr(ws), nbytes, sp), 'dvsg_tps_prepare')
        rc = lib.dvsg_tps_coord_bwd(ptr(cbuf), cstride, ptr(T), ptr(grad_T), ptr(grad_x), ptr(grad_y), ptr(g), B, oh, ow, pn, ptr(ws), nbytes, sp)
    _lib.check(rc, 'dvsg_tps_coord_bwd')
    return g


# ---- K2+K3 / K4 ------------------------------------------------------------------------------
def tps_warp_fwd(U, coord, T, out_size, want_grid=True, want_mask=False, flags=0):
    lib = _lib.load()
    B, H, W, C = U.shape
    oh, ow = out_hw(out_size)
    cbuf, cstride, pn = _mesh_args(coord, B, T.shape[2] - 3)
    dev = U.device
    out = torch.empty((B, oh, ow, C), dtype=torch.float32, device=dev)
    x = torch.empty(B * oh * ow, dtype=torch.float32, device=dev) if want_grid else None
    y = torch.empty(B * oh * ow, dtype=torch.float32, device=dev) if want_grid else None
    mask = torch.empty((B, oh, ow), dtype=torch.float32, device=dev) if want_mask else None
    with torch.cuda.device(dev):
        rc = lib.dvsg_tps_warp_fwd(ptr(U), ptr(cbuf), cstride, ptr(T), ptr(out), ptr(x), ptr(y), ptr(mask),
                                   B, H, W, C, oh, ow, pn, flags, stream_ptr(dev))
    _lib.check(rc, 'dvsg_tps_warp_fwd')
    return out, x, y, mask


def tps_warp_bwd(U, coord, T, out_size, grad_out, grad_x=None, grad_y=None, need_grad_U=True, want_grid_grad=False, grad_U_out=None, flags=0):
    lib = _lib.load()
    B, H, W, C = U.shape
    oh, ow = out_hw(out_size)
    cbuf, cstride, pn = _mesh_args(coord, B, T.shape[2] - 3)
    dev = U.device
    grad_out = as_cuda_f32(grad_out, 'grad_out', like=U)
    gx_in = as_cuda_f32(grad_x, 'grad_x', like=U) if grad_x is not None else None
    gy_in = as_cuda_f32(grad_y, 'grad_y', like=U) if grad_y is not None else None
    if (gx_in is None) != (gy_in is None):
        z = torch.zeros(B * oh * ow, dtype=torch.float32, device=dev)
        gx_in = z if gx_in is None else gx_in
        gy_in = z if gy_in is None else gy_in
    # grad_U is accumulated into: the caller may pass a zero-filled (or partially accumulated) buffer
    gU = (grad_U_out if grad_U_out is not None else torch.zeros_like(U)) if need_grad_U else None
    gT = torch.empty((B, 2, pn + 3), dtype=torch.float32, device=dev)
    gxs = torch.empty(B * oh * ow, dtype=torch.float32, device=dev) if want_grid_grad else None
    gys = torch.empty(B * oh * ow, dtype=torch.float32, device=dev) if want_grid_grad else None
    with torch.cuda.device(dev):
        # `flags` must be the flags of the forward call (DVSG_FLAG_TPS_EXACT selects the same coordinates in both)
        rc = lib.dvsg_tps_warp_bwd_ex(ptr(U), ptr(cbuf), cstride, ptr(T), ptr(grad_out), ptr(gx_in), ptr(gy_in), ptr(gU),
                                      ptr(gT), ptr(gxs), ptr(gys), B, H, W, C, oh, ow, pn, flags & _lib.FLAG_TPS_EXACT, stream_ptr(dev))
    _lib.check(rc, 'dvsg_tps_warp_bwd')
    return gU, gT, gxs, gys


class _TpsWarp(torch.autograd.Function):
    """(U, target) -> (output, x, y, mask) with coord a constant, as in every reference call site
    (model.py:62-68).  mask (N1) is the warp of an all-ones image from the same pass."""

    @staticmethod
    def forward(ctx, U, coord, target, out_size, want_grid, offsets=False, want_mask=False, cache_mesh=True):
        T = tps_solve(coord, target, offsets=offsets, cache_mesh=cache_mesh)   # offsets: `target` is the regressed `vector` (d target / d vector = I)
        out, x, y, mask = tps_warp_fwd(U, coord, T, out_size, want_grid=want_grid, want_mask=want_mask)
        ctx.save_for_backward(U, coord, T)
        ctx.out_size = out_size
        ctx.want_grid = want_grid
        ctx.cache_mesh = cache_mesh
        ctx.offsets = offsets
        if not want_grid:
            x, y = out.new_empty(0), out.new_empty(0)
            ctx.mark_non_differentiable(x, y)
        if not want_mask:
            mask = out.new_empty(0)
        # d mask / d (x, y) is identically zero: the four weights sum to (x1f - x0f) * (y1f - y0f), a piecewise
        # constant of the coordinates (and the warp of ones does not depend on U)
        ctx.mark_non_differentiable(mask)
        return out, x, y, mask

    @staticmethod
    def backward(ctx, grad_out, grad_x, grad_y, _grad_mask):
        U, coord, T = ctx.saved_tensors
        if not ctx.want_grid:
            grad_x = grad_y = None
        need_U = ctx.needs_input_grad[0]
        need_c = ctx.needs_input_grad[1]
        need_t = ctx.needs_input_grad[2]
        gU, gT, gxs, gys = tps_warp_bwd(U, coord, T, ctx.out_size, grad_out.contiguous(), grad_x, grad_y, need_grad_U=need_U,
                                        want_grid_grad=need_c)
        g_target = tps_solve_bwd(coord, gT, cache_mesh=ctx.cache_mesh) if (need_t or (need_c and ctx.offsets)) else None
        g_coord = None
        if need_c:
            # no reference caller differentiates the mesh (model.py:62-68): built for completeness (dvsg_tps_coord_bwd)
            g_coord = tps_coord_bwd(coord, T, gT, gxs, gys, ctx.out_size)
            if ctx.offsets:                       # coord also sits in the right-hand side coord + vector (ThinPlateSpline.py:161)
                g_coord = g_coord + g_target
            if coord.dim() == 2:
                g_coord = g_coord.sum(dim=0)
        return gU, g_coord, (g_target if need_t else None), None, None, None, None, None


def _tps_args(U, coord, target):
    U = as_cuda_f32(U, 'U')
    if U.dim() != 4:
        raise ValueError('U must have shape [num_batch, height, width, num_channels], got %s' % (tuple(U.shape),))
    cache_mesh = isinstance(coord, torch.Tensor)
    coord = _as_mesh(coord, U)
    target = as_cuda_f32(target, 'target', like=U)
    if coord.requires_grad:
        cache_mesh = False        # a mesh that is being optimised changes between calls: never reuse a cached inverse
    return U, coord, target, cache_mesh


def thin_plate_spline(U, coord, target, out_size, want_grid=True, offsets=False):
    """offsets=True: `target` holds the offsets `vector` of ThinPlateSpline(U, coord, vector, out_size); the sum
    coord + vector is formed inside the solve."""
    U, coord, target, cache_mesh = _tps_args(U, coord, target)
    out, x, y, _ = _TpsWarp.apply(U, coord, target, tuple(out_hw(out_size)), bool(want_grid), bool(offsets), False, cache_mesh)
    return (out, x, y) if want_grid else (out, None, None)


def thin_plate_spline_with_mask(U, coord, target, out_size, want_grid=True, offsets=False):
    """N1: the image warp and the warp of an all-ones image -- the pair of stn() calls at every
    training call site, model.py:81-85,120-121 -- from ONE pass.  Returns (output, mask, x, y).  output, x and y carry
    gradients to U and target exactly like thin_plate_spline; the mask is returned without a graph: in the reference it
    is differentiable too (trainer.py:232-243 multiplies by it), but its gradient w.r.t. the coordinates is identically
    zero (the four weights sum to a piecewise constant) and it does not depend on U."""
    U, coord, target, cache_mesh = _tps_args(U, coord, target)
    out, x, y, mask = _TpsWarp.apply(U, coord, target, tuple(out_hw(out_size)), bool(want_grid), bool(offsets), True, cache_mesh)
    mask = mask.unsqueeze(-1).expand(-1, -1, -1, U.shape[3])
    return (out, mask, x, y) if want_grid else (out, mask, None, None)


# ---- K5 --------------------------------------------------------------------------------------
class _Bilinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, im, x, y, out_size):
        lib = _lib.load()
        B, H, W, C = im.shape
        oh, ow = out_size
        out = torch.empty((B * oh * ow, C), dtype=torch.float32, device=im.device)
        with torch.cuda.device(im.device):
            rc = lib.dvsg_bilinear_fwd(ptr(im), ptr(x), ptr(y), ptr(out), B, H, W, C, oh, ow, 0, stream_ptr(im.device))
        _lib.check(rc, 'dvsg_bilinear_fwd')
        ctx.save_for_backward(im, x, y)
        ctx.out_size = out_size
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        im, x, y = ctx.saved_tensors
        B, H, W, C = im.shape
        oh, ow = ctx.out_size
        grad_out = grad_out.contiguous()
        g_im = torch.zeros_like(im) if ctx.needs_input_grad[0] else None
        need_xy = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        gx = torch.empty_like(x) if need_xy else None
        gy = torch.empty_like(y) if need_xy else None
        with torch.cuda.device(im.device):
            rc = lib.dvsg_bilinear_bwd(ptr(im), ptr(x), ptr(y), ptr(grad_out), ptr(g_im), ptr(gx), ptr(gy),
                                       B, H, W, C, oh, ow, stream_ptr(im.device))
        _lib.check(rc, 'dvsg_bilinear_bwd')
        return g_im, gx, gy, None


def bilinear_interp(im, x, y, out_size):
    im = as_cuda_f32(im, 'im')
    if im.dim() != 4:
        raise ValueError('im must have shape [batch, height, width, channels], got %s' % (tuple(im.shape),))
    oh, ow = out_hw(out_size)
    x = as_cuda_f32(x, 'x', like=im).reshape(-1)
    y = as_cuda_f32(y, 'y', like=im).reshape(-1)
    n = im.shape[0] * oh * ow
    if x.numel() != n or y.numel() != n:
        raise ValueError('x and y must hold batch*out_height*out_width = %d coordinates, got %d and %d' % (n, x.numel(), y.numel()))
    return _Bilinear.apply(im, x, y, (oh, ow))


# ---- K6 --------------------------------------------------------------------------------------
class _FlowWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, im, flow):
        lib = _lib.load()
        B, H, W, C = im.shape
        out = torch.empty_like(im)
        with torch.cuda.device(im.device):
            rc = lib.dvsg_flow_warp_fwd(ptr(im), ptr(flow), ptr(out), B, H, W, C, 0, stream_ptr(im.device))
        _lib.check(rc, 'dvsg_flow_warp_fwd')
        ctx.save_for_backward(im, flow)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        im, flow = ctx.saved_tensors
        B, H, W, C = im.shape
        grad_out = grad_out.contiguous()
        g_im = torch.zeros_like(im) if ctx.needs_input_grad[0] else None
        g_flow = torch.empty_like(flow) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(im.device):
            rc = lib.dvsg_flow_warp_bwd(ptr(im), ptr(flow), ptr(grad_out), ptr(g_im), ptr(g_flow), B, H, W, C, stream_ptr(im.device))
        _lib.check(rc, 'dvsg_flow_warp_bwd')
        return g_im, g_flow


def flow_warp(im, flow, out_height, out_width):
    im = as_cuda_f32(im, 'im')
    flow = as_cuda_f32(flow, 'flow', like=im)
    if im.dim() != 4:
        raise ValueError('im must have shape [batch, height, width, channels], got %s' % (tuple(im.shape),))
    B, H, W, _ = im.shape
    if tuple(flow.shape) != (B, H, W, 2):
        raise ValueError('flow must have shape [batch, height, width, 2] matching im, got %s' % (tuple(flow.shape),))
    if (int(out_height), int(out_width)) != (H, W):
        # the reference builds its gather base from out_height*out_width while x, y come from im's
        # own size (warp_with_optical_flow.py:107,148): any other value is ill-formed there too
        raise ValueError('tf_warp requires out_height, out_width == im height, width (%d, %d)' % (H, W))
    return _FlowWarp.apply(im, flow)


# ---- B1 / N2 ---------------------------------------------------------------------------------
def st_meshgrid(out_size, device=None):
    lib = _lib.load()
    oh, ow = out_hw(out_size)
    device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    grid = torch.empty(3 * oh * ow, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        rc = lib.dvsg_st_meshgrid(ptr(grid), oh, ow, stream_ptr(device))
    _lib.check(rc, 'dvsg_st_meshgrid')
    return grid


class _HomographyWarp(torch.autograd.Function):
    """(inp, theta) -> (out, x_s, y_s) of ProjectiveTransformer / AffineTransformer (spatial_transformer.py:384-452, 31-91).
    Backward: the sampler's (dvsg_bilinear_bwd) chained through the grid stage (dvsg_homography_grid_bwd)."""

    @staticmethod
    def forward(ctx, inp, theta, out_size, projective, want_grid):
        lib = _lib.load()
        B, H, W, C = inp.shape
        oh, ow = out_size
        need_bwd = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        out = torch.empty((B, oh, ow, C), dtype=torch.float32, device=inp.device)
        keep = want_grid or need_bwd
        x = torch.empty(B * oh * ow, dtype=torch.float32, device=inp.device) if keep else None
        y = torch.empty(B * oh * ow, dtype=torch.float32, device=inp.device) if keep else None
        with torch.cuda.device(inp.device):
            rc = lib.dvsg_homography_warp_fwd(ptr(inp), ptr(theta), 1 if projective else 0, ptr(out), ptr(x), ptr(y),
                                              B, H, W, C, oh, ow, stream_ptr(inp.device))
        _lib.check(rc, 'dvsg_homography_warp_fwd')
        ctx.out_size, ctx.projective = out_size, projective
        if need_bwd:
            ctx.save_for_backward(inp, theta, x, y)
        if not keep:
            x, y = out.new_empty(0), out.new_empty(0)
            ctx.mark_non_differentiable(x, y)
        return out, x, y

    @staticmethod
    def backward(ctx, grad_out, grad_x, grad_y):
        lib = _lib.load()
        inp, theta, x, y = ctx.saved_tensors
        B, H, W, C = inp.shape
        oh, ow = ctx.out_size
        dev = inp.device
        grad_out = grad_out.contiguous()
        g_im = torch.zeros_like(inp) if ctx.needs_input_grad[0] else None
        g_theta = None
        need_theta = ctx.needs_input_grad[1]
        gx = torch.empty_like(x) if need_theta else None
        gy = torch.empty_like(y) if need_theta else None
        with torch.cuda.device(dev):
            rc = lib.dvsg_bilinear_bwd(ptr(inp), ptr(x), ptr(y), ptr(grad_out), ptr(g_im), ptr(gx), ptr(gy), B, H, W, C, oh, ow, stream_ptr(dev))
            _lib.check(rc, 'dvsg_bilinear_bwd')
            if need_theta:
                if grad_x is not None and grad_x.numel():          # x_s / y_s consumed directly (_transform)
                    gx = gx + grad_x.reshape(-1)
                if grad_y is not None and grad_y.numel():
                    gy = gy + grad_y.reshape(-1)
                g_theta = torch.empty_like(theta)
                rc = lib.dvsg_homography_grid_bwd(ptr(theta), ptr(gx), ptr(gy), 1 if ctx.projective else 0, ptr(g_theta), B, oh, ow,
                                                  stream_ptr(dev))
                _lib.check(rc, 'dvsg_homography_grid_bwd')
        return g_im, g_theta, None, None, None


def homography_warp(inp, theta, out_size, projective, want_grid=False):
    inp = as_cuda_f32(inp, 'inp')
    if inp.dim() != 4:
        raise ValueError('inp must have shape [batch, height, width, channels], got %s' % (tuple(inp.shape),))
    B = inp.shape[0]
    nt = 8 if projective else 6
    theta = as_cuda_f32(theta, 'theta', like=inp).reshape(B, nt)
    out, x, y = _HomographyWarp.apply(inp, theta, out_hw(out_size), bool(projective), bool(want_grid))
    return (out, x, y) if want_grid else out


# ---- online loop -----------------------------------------------------------------------------
class OnlineWarper(object):
    """ThinPlateSpline for the online loop of eval.py:106-110 (one small batch per call, constant mesh, inference only):
    the mesh's system is inverted once, buffers are allocated once, and each call is ONE trip through the C ABI
    (dvsg_tps_warp_frames_offsets = prepared solve of coord + vector + fused warp).  Same results as
    ThinPlateSpline(U, coord, vector, [h, w])."""

    def __init__(self, mesh, B, H, W, C=3, device=None):
        lib = _lib.load()
        self.mesh = as_cuda_f32(mesh, 'mesh').reshape(-1, 2).contiguous()
        dev = self.mesh.device if device is None else torch.device(device)
        self.B, self.H, self.W, self.C, self.pn = int(B), int(H), int(W), int(C), self.mesh.shape[0]
        self.nbytes = lib.dvsg_tps_prepare_workspace_bytes(self.B, self.pn, 0)
        self.ws = torch.empty(self.nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.dvsg_tps_prepare(ptr(self.mesh), 0, self.B, self.pn, ptr(self.ws), self.nbytes, stream_ptr(dev))
        _lib.check(rc, 'dvsg_tps_prepare')
        self.T = torch.empty((self.B, 2, self.pn + 3), dtype=torch.float32, device=dev)
        self.out = torch.empty((self.B, self.H, self.W, self.C), dtype=torch.float32, device=dev)
        self._fn = lib.dvsg_tps_warp_frames_offsets
        self._mesh_ptr = ptr(self.mesh)
        self._args = (ptr(self.ws), self.nbytes, ptr(self.T), ptr(self.out), None, None, None,
                      self.B, self.H, self.W, self.C, self.H, self.W, self.pn)
        self._dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
        self._U_shape = (self.B, self.H, self.W, self.C)
        self._v_shape = (self.B, self.pn, 2)
        # raw handle of the current stream without constructing a torch.cuda.Stream per frame (the per-frame host cost of
        # the online loop is what bounds it: the kernel itself lasts ~6 us)
        self._raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)

    def warp(self, U, vector):
        """U [B,H,W,C] and vector [B,pn,2]: contiguous fp32 CUDA tensors.  Returns the warped frames (a buffer owned by this
        object, overwritten by the next call).  One C-ABI call = one kernel launch per call for meshes up to 29 points."""
        if U.shape != self._U_shape or U.dtype is not torch.float32 or not U.is_cuda or not U.is_contiguous():
            raise ValueError('U must be a contiguous fp32 CUDA tensor of shape %r' % (self._U_shape,))
        if vector.shape != self._v_shape or vector.dtype is not torch.float32 or not vector.is_cuda or not vector.is_contiguous():
            raise ValueError('vector must be a contiguous fp32 CUDA tensor of shape %r' % (self._v_shape,))
        # coord + vector (ThinPlateSpline.py:161) is formed inside the prepared solve, which runs in the warp kernel's prologue
        if torch.cuda.current_device() != self._dev_index:
            with torch.cuda.device(self._dev_index):
                return self._call(U, vector)
        return self._call(U, vector)

    def _call(self, U, vector):
        st = self._raw_stream(self._dev_index) if self._raw_stream is not None else torch.cuda.current_stream(self.out.device).cuda_stream
        rc = self._fn(U.data_ptr(), self._mesh_ptr, vector.data_ptr(), *self._args, st)
        if rc:
            _lib.check(rc, 'dvsg_tps_warp_frames_offsets')
        return self.out


# ---- host-buffer pipeline --------------------------------------------------------------------
class HostPipeline(object):
    """ThinPlateSpline on HOST buffers: H2D, solve, fused warp and D2H overlapped across
    `n_slots` staging slots (the e2e path measured by bench.py)."""

    def __init__(self, H, W, C, pn, frames_per_chunk=8, n_slots=3, device=None):
        lib = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.shape = (H, W, C)
        self.pn = pn
        h = ctypes.c_void_p()
        rc = lib.dvsg_host_pipeline_create(ctypes.byref(h), self.device, H, W, C, pn, frames_per_chunk, n_slots)
        if rc != 0:
            if h.value:
                lib.dvsg_host_pipeline_destroy(h)
            _lib.check(rc, 'dvsg_host_pipeline_create')
        self._h = h

    def thin_plate_spline(self, U_host, coord_host, vector_host, out_host=None):
        """U_host [B,H,W,C], coord_host [pn,2], vector_host [B,pn,2]: CPU fp32 contiguous torch
        tensors (pinned for full PCIe bandwidth).  Returns out_host."""
        lib = _lib.load()
        for name, t in (('U', U_host), ('coord', coord_host), ('vector', vector_host)):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError('%s must be a contiguous fp32 CPU tensor' % name)
        B = U_host.shape[0]
        if tuple(U_host.shape[1:]) != self.shape or tuple(coord_host.shape) != (self.pn, 2) or tuple(vector_host.shape) != (B, self.pn, 2):
            raise ValueError('shape mismatch with the pipeline configuration')
        if out_host is None:
            out_host = torch.empty_like(U_host, pin_memory=U_host.is_pinned())
        rc = lib.dvsg_host_tps_warp(self._h, U_host.data_ptr(), coord_host.data_ptr(), vector_host.data_ptr(), out_host.data_ptr(), B)
        _lib.check(rc, 'dvsg_host_tps_warp')
        return out_host

    def thin_plate_spline_u8(self, U_host, coord_host, vector_host, out_host=None, swap_rb=True):
        """Same call with uint8 frames on the host side: U_host [B,H,W,3] uint8 (BGR when swap_rb,
        as cv2 delivers it) -> ingest u/255 on the device -> fp32 ThinPlateSpline -> egress uint8(x*255) -> out_host
        uint8.  3 bytes per pixel cross PCIe each way instead of 12."""
        lib = _lib.load()
        if U_host.is_cuda or U_host.dtype != torch.uint8 or not U_host.is_contiguous():
            raise ValueError('U must be a contiguous uint8 CPU tensor')
        for name, t in (('coord', coord_host), ('vector', vector_host)):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError('%s must be a contiguous fp32 CPU tensor' % name)
        B = U_host.shape[0]
        if tuple(U_host.shape[1:]) != self.shape or tuple(coord_host.shape) != (self.pn, 2) or tuple(vector_host.shape) != (B, self.pn, 2):
            raise ValueError('shape mismatch with the pipeline configuration')
        if out_host is None:
            out_host = torch.empty_like(U_host, pin_memory=U_host.is_pinned())
        elif out_host.is_cuda or out_host.dtype != torch.uint8 or not out_host.is_contiguous() or out_host.shape != U_host.shape:
            raise ValueError('out must be a contiguous uint8 CPU tensor shaped like U')
        rc = lib.dvsg_host_tps_warp_u8(self._h, U_host.data_ptr(), coord_host.data_ptr(), vector_host.data_ptr(), out_host.data_ptr(),
                                       B, 1 if swap_rb else 0)
        _lib.check(rc, 'dvsg_host_tps_warp_u8')
        return out_host

    def tf_warp(self, im_host, flow_host, out_host=None):
        """tf_warp(im, flow, H, W) on HOST buffers: im_host [B,H,W,C], flow_host [B,H,W,2] CPU fp32 contiguous (pinned for full
        PCIe bandwidth).  Returns out_host."""
        lib = _lib.load()
        for name, t in (('im', im_host), ('flow', flow_host)):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError('%s must be a contiguous fp32 CPU tensor' % name)
        B = im_host.shape[0]
        if tuple(im_host.shape[1:]) != self.shape or tuple(flow_host.shape) != (B, self.shape[0], self.shape[1], 2):
            raise ValueError('shape mismatch with the pipeline configuration')
        if out_host is None:
            out_host = torch.empty_like(im_host, pin_memory=im_host.is_pinned())
        rc = lib.dvsg_host_flow_warp(self._h, im_host.data_ptr(), flow_host.data_ptr(), out_host.data_ptr(), B)
        _lib.check(rc, 'dvsg_host_flow_warp')
        return out_host

    def set_async(self, flag=True):
        """Streaming mode: calls return once enqueued, so consecutive batches overlap (upload of batch k+1 with the download of
        batch k); sync() waits.  Host buffers of a call must stay valid, and its outputs unread, until sync()."""
        _lib.check(_lib.load().dvsg_host_pipeline_set_async(self._h, 1 if flag else 0), 'dvsg_host_pipeline_set_async')

    def sync(self):
        _lib.check(_lib.load().dvsg_host_pipeline_sync(self._h), 'dvsg_host_pipeline_sync')

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            _lib.load().dvsg_host_pipeline_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
