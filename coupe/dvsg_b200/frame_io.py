"""Frame ingest / egress on the device.

The reference's inference loop converts every frame on the CPU (eval.py:76-81: BGR uint8 -> RGB, / 255.) and converts
the stabilised frame back (eval.py:112-113: np.uint8(x * 255.), RGB -> BGR).  These are the same two conversions as
CUDA kernels behind the C ABI (dvsg_frames_u8_to_f32 / dvsg_frames_f32_to_u8), so frames can cross PCIe as uint8.
read_frames() also performs the resize of eval.py:80 (cv2.resize of the float64 frame, INTER_LINEAR) on the device;
read_flow() is the optical-flow ingest of data_loader.py:239.
"""
import torch

from . import _lib
from ._tensors import stream_ptr


def _check_u8(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError('%s: expected a CUDA torch tensor (no CPU fallback on this path)' % name)
    if t.shape[-1] != 3:
        raise ValueError('%s: frames are [..., 3], got %r' % (name, tuple(t.shape)))
    return t.contiguous()


def frames_u8_to_f32(frames_u8, swap_rb=True):
    """uint8 [..., H, W, 3] (BGR when swap_rb) -> fp32 RGB in [0, 1] = u / 255  (eval.py:79-80)."""
    f = _check_u8(frames_u8, 'frames_u8')
    if f.dtype != torch.uint8:
        raise ValueError('frames_u8: dtype must be uint8, got %s' % f.dtype)
    out = torch.empty(f.shape, dtype=torch.float32, device=f.device)
    with torch.cuda.device(f.device):
        rc = _lib.load().dvsg_frames_u8_to_f32(f.data_ptr(), out.data_ptr(), f.numel() // 3, 1 if swap_rb else 0, stream_ptr(f.device))
    _lib.check(rc, 'dvsg_frames_u8_to_f32')
    return out


def frames_f32_to_u8(frames_f32, swap_rb=True):
    """fp32 [..., H, W, 3] RGB -> uint8 (BGR when swap_rb) = uint8(x * 255.)  (eval.py:112-113)."""
    f = _check_u8(frames_f32, 'frames_f32')
    if f.dtype != torch.float32:
        raise ValueError('frames_f32: dtype must be float32, got %s' % f.dtype)
    out = torch.empty(f.shape, dtype=torch.uint8, device=f.device)
    with torch.cuda.device(f.device):
        rc = _lib.load().dvsg_frames_f32_to_u8(f.data_ptr(), out.data_ptr(), f.numel() // 3, 1 if swap_rb else 0, stream_ptr(f.device))
    _lib.check(rc, 'dvsg_frames_f32_to_u8')
    return out


def read_frames(frames_u8, out_size, swap_rb=True):
    """read_frame (eval.py:76-81) for a batch: uint8 [B, Hs, Ws, 3] (BGR when swap_rb) -> fp32 RGB [B, h, w, 3] =
    cv2.resize(cv2.cvtColor(frame, BGR2RGB) / 255., (w, h)), cast to fp32 as the feed does.  out_size = (h, w)."""
    f = _check_u8(frames_u8, 'frames_u8')
    if f.dtype != torch.uint8 or f.dim() != 4:
        raise ValueError('frames_u8: expected a uint8 [B, H, W, 3] tensor, got %s %r' % (f.dtype, tuple(f.shape)))
    h, w = int(out_size[0]), int(out_size[1])
    B, Hs, Ws = f.shape[0], f.shape[1], f.shape[2]
    out = torch.empty((B, h, w, 3), dtype=torch.float32, device=f.device)
    with torch.cuda.device(f.device):
        rc = _lib.load().dvsg_frames_u8_resize_to_f32(f.data_ptr(), out.data_ptr(), B, Hs, Ws, h, w, 1 if swap_rb else 0, stream_ptr(f.device))
    _lib.check(rc, 'dvsg_frames_u8_resize_to_f32')
    return out


def read_flow(flow, out_size):
    """The optical-flow ingest of data_loader.py:239, `cv2.resize(np.load(of_file), (w, h)) * [w, h]`, on the device: flow is
    the stored float32 field of normalised displacements, [Hs, Ws, 2] or a batch [B, Hs, Ws, 2]; the result is in pixels at
    the working size out_size = (h, w) -- the `of` argument of tf_warp (channel 0 = dx, 1 = dy)."""
    if not isinstance(flow, torch.Tensor) or not flow.is_cuda:
        raise ValueError('flow: expected a CUDA torch tensor (no CPU fallback on this path)')
    if flow.dtype != torch.float32 or flow.shape[-1] != 2 or flow.dim() not in (3, 4):
        raise ValueError('flow: expected a float32 [Hs, Ws, 2] or [B, Hs, Ws, 2] tensor, got %s %r' % (flow.dtype, tuple(flow.shape)))
    single = flow.dim() == 3
    f = (flow.unsqueeze(0) if single else flow).contiguous()
    h, w = int(out_size[0]), int(out_size[1])
    B, Hs, Ws = f.shape[0], f.shape[1], f.shape[2]
    out = torch.empty((B, h, w, 2), dtype=torch.float32, device=f.device)
    with torch.cuda.device(f.device):
        rc = _lib.load().dvsg_flow_resize_scale(f.data_ptr(), out.data_ptr(), B, Hs, Ws, h, w, stream_ptr(f.device))
    _lib.check(rc, 'dvsg_flow_resize_scale')
    return out[0] if single else out
