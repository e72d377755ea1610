"""Tensor plumbing between callers and the C ABI: DLPack in, torch tensors out.

PyTorch is used only to own device memory and streams; every computation on this path runs
in libdvsg_warp.so.
"""
import torch


def as_cuda_f32(obj, name, like=None):
    """Any DLPack producer (torch / cupy / jax / numba ...) or torch tensor -> contiguous fp32
    CUDA torch tensor, zero-copy when the producer already is one.  The reference casts
    images to float32 itself (ThinPlateSpline.py:73-74), hence the dtype conversion."""
    if not isinstance(obj, torch.Tensor):
        if hasattr(obj, '__dlpack__'):
            obj = torch.from_dlpack(obj)
        else:
            raise TypeError('%s: expected a torch.Tensor or a DLPack-capable array, got %s' % (name, type(obj).__name__))
    if not obj.is_cuda:
        raise ValueError('%s: tensor lives on %s -- this path runs on CUDA only (no CPU fallback); move it '
                         'with .cuda() or use coupe.dvsg_b200.ops.HostPipeline for host buffers' % (name, obj.device))
    if like is not None and obj.device != like.device:
        raise ValueError('%s: device %s differs from %s' % (name, obj.device, like.device))
    if obj.dtype != torch.float32:
        obj = obj.float()
    return obj.contiguous()


def ptr(t):
    return 0 if t is None else t.data_ptr()


def stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def out_hw(out_size):
    h, w = int(out_size[0]), int(out_size[1])
    if h < 0 or w < 0:
        raise ValueError('out_size must be non-negative, got %r' % (out_size,))
    return h, w
