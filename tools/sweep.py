"""Times the forward kernel variants with CUDA events (tile kernel vs the generic direct kernel) -- run on the GPU box."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coupe.dvsg_b200 import _lib, ops

lib = _lib.load()
dev = torch.device('cuda', 0)


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) for x, y in evs)
    return ts[len(ts) // 2]


def tps_case(B, H, W, m, amp=0.2):
    torch.manual_seed(0)
    U = torch.rand((B, H, W, 3), device=dev)
    lin = torch.arange(m, device=dev, dtype=torch.float32) * (2.0 / (m - 1)) - 1.0
    mesh = torch.stack(torch.meshgrid(lin, lin, indexing='xy'), dim=-1).reshape(m * m, 2).contiguous()
    coord = mesh.unsqueeze(0).expand(B, -1, -1)
    vec = (torch.rand((B, m * m, 2), device=dev) - 0.5) * amp
    T = ops.tps_solve(coord, coord + vec)
    return U, coord, T


def smooth_flow(B, H, W, amp=8.0, jitter=0.5, lattice=(9, 16)):
    lat = (torch.rand((B, 2) + tuple(lattice), device=dev) - 0.5) * 2 * amp
    f = torch.nn.functional.interpolate(lat, size=(H, W), mode='bilinear', align_corners=True)
    f = f + (torch.rand((B, 2, H, W), device=dev) - 0.5) * 2 * jitter
    return f.permute(0, 2, 3, 1).contiguous()


res = []


def rec(name, ms, px, bpp):
    gbs = px * bpp / ms / 1e6
    res.append((name, ms, gbs))
    print('%-56s %8.3f ms  %8.1f GB/s  %5.1f%%  %7.1f Gpix/s' % (name, ms, gbs, 100 * gbs / 6548.2, px / ms / 1e6), flush=True)


def flow_call(im, flow, out, flags):
    rc = lib.dvsg_flow_warp_fwd(im.data_ptr(), flow.data_ptr(), out.data_ptr(), im.shape[0], im.shape[1], im.shape[2], 3, flags, 0)
    assert rc == 0, lib.dvsg_last_error()


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else 'all'
    B, H, W = 64, 720, 1280
    px = B * H * W
    if only in ('all', 'tps'):
        for amp in (0.2, 0.04, 0.0, 0.4, 0.6):
            U, coord, T = tps_case(B, H, W, 4, amp)
            ms = timeit(lambda: ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=False))
            rec('tps720 4x4 amp=%.2f tile' % amp, ms, px, 24)
            if amp >= 0.4:
                lib.dvsg_set_tile_tuning(4, -1, -1)
                ms = timeit(lambda: ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=False))
                rec('tps720 4x4 amp=%.2f tile, no L2 prefetch of per-pixel tiles' % amp, ms, px, 24)
                lib.dvsg_set_tile_tuning(0, -1, -1)
            if amp == 0.2:
                ms = timeit(lambda: ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=True))
                rec('tps720 4x4 tile +xy', ms, px, 32)
                ms = timeit(lambda: ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=False, want_mask=True))
                rec('tps720 4x4 tile +mask', ms, px, 28)
                ms = timeit(lambda: ops.tps_warp_fwd(U, coord, T, (H, W), want_grid=False, flags=1))
                rec('tps720 4x4 generic direct kernel', ms, px, 24)
        U5, c5, T5 = tps_case(B, H, W, 5)
        ms = timeit(lambda: ops.tps_warp_fwd(U5, c5, T5, (H, W), want_grid=False))
        rec('tps720 5x5 tile', ms, px, 24)
        del U, U5
        U8, c8, T8 = tps_case(64, 1080, 1920, 4)
        ms = timeit(lambda: ops.tps_warp_fwd(U8, c8, T8, (1080, 1920), want_grid=False))
        rec('tps1080 4x4 B=64 tile', ms, 64 * 1080 * 1920, 24)
        del U8
        Uc, cc, Tc = tps_case(4, 2160, 3840, 16)
        ms = timeit(lambda: ops.tps_warp_fwd(Uc, cc, Tc, (2160, 3840), want_grid=False), n=5, warm=1)
        rec('tps4k 16x16 B=4 tile', ms, 4 * 2160 * 3840, 24)
        del Uc
        Us, cs, Ts = tps_case(32, 288, 512, 4)
        ms = timeit(lambda: ops.tps_warp_fwd(Us, cs, Ts, (288, 512), want_grid=True))
        rec('tps 288x512 B=32 +xy (L2-resident) tile', ms, 32 * 288 * 512, 32)
        U1, c1, T1 = tps_case(1, 288, 512, 4)
        ms = timeit(lambda: ops.tps_warp_fwd(U1, c1, T1, (288, 512), want_grid=False))
        rec('tps 288x512 B=1 tile (latency)', ms, 288 * 512, 24)
    if only in ('all', 'flow'):
        im = torch.rand((16, 1080, 1920, 3), device=dev)
        out = torch.empty_like(im)
        pxf = 16 * 1080 * 1920
        for fname, flow in (('smooth', smooth_flow(16, 1080, 1920)), ('random+-8', (torch.rand((16, 1080, 1920, 2), device=dev) - 0.5) * 16),
                            ('zero', torch.zeros((16, 1080, 1920, 2), device=dev))):
            ms = timeit(lambda: flow_call(im, flow, out, 0))
            rec('flow1080 %s tile' % fname, ms, pxf, 32)
            for name, flags in (('generic direct kernel', 1),):
                ms = timeit(lambda: flow_call(im, flow, out, flags))
                rec('flow1080 %s %s' % (fname, name), ms, pxf, 32)
        x = torch.rand(pxf, device=dev) * 2 - 1
        y = torch.rand(pxf, device=dev) * 2 - 1
        ms = timeit(lambda: ops.bilinear_interp(im, x, y, (1080, 1920)))
        rec('bilinear1080 random xy tile', ms, pxf, 32)
        ms = timeit(lambda: out.copy_(im))
        rec('torch copy 398MB', ms, pxf, 24)





def c18():
    """the only live call sites of bilinear_interp: C = 18 (model.py:156-167), generic direct kernel"""
    B, H, W, C = 32, 288, 512, 18
    im = torch.rand((B, H, W, C), device=dev)
    px = B * H * W
    theta = torch.tensor([1.02, 0.01, 0.0, -0.01, 0.98, 0.01, 1e-3, -1e-3], device=dev).repeat(B, 1)
    ms = timeit(lambda: ops.homography_warp(im, theta, (H, W), True))
    rec('projective C=18 288x512 B=32 (wide-pixel kernel)', ms, px, 8 * C)
    lat = (torch.rand((B, 2, 9, 16), device=dev) - 0.5) * 0.05
    g = torch.nn.functional.interpolate(lat, size=(H, W), mode='bilinear', align_corners=True)
    lin_x = torch.linspace(-1, 1, W, device=dev).view(1, 1, W).expand(B, H, W)
    lin_y = torch.linspace(-1, 1, H, device=dev).view(1, H, 1).expand(B, H, W)
    x = (lin_x + g[:, 0]).reshape(-1).contiguous()
    y = (lin_y + g[:, 1]).reshape(-1).contiguous()
    ms = timeit(lambda: ops.bilinear_interp(im, x, y, (H, W)))
    rec('bilinear_interp C=18 288x512 B=32 smooth grid (wide-pixel kernel)', ms, px, 8 * C + 8)
    B2 = 8
    im2 = torch.rand((B2, 1080, 1920, C), device=dev)
    theta2 = theta[:B2].contiguous()
    ms = timeit(lambda: ops.homography_warp(im2, theta2, (1080, 1920), True), n=10)
    rec('projective C=18 1080p B=8 (wide-pixel kernel, 2.4 GB per call)', ms, B2 * 1080 * 1920, 8 * C)



def bwd():
    """backward kernels (CUDA events; grad_image zero fill not included in the time, counted in the bytes as in DESIGN.md)"""
    for (B, H, W, tag) in ((32, 288, 512, 'training shape'), (16, 1080, 1920, '1080p')):
        U, coord, T = tps_case(B, H, W, 4, 0.2)
        g = torch.rand((B, H, W, 3), device=dev)
        gU = torch.zeros_like(U)
        px = B * H * W
        ms = timeit(lambda: ops.tps_warp_bwd(U, coord, T, (H, W), g, None, None, need_grad_U=True, want_grid_grad=True, grad_U_out=gU))
        rec('tps bwd 4x4 %s B=%d (grad image, grid, T)' % (tag, B), ms, px, 56)
        ms = timeit(lambda: ops.tps_warp_bwd(U, coord, T, (H, W), g, None, None, need_grad_U=False, want_grid_grad=True))
        rec('tps bwd 4x4 %s B=%d (grad grid, T only)' % (tag, B), ms, px, 32)
        flow = smooth_flow(B, H, W)
        gf = torch.empty_like(flow)
        s = torch.cuda.current_stream().cuda_stream

        def fb(want_flow):
            rc = lib.dvsg_flow_warp_bwd(U.data_ptr(), flow.data_ptr(), g.data_ptr(), gU.data_ptr(), gf.data_ptr() if want_flow else None, B, H, W, 3, s)
            assert rc == 0
        ms = timeit(lambda: fb(False))
        rec('tf_warp bwd %s B=%d smooth flow (grad image only, trainer.py:246-247)' % (tag, B), ms, px, 44)
        ms = timeit(lambda: fb(True))
        rec('tf_warp bwd %s B=%d smooth flow (grad image and flow)' % (tag, B), ms, px, 64)
        if H < 1080:
            # the default field (+-8 px on a 9 x 16 lattice) has 32-px cells at this size: slope 0.5, most tiles fit no box.
            # Same slope as the 1080p case: amplitude and lattice scaled with the frame
            flow = smooth_flow(B, H, W, amp=8.0 * H / 1080.0, jitter=0.5 * H / 1080.0, lattice=(3, 5))
            ms = timeit(lambda: fb(False))
            rec('tf_warp bwd %s B=%d flow scaled to the frame (grad image only)' % (tag, B), ms, px, 44)
            ms = timeit(lambda: fb(True))
            rec('tf_warp bwd %s B=%d flow scaled to the frame (grad image and flow)' % (tag, B), ms, px, 64)
        del U, g, gU, flow, gf


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'c18':
        c18()
    elif len(sys.argv) > 1 and sys.argv[1] == 'bwd':
        bwd()
    else:
        main()
