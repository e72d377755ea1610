import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coupe.dvsg_b200 import ops, _lib
from tools.sweep import timeit, tps_case, smooth_flow, dev
lib = _lib.load()
what = sys.argv[1]
if what == 'tps':
    B, H, W = 32, 288, 512
    for amp in (0.0, 0.2):
        U, coord, T = tps_case(B, H, W, 4, amp)
        g = torch.rand_like(U)
        for notile in (0, 1):
            lib.dvsg_set_bwd_tuning(1 | (notile << 1))
            ms = timeit(lambda: ops.tps_warp_bwd(U, coord, T, (H, W), g, None, None, need_grad_U=True, want_grid_grad=True))
            print('tps bwd amp=%.1f notile=%d  %.3f ms (incl. zero fill of grad_U)' % (amp, notile, ms), flush=True)
            ms = timeit(lambda: ops.tps_warp_bwd(U, coord, T, (H, W), g, None, None, need_grad_U=False, want_grid_grad=True))
            print('tps bwd amp=%.1f notile=%d  %.3f ms (no grad_U)' % (amp, notile, ms), flush=True)
        a = ops.tps_warp_bwd(U, coord, T, (H, W), g, None, None, need_grad_U=True, want_grid_grad=True)
        lib.dvsg_set_bwd_tuning(1)
        b = ops.tps_warp_bwd(U, coord, T, (H, W), g, None, None, need_grad_U=True, want_grid_grad=True)
        for x, y, n in zip(a, b, ('gU', 'gT', 'gxs', 'gys')):
            print(n, float((x - y).abs().max()), float(y.abs().max()))
elif what == 'given':
    B, H, W = 2, 72, 100
    lo, hi = float(sys.argv[2]), float(sys.argv[3])
    im = torch.rand((B, H, W, 3), device=dev, requires_grad=True)
    gx, gy = torch.meshgrid(torch.linspace(lo, hi, W, device=dev), torch.linspace(lo, hi, H, device=dev), indexing='xy')
    x = gx.reshape(-1).repeat(B).contiguous().requires_grad_(True)
    y = gy.reshape(-1).repeat(B).contiguous().requires_grad_(True)
    out = ops.bilinear_interp(im, x, y, (H, W))
    out.sum().backward()
    torch.cuda.synchronize()
    print('given ok', lo, hi, float(im.grad.sum()), float(x.grad.abs().sum()))
