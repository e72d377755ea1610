"""experiment: tile vs generic backward kernels by shape and mode"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
exec(open(os.path.join(os.path.dirname(__file__), 'sweep.py')).read().split("def main():")[0])
s = torch.cuda.current_stream().cuda_stream
for (B, H, W) in ((32, 288, 512), (16, 720, 1280), (16, 1080, 1920)):
    U = torch.rand((B, H, W, 3), device=dev); g = torch.rand_like(U); gU = torch.zeros_like(U)
    for name, flow in (('zero', torch.zeros((B, H, W, 2), device=dev)), ('smooth no jitter', smooth_flow(B, H, W, 8.0, 0.0)), ('smooth jitter 0.5', smooth_flow(B, H, W, 8.0, 0.5))):
        def fb():
            rc = lib.dvsg_flow_warp_bwd(U.data_ptr(), flow.data_ptr(), g.data_ptr(), gU.data_ptr(), None, B, H, W, 3, s)
            assert rc == 0
        a = timeit(fb)
        lib.dvsg_set_bwd_tuning(1 | 2)
        b = timeit(fb)
        lib.dvsg_set_bwd_tuning(1)
        print('%dx%dx%d flow bwd %-20s tile %.4f ms  generic %.4f ms' % (B, H, W, name, a, b))
    Uc, coord, T = tps_case(B, H, W, 4, 0.2)
    for want_U in (True, False):
        def tb():
            ops.tps_warp_bwd(Uc, coord, T, (H, W), g, None, None, need_grad_U=want_U, want_grid_grad=True, grad_U_out=gU if want_U else None)
        a = timeit(tb)
        lib.dvsg_set_bwd_tuning(1 | 2)
        b = timeit(tb)
        lib.dvsg_set_bwd_tuning(1)
        print('%dx%dx%d tps bwd grad_U=%-5s          tile %.4f ms  generic %.4f ms' % (B, H, W, want_U, a, b))
    del U, g, gU, Uc
