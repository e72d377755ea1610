"""experiment: cost of the backward kernel without its grad_T pass (P2)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coupe.dvsg_b200 import _lib, ops
exec(open(os.path.join(os.path.dirname(__file__), 'sweep.py')).read().split("def main():")[0])
B, H, W = 32, 288, 512
U, coord, T = tps_case(B, H, W, 4, 0.2)
g = torch.rand((B, H, W, 3), device=dev)
gU = torch.zeros_like(U); gT = torch.empty((B, 2, 19), device=dev); gx = torch.empty(B * H * W, device=dev); gy = torch.empty_like(gx)
cb = coord[0].contiguous()
s = torch.cuda.current_stream().cuda_stream
def run(gT_ptr, gx_ptr, gy_ptr):
    rc = lib.dvsg_tps_warp_bwd(U.data_ptr(), cb.data_ptr(), 0, T.data_ptr(), g.data_ptr(), None, None, gU.data_ptr(), gT_ptr, gx_ptr, gy_ptr, B, H, W, 3, H, W, 16, s)
    assert rc == 0
for name, a in (('full (grad_U, grad_T, grad_xy)', (gT.data_ptr(), gx.data_ptr(), gy.data_ptr())), ('no grad_T', (None, gx.data_ptr(), gy.data_ptr())),
                ('grad_T, no grad_xy', (gT.data_ptr(), None, None))):
    ms = timeit(lambda: run(*a))
    print('%-40s %.4f ms' % (name, ms))
